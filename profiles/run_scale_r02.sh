#!/bin/bash
# Round-2 scaling run on one 8-GPU box (gpurun --gpus 8): outputs under gpurun_out/, copied into profiles/ afterwards.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02_bench_1gpu.json 2> $OUT/r02_bench_1gpu.err
timeout 300 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/r02_bench_8gpu_pipelined.json 2> $OUT/r02_bench_8gpu_pipelined.err
timeout 300 python bench.py --gpus 8 --inprocess --steps 5 --warmup 3 > $OUT/r02_bench_8gpu_inprocess.json 2> $OUT/r02_bench_8gpu_inprocess.err
timeout 300 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --exchange multicast --steps 3 --warmup 3 --no-e2e > $OUT/r02_bench_8gpu_multicast.json 2> $OUT/r02_bench_8gpu_multicast.err
timeout 300 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload C4 --steps 3 --warmup 2 > $OUT/r02_bench_8gpu_c4.json 2> $OUT/r02_bench_8gpu_c4.err
timeout 400 python -m pytest tests -m gpu -q -rf --timeout 380 -k "multi_gpu or ipc_pipelined or fused_exchange" > $OUT/r02_pytest_multigpu_8gpu_box.log 2>&1
tail -5 $OUT/r02_pytest_multigpu_8gpu_box.log
for f in $OUT/r02_bench_*.err; do echo "== $f"; tail -c 400 $f; done
for f in $OUT/r02_bench_*.json; do echo "== $f"; cut -c1-330 $f; done
