#!/bin/bash
# Round-2 scaling run on one 8-GPU box (gpurun --gpus 8): outputs under gpurun_out/, copied into profiles/ afterwards.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
timeout 400 python bench.py --steps 5 --warmup 3 > $OUT/r02_bench_1gpu.json 2> $OUT/r02_bench_1gpu.err
for N in 2 4 8; do
  timeout 300 $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 > $OUT/r02_bench_${N}gpu_pipelined.json 2> $OUT/r02_bench_${N}gpu_pipelined.err
done
timeout 300 python bench.py --gpus 8 --inprocess --steps 5 --warmup 3 > $OUT/r02_bench_8gpu_inprocess.json 2> $OUT/r02_bench_8gpu_inprocess.err
timeout 400 python -m pytest tests -m gpu -q -rf --timeout 380 -k "multi_gpu or ipc_pipelined or fused_exchange or several_gpus" > $OUT/r02_pytest_multigpu_8gpu_box.log 2>&1
tail -5 $OUT/r02_pytest_multigpu_8gpu_box.log
for f in $OUT/r02_bench_*.err; do echo "== $f"; tail -c 300 $f; done
for f in $OUT/r02_bench_*.json; do echo "== $f"; cut -c1-250 $f; done
