#!/bin/bash
# Re-measure of the one-process 8-GPU path after the per-device exchange enqueue (no cross-device events), torchrun line next to it.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 8 --inprocess --steps 5 --warmup 3 > $OUT/r02c_bench_8gpu_inprocess.json 2> $OUT/r02c_bench_8gpu_inprocess.err
timeout 300 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > $OUT/r02c_bench_8gpu_pipelined.json 2> $OUT/r02c_bench_8gpu_pipelined.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r02c_bench_1gpu.json 2> $OUT/r02c_bench_1gpu.err
timeout 300 python -m pytest tests -m gpu -q --timeout 280 -k "multi_gpu or several_gpus" 2>&1 | tail -3
for f in $OUT/r02c_bench_8gpu_inprocess.err $OUT/r02c_*.err; do echo "== $f"; tail -c 300 $f; done
for f in $OUT/r02c_bench_8gpu_inprocess.json $OUT/r02c_*.json; do echo "== $f"; cut -c1-250 $f; done
timeout 120 python profiles/d2h_probe.py > $OUT/r02c_d2h_probe_8gpu.json 2>&1; cat $OUT/r02c_d2h_probe_8gpu.json
