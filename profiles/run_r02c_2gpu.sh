#!/bin/bash
# 2-GPU box: the whole GPU suite (all multi-GPU tests run), then the one-process and the torchrun bench lines at N=2.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
DXB_PARITY_LOG=$OUT/r02c_parity_metrics.jsonl timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rs > $OUT/r02c_pytest_gpu_2gpu.log 2>&1
tail -8 $OUT/r02c_pytest_gpu_2gpu.log
timeout 300 python bench.py --gpus 2 --inprocess --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02c_bench_2gpu_inprocess.json 2> $OUT/r02c_bench_2gpu_inprocess.err
timeout 300 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > $OUT/r02c_bench_2gpu_pipelined.json 2> $OUT/r02c_bench_2gpu_pipelined.err
for f in $OUT/r02c_bench_2gpu_*.err; do echo "== $f"; tail -c 300 $f; done
for f in $OUT/r02c_bench_2gpu_*.json; do echo "== $f"; cut -c1-300 $f; done
