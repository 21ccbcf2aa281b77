#!/bin/bash
# Final round-2 lines with the shipped code: torchrun bench at N = $1 (the driver's command), plus the GPU suite when N = 2.
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  DXB_PARITY_LOG=$OUT/r02g_parity_metrics.jsonl timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rs > $OUT/r02g_pytest_gpu_2gpu.log 2>&1
  tail -4 $OUT/r02g_pytest_gpu_2gpu.log
fi
timeout 400 $TR --nproc-per-node $N --master-port 2957$N bench.py --gpus $N --steps 5 --warmup 3 > $OUT/r02g_bench_${N}gpu_pipelined.json 2> $OUT/r02g_bench_${N}gpu_pipelined.err
grep '^{' $OUT/r02g_bench_${N}gpu_pipelined.json | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print(j['n_gpus'], j['value'], j['ms_per_step'], j['exchange_ms']['non_kernel_ms_per_step'], j['e2e']['value'], j['e2e']['parts_ms']['set_grid'], j['e2e']['parts_ms']['transport'], j['e2e']['parts_ms']['get_dose'], j['clocks'])"
tail -c 400 $OUT/r02g_bench_${N}gpu_pipelined.err
