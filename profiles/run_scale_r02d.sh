#!/bin/bash
# 8-GPU box: the one-process path (resident per-device host threads, exchange enqueued at dxb_finish_beam) and the torchrun
# path, alternating on the same box (box-to-box kernel-time differences are ~2 %, larger than the effect measured).
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
DXB_TRACE_HOST=1 timeout 300 python bench.py --gpus 8 --inprocess --steps 5 --warmup 3 > $OUT/r02d_bench_8gpu_inprocess.json 2> $OUT/r02d_bench_8gpu_inprocess.err
timeout 300 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r02d_bench_8gpu_pipelined.json 2> $OUT/r02d_bench_8gpu_pipelined.err
timeout 300 python bench.py --gpus 8 --inprocess --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r02d_bench_8gpu_inprocess_b.json 2> $OUT/r02d_bench_8gpu_inprocess_b.err
timeout 300 $TR --nproc-per-node 8 --master-port 29562 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r02d_bench_8gpu_pipelined_b.json 2> $OUT/r02d_bench_8gpu_pipelined_b.err
timeout 300 python -m pytest tests -m gpu -q --timeout 280 -k "multi_gpu or several_gpus" 2>&1 | tail -3
grep "dxb host" $OUT/r02d_bench_8gpu_inprocess.err | tail -8
for f in $OUT/r02d_bench_8gpu_*.json; do echo "== $f"; grep '^{' $f | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print(j['ms_per_step'], j.get('exchange_ms',{}).get('non_kernel_ms_per_step'), j['clocks']['sm_mhz'], j['clocks']['power_w_max'], (j.get('e2e') or {}).get('value'))"; done
