#!/bin/bash
# Final round-2 lines with the shipped code on an N-GPU box ($1): the GPU suite (all multi-GPU tests run at N >= 2), the 1-GPU
# bench line, the one-process and the torchrun (the driver's command) lines at N.
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { grep '^{' $1 | python -c "import sys,json; j=json.loads(sys.stdin.readline()); e=j.get('e2e') or {}; print(j['n_gpus'], 'value %.4g' % j['value'], 'ms/step %.2f' % j['ms_per_step'], 'non-kernel', (j.get('exchange_ms') or {}).get('non_kernel_ms_per_step'), 'e2e', e.get('value'), (e.get('parts_ms') or {}).get('set_grid'), (e.get('parts_ms') or {}).get('transport'), (e.get('parts_ms') or {}).get('get_dose'), 'traffic', j['roofline']['traffic_bytes_per_history'], 'frac %.4f sector %.4f' % (j['roofline']['frac'], j['roofline']['sector_frac']), j['clocks']['sm_mhz'], j['clocks']['reasons'])"; }
if [ "$N" = "2" ]; then
  DXB_PARITY_LOG=$OUT/r02k_parity_metrics.jsonl timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rs > $OUT/r02k_pytest_gpu_2gpu.log 2>&1
  tail -4 $OUT/r02k_pytest_gpu_2gpu.log
fi
timeout 400 python bench.py --steps 5 --warmup 3 > $OUT/r02k_bench_1gpu_on_${N}gpu_box.json 2> $OUT/r02k_bench_1gpu.err; show $OUT/r02k_bench_1gpu_on_${N}gpu_box.json
timeout 400 python bench.py --gpus $N --inprocess --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02k_bench_${N}gpu_inprocess.json 2> $OUT/r02k_bench_${N}gpu_inprocess.err; show $OUT/r02k_bench_${N}gpu_inprocess.json
timeout 400 $TR --nproc-per-node $N --master-port 2958$N bench.py --gpus $N --steps 5 --warmup 3 > $OUT/r02k_bench_${N}gpu_pipelined.json 2> $OUT/r02k_bench_${N}gpu_pipelined.err; show $OUT/r02k_bench_${N}gpu_pipelined.json
tail -c 300 $OUT/r02k_bench_${N}gpu_inprocess.err
