#!/bin/bash
# Round-2 ncu evidence, one GPU (gpurun): (1) launch list of bench.py, (2) --set full capture of the shipped transport kernel on
# C2 (3e7 histories of the 512x512x300 beam), (3) the slab-local-majorant build on C3.  Reports come back in gpurun_out/ and are
# summarised here with profiles/ncu_summary.py / ncu_buckets.py; numbers printed under ncu are never bench values.
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:transportKernelPool -s 1 -c 1 -o $OUT/r02_pool -f \
    python profiles/ncu_target.py c2 > $OUT/r02_pool_target.json 2> $OUT/r02_pool_target.err
ncu --set full --clock-control none --import-source on -k regex:transportKernelPool -s 1 -c 1 -o $OUT/r02_pool_lm -f \
    python profiles/ncu_target.py c3 > $OUT/r02_pool_lm_target.json 2> $OUT/r02_pool_lm_target.err
tail -n 2 $OUT/r02_pool_target.json; tail -n 2 $OUT/r02_pool_lm_target.json
ls -la $OUT/*.ncu-rep
