#!/bin/bash
# compute-sanitizer over the device code paths (one GPU): memcheck (out-of-bounds / misaligned accesses, leaks of device
# memory reported at exit), initcheck (reads of uninitialised global memory) and synccheck (invalid barrier / warp-sync use).
# racecheck is not run: the pool kernel hands slots between warps through shared-memory status words with fences and atomics
# by design, which racecheck - a detector for barrier-separated accesses - cannot model.
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python profiles/sanitizer_target.py 60000 > $OUT/r02_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|LEAK SUMMARY|C1 mode|C2/8|C3/4|postprocessed|Error|error" $OUT/r02_sanitizer_$tool.log | head -20
done
