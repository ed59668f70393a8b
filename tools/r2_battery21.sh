#!/bin/bash
# round 2, battery 21: evidence run of the final tree -- full GPU suite, the default bench line (CPU reference arm and the hac / sup
# sub-results included), the reference arm alone, then the ncu launch lists and captures of tools/r2_profile_all.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b21_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b21_tests.log
( time timeout 900 python bench.py ) > gpurun_out/b21_bench_default.json 2> gpurun_out/b21_bench_default.err
timeout 1500 bash tools/r2_profile_all.sh > gpurun_out/b21_profile.log 2>&1
echo done > gpurun_out/b21_done
