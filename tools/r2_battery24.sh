#!/bin/bash
# round 2, battery 24: evidence refresh after the GEMM tile-order change -- full GPU suite and default bench line of the final tree,
# re-capture of the GEMM launches (hac x-projection, sup layer 0) and the hac launch list; the other captures of battery 21 stand
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
( time timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b24_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b24_tests.log
( time timeout 900 python bench.py ) > gpurun_out/b24_bench_default.json 2> gpurun_out/b24_bench_default.err
B="python bench.py --no-cpu-baseline --no-sub-models --steps 1 --warmup 1"
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FULL="ncu --set full --clock-control none --import-source on -f"
$LL -c 70 --log-file gpurun_out/r02_launches_hac_n512.csv $B --model hac --batch 512 --runners 1 > gpurun_out/ncu_ll_hac.log 2>&1
$FULL -k "regex:gemm_f16" -s 1 -c 1 -o /tmp/ncu/r02_gx_gemm_hac_n512 $B --model hac --batch 512 --runners 1 > gpurun_out/ncu_hac_gx.log 2>&1
$FULL -k regex:gemm_f16 -s 4 -c 4 -o /tmp/ncu/r02_gemm_sup_layer0 $B --model sup --batch 128 --runners 1 > gpurun_out/ncu_sup_gemm.log 2>&1
for rep in /tmp/ncu/*.ncu-rep; do
  stem=$(basename "$rep" .ncu-rep)
  ncu -i "$rep" --page raw --csv > gpurun_out/${stem}.raw.csv 2>/dev/null
  ncu -i "$rep" --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${stem}.source.csv.gz
done
cp /tmp/ncu/r02_gx_gemm_hac_n512.ncu-rep gpurun_out/
du -sh gpurun_out
echo done > gpurun_out/b24_done
