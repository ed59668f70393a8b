#!/bin/bash
# round 2, battery 15: pipelined attention (guarded first), SFU activations in every epilogue -- full suite, sup / hac / fast lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "tx_model_scores" -p no:cacheprovider ) > gpurun_out/b15_attn_first.log 2>&1
echo "attn first rc=$?" >> gpurun_out/b15_attn_first.log
( time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b15_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b15_tests.log
timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b15_sup.json 2> gpurun_out/b15_bench.err
timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b15_hac.json 2>> gpurun_out/b15_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-sub-models > gpurun_out/b15_fast.json 2>> gpurun_out/b15_bench.err
echo done > gpurun_out/b15_done
