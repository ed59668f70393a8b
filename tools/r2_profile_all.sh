#!/bin/bash
# Round-2 ncu evidence for the FINAL kernels (run on the GPU box: gpurun -- bash tools/r2_profile_all.sh).
#   1. launch lists (gpu__time_duration.sum) of the bench command for fast / hac / sup -> gpurun_out/r02_launches_*.csv
#   2. one `--set full` capture per hot kernel -> gpurun_out/r02_*.ncu-rep (summarised by tools/ncu_summary.py into profiles/)
# A number printed by a run under ncu is never a bench value; these runs only feed profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-sub-models --steps 1 --warmup 1"
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FULL="ncu --set full --clock-control none --import-source on -f"
# ---- launch lists: one step of one runner (warm-up step + timed step + the three profiled passes follow; -c bounds the list)
$LL -c 60  --log-file gpurun_out/r02_launches_fast_n512.csv $B --runners 1 > gpurun_out/ncu_ll_fast.log 2>&1
$LL -c 70  --log-file gpurun_out/r02_launches_hac_n512.csv $B --model hac --batch 512 --runners 1 > gpurun_out/ncu_ll_hac.log 2>&1
$LL -c 260 --log-file gpurun_out/r02_launches_sup_n128.csv $B --model sup --batch 128 --runners 1 > gpurun_out/ncu_ll_sup.log 2>&1
# ---- fast: the recurrence in the shape the default bench runs (4 runners: 32 CTAs x 2 groups x 8 chunks), then the rest of a step
$FULL -k regex:lstm_layer -s 2 -c 1 -o gpurun_out/r02_lstm_layer_fast_n512 $B > gpurun_out/ncu_fast_lstm.log 2>&1
$FULL -k "regex:conv12|gemm_f16|crf_" -c 6 -o gpurun_out/r02_step_fast_n512 $B --runners 1 > gpurun_out/ncu_fast_step.log 2>&1
# ---- hac: cluster recurrence (default: 2 runners -> 64 chunks per cluster), gx GEMM + linear + decode
$FULL -k regex:lstm_cluster -s 1 -c 1 -o gpurun_out/r02_lstm_rec_hac_n512 $B --model hac --batch 512 > gpurun_out/ncu_hac_rec.log 2>&1
$FULL -k "regex:gemm_f16|crf_" -s 1 -c 1 -o gpurun_out/r02_gx_gemm_hac_n512 $B --model hac --batch 512 --runners 1 > gpurun_out/ncu_hac_gx.log 2>&1
$FULL -k "regex:crf_" -c 3 -o gpurun_out/r02_decode_hac_n512 $B --model hac --batch 512 --runners 1 > gpurun_out/ncu_hac_dec.log 2>&1
# ---- sup: layer 0 GEMMs (QKV+RoPE, out-proj, FC1+SwiGLU, FC2), attention, first conv, CRF GEMM + decode
S="$B --model sup --batch 128 --runners 1"
$FULL -k regex:gemm_f16 -s 4 -c 4 -o gpurun_out/r02_gemm_sup_layer0 $S > gpurun_out/ncu_sup_gemm.log 2>&1
$FULL -k regex:tx_attention -s 1 -c 1 -o gpurun_out/r02_tx_attention_sup_n128 $S > gpurun_out/ncu_sup_attn.log 2>&1
$FULL -k "regex:tx_conv1|crf_" -c 4 -o gpurun_out/r02_conv1_decode_sup_n128 $S > gpurun_out/ncu_sup_dec.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
