#!/bin/bash
# Round-2 ncu evidence for the FINAL kernels (run on the GPU box: gpurun -- bash tools/r2_profile_all.sh).
#   1. launch lists (gpu__time_duration.sum) of the bench command for fast / hac / sup -> gpurun_out/r02_launches_*.csv
#   2. one `--set full` capture per hot kernel.  gpurun pulls at most 64 MiB back, and a capture with sources is 2-10 MB per
#      launch, so the .ncu-rep files are written to /tmp on the box; what comes back for every capture is its raw-metrics CSV
#      and its per-SASS-line source page (`ncu -i ... --page raw|source --csv`, the same views one would open from the
#      file), and the .ncu-rep files themselves in the priority order below until 40 MiB are used.
# A number printed by a run under ncu is never a bench value; these runs only feed profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
B="python bench.py --no-cpu-baseline --no-sub-models --steps 1 --warmup 1"
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FULL="ncu --set full --clock-control none --import-source on -f"
# ---- launch lists: one step of one runner (warm-up step + timed step + the three profiled passes follow; -c bounds the list)
$LL -c 60  --log-file gpurun_out/r02_launches_fast_n512.csv $B --runners 1 > gpurun_out/ncu_ll_fast.log 2>&1
$LL -c 70  --log-file gpurun_out/r02_launches_hac_n512.csv $B --model hac --batch 512 --runners 1 > gpurun_out/ncu_ll_hac.log 2>&1
$LL -c 260 --log-file gpurun_out/r02_launches_sup_n128.csv $B --model sup --batch 128 --runners 1 > gpurun_out/ncu_ll_sup.log 2>&1
# ---- fast: the recurrence in the shape the default bench runs (4 runners: 32 CTAs x 2 groups x 8 chunks), then the rest of a step
$FULL -k regex:lstm_layer -s 2 -c 1 -o /tmp/ncu/r02_lstm_layer_fast_n512 $B > gpurun_out/ncu_fast_lstm.log 2>&1
$FULL -k "regex:conv12|gemm_f16|crf_" -c 6 -o /tmp/ncu/r02_step_fast_n512 $B --runners 1 > gpurun_out/ncu_fast_step.log 2>&1
# ---- hac: cluster recurrence (default: 4 runners -> 64 chunks per cluster), gx GEMM, decode
$FULL -k regex:lstm_cluster -s 1 -c 1 -o /tmp/ncu/r02_lstm_rec_hac_n512 $B --model hac --batch 512 > gpurun_out/ncu_hac_rec.log 2>&1
$FULL -k "regex:gemm_f16" -s 1 -c 1 -o /tmp/ncu/r02_gx_gemm_hac_n512 $B --model hac --batch 512 --runners 1 > gpurun_out/ncu_hac_gx.log 2>&1
$FULL -k "regex:crf_" -c 3 -o /tmp/ncu/r02_decode_hac_n512 $B --model hac --batch 512 --runners 1 > gpurun_out/ncu_hac_dec.log 2>&1
# ---- sup: layer 0 GEMMs (QKV+RoPE, out-proj, FC1+SwiGLU, FC2), attention, first conv, decode
S="$B --model sup --batch 128 --runners 1"
$FULL -k regex:gemm_f16 -s 4 -c 4 -o /tmp/ncu/r02_gemm_sup_layer0 $S > gpurun_out/ncu_sup_gemm.log 2>&1
$FULL -k regex:tx_attention -s 1 -c 1 -o /tmp/ncu/r02_tx_attention_sup_n128 $S > gpurun_out/ncu_sup_attn.log 2>&1
$FULL -k "regex:tx_conv1|crf_" -c 4 -o /tmp/ncu/r02_conv1_decode_sup_n128 $S > gpurun_out/ncu_sup_dec.log 2>&1
# ---- exports
for rep in /tmp/ncu/*.ncu-rep; do
  stem=$(basename "$rep" .ncu-rep)
  ncu -i "$rep" --page raw --csv > gpurun_out/${stem}.raw.csv 2>/dev/null
  ncu -i "$rep" --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${stem}.source.csv.gz
done
used=0
for stem in r02_step_fast_n512 r02_lstm_layer_fast_n512 r02_lstm_rec_hac_n512 r02_tx_attention_sup_n128 r02_gx_gemm_hac_n512 \
            r02_decode_hac_n512 r02_gemm_sup_layer0 r02_conv1_decode_sup_n128; do
  f=/tmp/ncu/$stem.ncu-rep
  [ -f "$f" ] || continue
  sz=$(stat -c %s "$f")
  if [ $((used + sz)) -le $((40 * 1024 * 1024)) ]; then cp "$f" gpurun_out/; used=$((used + sz)); fi
done
ls -la /tmp/ncu gpurun_out/*.ncu-rep; du -sh gpurun_out
