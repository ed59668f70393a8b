#!/usr/bin/env python
"""Summarise profiles/<round>_*.ncu-rep (`ncu --set full` captures, tools/r2_profile_all.sh) into
profiles/<round>_ncu_full_summary.txt and profiles/<round>_traffic.json (the DRAM bytes per launch that bench.py reports as
roofline.traffic, keyed "<model>_n<batch>_<bench kernel name>").

Usage: python tools/ncu_summary.py [round-prefix]   (default r02)
"""
import csv
import io
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

# capture file stem (without the round prefix) -> (bench workload key, bench kernel name of each captured launch, in order)
CAPTURES = {
    "lstm_layer_fast_n512": ("fast_n512", ["lstm_layer"]),
    "step_fast_n512": ("fast_n512", ["conv12", "conv3_gemm", "linear_gemm", "crf_bwd_scan", "crf_fwd_beam", "crf_traceback"]),
    "lstm_rec_hac_n512": ("hac_n512", ["lstm_rec"]),
    "gx_gemm_hac_n512": ("hac_n512", ["lstm_gx_gemm"]),
    "decode_hac_n512": ("hac_n512", ["crf_bwd_scan", "crf_fwd_beam", "crf_traceback"]),
    "gemm_sup_layer0": ("sup_n128", ["qkv_gemm", "out_proj_gemm", "fc1_swiglu_gemm", "fc2_gemm"]),
    "tx_attention_sup_n128": ("sup_n128", ["tx_attention"]),
    "conv1_decode_sup_n128": ("sup_n128", ["tx_conv1", "crf_bwd_scan", "crf_fwd_beam", "crf_traceback"]),
}


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = [f"# {prefix}: `ncu --set full --clock-control none --import-source on` captures of the final kernels, one launch each, "
           f"taken by tools/r2_profile_all.sh from `python bench.py --steps 1 --warmup 1`; regenerate with tools/ncu_summary.py.",
           "# Durations here are cold-cache, serialised replays: the bench line's live CUDA-event times are the ones that count.", ""]
    traffic = {}
    # a capture is either the .ncu-rep itself or its `ncu -i ... --page raw --csv` export made on the GPU box (the box returns
    # at most 64 MiB per call, so not every .ncu-rep travels; tools/r2_profile_all.sh)
    prof = ROOT / "profiles"
    stems = sorted({f.name[:-len(".ncu-rep")] for f in prof.glob(f"{prefix}_*.ncu-rep")} |
                   {f.name[:-len(".raw.csv")] for f in prof.glob(f"{prefix}_*.raw.csv")})
    for full_stem in stems:
        csv_file, rep = prof / f"{full_stem}.raw.csv", prof / f"{full_stem}.ncu-rep"
        if csv_file.exists():
            raw = csv_file.read_text()
            rep = csv_file
        else:
            raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            out.append(f"== {rep.name}: unreadable")
            continue
        head, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(head)}
        stem = full_stem[len(prefix) + 1:]
        workload, names = CAPTURES.get(stem, (None, []))
        for li, vals in enumerate(rows[2:]):
            kname = vals[col["Kernel Name"]] if "Kernel Name" in col else "?"
            short = kname.split("(")[0].split("::")[-1].strip()
            label = names[li] if li < len(names) else "?"
            out.append(f"== {rep.name} launch {li}: {short}  [{label}]  grid {vals[col.get('Grid Size', 0)]} block {vals[col.get('Block Size', 0)]}")
            for k in KEEP:
                if k in col:
                    out.append(f"{k} [{units[col[k]]}] = {vals[col[k]]}")
            try:
                rd = float(vals[col["dram__bytes_read.sum"]]) * MULT[units[col["dram__bytes_read.sum"]]]
                wr = float(vals[col["dram__bytes_write.sum"]]) * MULT[units[col["dram__bytes_write.sum"]]]
                if workload and li < len(names):
                    traffic[f"{workload}_{names[li]}"] = {"file": rep.name, "launch": li, "kernel": short,
                                                         "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr}
            except (KeyError, ValueError):
                pass
            out.append("")
    (ROOT / "profiles" / f"{prefix}_ncu_full_summary.txt").write_text("\n".join(out))
    (ROOT / "profiles" / f"{prefix}_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    print("\n".join(out[:60]))


if __name__ == "__main__":
    main()
