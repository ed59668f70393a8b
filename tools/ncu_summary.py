#!/usr/bin/env python
"""Summarise every profiles/*.ncu-rep (one `ncu --set full` capture per kernel) into profiles/r01_ncu_full_summary.txt.

Usage: python tools/ncu_summary.py [round-prefix]   (default r01)
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
        "launch__cluster_size", "launch__cluster_max_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
# which kernel version each capture shows (commit of the capture; later commits that changed the kernel)
NOTES = {
    "lstm_layer_fast": "captured at 25d6e20, before 6298545 / 3649b83 (shared-space pointers, weights in tensor memory): "
                       "that version took 1.68 ms/launch under ncu, the final kernel 1.02 ms (r01_launches_fast_n512.csv); "
                       "no GPU budget was left to re-capture",
    "crf_fwd_beam_fast": "captured at 25d6e20 (decode v2); the final kernel (35b17a9: redux max, deferred pow) takes 2.80 ms "
                         "(r01_launches_fast_n512.csv)",
    "crf_bwd_scan_fast": "captured at 25d6e20 together with the forward kernel; the final backward scan takes 0.89 ms (r01_launches_fast_n512.csv)",
    "lstm_cluster_hac_n512_tmem": "final kernel (43cb863): W_hh in tensor memory, 32 chunks per cluster, 16 clusters",
}


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r01"
    out = [f"# {prefix}: one `ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 python bench.py "
           f"--steps 1 --warmup 1` capture per kernel; regenerate with tools/ncu_summary.py", ""]
    traffic = {}  # bench.py's roofline.traffic: DRAM bytes per launch of each captured kernel, keyed by capture name
    for rep in sorted((ROOT / "profiles").glob(f"{prefix}_*.ncu-rep")):
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            out.append(f"== {rep.name}: unreadable")
            continue
        head, units, vals = rows[0], rows[1], rows[2]
        col = {h: i for i, h in enumerate(head)}
        name = vals[col["Kernel Name"]] if "Kernel Name" in col else "?"
        grid = vals[col["Grid Size"]] if "Grid Size" in col else ""
        block = vals[col["Block Size"]] if "Block Size" in col else ""
        out.append(f"== {rep.name}: {name}  grid {grid} block {block}")
        for key, note in NOTES.items():
            if key in rep.name:
                out.append(f"   note: {note}")
        for k in KEEP:
            if k in col:
                out.append(f"{k} [{units[col[k]]}] = {vals[col[k]]}")
        try:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(vals[col["dram__bytes_read.sum"]]) * mult[units[col["dram__bytes_read.sum"]]]
            wr = float(vals[col["dram__bytes_write.sum"]]) * mult[units[col["dram__bytes_write.sum"]]]
            traffic[rep.stem] = {"kernel": name.split("(")[0].split("::")[-1].strip(), "dram_read_bytes": rd,
                                 "dram_write_bytes": wr, "dram_bytes": rd + wr}
        except (KeyError, ValueError):
            pass
        out.append("")
    (ROOT / "profiles" / f"{prefix}_ncu_full_summary.txt").write_text("\n".join(out))
    (ROOT / "profiles" / f"{prefix}_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
