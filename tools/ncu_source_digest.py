#!/usr/bin/env python
"""Digest of the per-SASS-line source pages exported by tools/r2_profile_all.sh (profiles/<round>_*.source.csv.gz):
for every captured launch the executed warp-instructions, the stall samples, and the SASS lines that collected the most
samples.  DESIGN.md quotes the executed-instruction counts from here.

Usage: python tools/ncu_source_digest.py [round-prefix]   ->  profiles/<round>_ncu_source_digest.txt
"""
import csv
import gzip
import io
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
TOP = 10


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = [f"# {prefix}: per-launch digest of the ncu source pages (SASS view): executed warp-instructions, stall samples and the "
           f"{TOP} lines with the most samples (samples | executions | instruction).  Regenerate with tools/ncu_source_digest.py."]
    for f in sorted((ROOT / "profiles").glob(f"{prefix}_*.source.csv.gz")):
        rows = list(csv.reader(io.StringIO(gzip.open(f, "rt").read())))
        starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
        seen = {}
        for k, st in enumerate(starts):
            end = starts[k + 1] if k + 1 < len(starts) else len(rows)
            hdr = rows[st + 1] if st + 1 < len(rows) else []
            if "Instructions Executed" not in hdr or "# Samples" not in hdr:
                continue
            name = rows[st][1].replace("(int)", "").split("(")[0].replace("b200::<unnamed>::", "").replace("void ", "").strip()
            launch = seen.get(name, 0)
            seen[name] = launch + 1
            ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
            sec = [r for r in rows[st + 2:end] if len(r) > max(ci, cs) and r[ci].isdigit()]
            if not sec:
                continue
            # every launch appears twice (two views of the same SASS); keep the first
            if launch % 2 == 1:
                continue
            tot = sum(int(r[ci]) for r in sec)
            samp = sum(int(r[cs]) for r in sec)
            out.append("")
            out.append(f"== {f.name}: {name}  (capture {launch // 2})  SASS lines {len(sec)}, executed warp-instructions {tot:,}, "
                       f"stall samples {samp:,}")
            for r in sorted(sec, key=lambda r: -int(r[cs]))[:TOP]:
                out.append(f"   {int(r[cs]):>8,} | {int(r[ci]):>12,} | {r[1].strip()[:110]}")
    dst = ROOT / "profiles" / f"{prefix}_ncu_source_digest.txt"
    dst.write_text("\n".join(out) + "\n")
    print(f"{dst}: {len(out)} lines")


if __name__ == "__main__":
    main()
