"""cuobjdump -sass of the built library, summarised per kernel for profiles/ (no GPU needed).

usage: python tools/sass_extract.py [lib.so] [out_prefix]
Writes <out_prefix>_summary.txt (per kernel: instruction count, registers from the ELF, and the counts of the
mnemonics that prove which pipes a kernel uses: UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG =
TMA load/store, UBLKCP = bulk async copy, HMMA = legacy mma.sync, MUFU, REDUX, BAR) and <out_prefix>_<kernel>.txt with the
full listing of the kernels named in FULL below.
"""
import collections, re, subprocess, sys, pathlib

lib = sys.argv[1] if len(sys.argv) > 1 else "dorado_b200/libb200call.so"
prefix = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_sass"
FULL = ["crf_fwd_beam_kernel<3>", "crf_bwd_scan_kernel<3>", "crf_traceback_kernel", "tx_attention_tc_kernel",
        "lstm_layer_kernel<96, 8, 2>", "conv12_tc_kernel", "lstm_cluster_kernel<384, 6, 2, 32>"]
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "IMMA", "MUFU",
        "REDUX", "MATCH", "SHFL", "BAR", "MEMBAR", "LDS", "STS", "LDG", "STG", "FFMA", "FADD", "FMUL", "DFMA"]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur, name = [], None, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = demangle(m.group(1))
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"^b200::", "", name)
        name = re.sub(r"\(.*$", "", name)
        cur = []
        kernels.append((name, cur))
    elif cur is not None:
        cur.append(line)

def mnemonic(line):
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    return m.group(1) if m else None

with open(prefix + "_summary.txt", "w") as out:
    out.write(f"# cuobjdump -sass {lib} -- per-kernel mnemonic counts (tools/sass_extract.py)\n")
    out.write("# kernel | instructions | " + " ".join(KEYS) + "\n")
    for name, lines in kernels:
        ops = [mnemonic(l) for l in lines]
        ops = [o for o in ops if o]
        cnt = collections.Counter(o.split(".")[0] for o in ops)
        out.write(f"{name} | {len(ops)} | " + " ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]) + "\n")
    tot = collections.Counter()
    for _, lines in kernels:
        for l in lines:
            o = mnemonic(l)
            if o:
                tot[o.split(".")[0]] += 1
    out.write("# whole library: " + " ".join(f"{k}={tot[k]}" for k in KEYS) + "\n")
for name, lines in kernels:
    if name in FULL:
        fn = re.sub(r"[^A-Za-z0-9]+", "_", name).strip("_")
        with open(f"{prefix}_{fn}.txt", "w") as f:
            f.write(f"# {name}\n")
            # keep the instruction column only (drop the encoding words)
            for l in lines:
                m = re.match(r"(\s+/\*[0-9a-f]{4}\*/\s+.*?;)", l)
                if m:
                    f.write(m.group(1).rstrip() + "\n")
                elif l.strip().startswith(".L_"):
                    f.write(l.rstrip() + "\n")
print(open(prefix + "_summary.txt").read())
