"""Generate tests/golden/frontend.npz from the reference's own front-end code (oracle/_ref/libdorado_ref.so built from
/root/reference: chunk.cpp, stitch.cpp, tensor_utils.cpp + BasecallerNode's slice/repeat-pad calls in ref_driver.cpp).

Run where /root/reference exists:  python tools/make_golden_frontend.py
A few synthetic reads (short, exactly one chunk, ragged tail, long) -> chunk offsets and the fp16 model-input rows the
reference would feed its runner.  The GPU box has no /root/reference; tests compare the engine with these rows.
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.oracle import Reference  # noqa: E402

CHUNK, STRIDE, OVERLAP = 1200, 6, 120
READ_LENS = [7, 500, 1200, 1201, 1733, 2400, 5003, 9000]


def main():
    ref = Reference()
    rng = np.random.default_rng(77)
    out = {"chunk_size": CHUNK, "stride": STRIDE, "overlap": OVERLAP, "read_lens": np.array(READ_LENS)}
    rows, meta = [], []
    for r, n in enumerate(READ_LENS):
        raw = np.clip(rng.standard_normal(n) * 180 + 420, -32768, 32767).astype(np.int16)
        shift, scale = np.float32(400 + 13.7 * r), np.float32(150 + 3.3 * r)
        offs = ref.generate_chunks(n, CHUNK, STRIDE, OVERLAP)
        out[f"raw_{r}"] = raw
        out[f"offsets_{r}"] = np.array(offs, np.uint64)
        out[f"shift_scale_{r}"] = np.array([shift, scale], np.float32)
        for o in offs:
            rows.append(ref.make_chunk_input(raw, o, CHUNK, float(shift), float(scale)).view(np.uint16))
            meta.append((r, o))
    out["input_rows_f16_bits"] = np.stack(rows)
    out["row_read_offset"] = np.array(meta, np.uint64)
    np.savez_compressed(ROOT / "tests" / "golden" / "frontend.npz", **out)
    print("reads", READ_LENS, "chunks", len(rows))


if __name__ == "__main__":
    main()
