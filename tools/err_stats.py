import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import model_dir
from dorado_b200.config import load_model_config
from dorado_b200.runner import B200Caller, B200ModelRunner
from dorado_b200.weights import synthetic_weights
from oracle import nn_oracle
kind, N, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = load_model_config(model_dir(kind)); w = synthetic_weights(cfg, 42)
caller = B200Caller(cfg, w); runner = B200ModelRunner(caller, N, T)
sig = np.random.default_rng(1234).standard_normal((N, runner.chunk_size())).astype(np.float16)
for i in range(N): runner.accept_chunk(i, sig[i])
got = runner.forward_scores(N).astype(np.float32)
if cfg.clamp: got = np.clip(got, -5, 5)
print("got finite", np.isfinite(got).all(), "std", got.std())
for name, ref in [("fp32 oracle", nn_oracle.forward(cfg, w, sig.astype(np.float32))), ("fp16-storage oracle", nn_oracle.forward(cfg, w, sig.astype(np.float32), emulate_fp16=True))]:
    e = np.abs(got - ref); sc = np.abs(ref).max()
    print(f"{kind} vs {name}: scale {sc:.2f} max {e.max():.4f} mean {e.mean():.5f} relL2 {np.linalg.norm(got-ref)/np.linalg.norm(ref):.2e} frac>1e-3*scale {(e>1e-3*sc).mean():.4f} frac>4e-3*scale {(e>4e-3*sc).mean():.5f} p99.9 {np.quantile(e,0.999):.4f}")
    print("   per-t mean err (every 16th)", np.round(e.mean(axis=(0,2))[::16], 4))
