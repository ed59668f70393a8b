"""In-kernel clock64 timeline of the LSTM recurrence (CTA 0, steps 64..67), printed by the plan on stderr.
usage: python tools/lstm_timeline.py [fast|hac] [batch]   (B200_LSTM_V1=1 / B200_CLUSTER_V1=1 select the first-generation kernels)"""
import os, sys, numpy as np
os.environ["B200_DEBUG_LSTM_TIMELINE"] = "1"
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import model_dir
from dorado_b200.config import load_model_config
from dorado_b200.runner import B200Caller, B200ModelRunner
from dorado_b200.weights import synthetic_weights
kind = sys.argv[1] if len(sys.argv) > 1 else "fast"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
cfg = load_model_config(model_dir(kind)); w = synthetic_weights(cfg, 42)
caller = B200Caller(cfg, w); runner = B200ModelRunner(caller, N, 9996)
runner.input_view()[:] = np.random.default_rng(0).standard_normal((N, 9996)).astype(np.float16)
runner.upload(); runner.step_device(N, 2)
