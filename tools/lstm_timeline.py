import os, sys, numpy as np
os.environ["B200_DEBUG_LSTM_TIMELINE"] = "1"
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import model_dir
from dorado_b200.config import load_model_config
from dorado_b200.runner import B200Caller, B200ModelRunner
from dorado_b200.weights import synthetic_weights
cfg = load_model_config(model_dir("fast")); w = synthetic_weights(cfg, 42)
caller = B200Caller(cfg, w); runner = B200ModelRunner(caller, 512, 9996)
runner.input_view()[:] = np.random.default_rng(0).standard_normal((512, 9996)).astype(np.float16)
runner.upload(); runner.step_device(512, 2)
