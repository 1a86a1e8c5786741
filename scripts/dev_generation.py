"""Developer driver for ncu: `reps` population evaluations (P candidates, `seconds` of stereo audio, given chain)
bracketed by cudaProfilerStart/Stop, so that `ncu --profile-from-start off` sees exactly those generations.

    python scripts/dev_generation.py [P] [reps] [seconds] [chain]
"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from st_ito_b200 import effects
from st_ito_b200.engine import compile_chain
from st_ito_b200.style_transfer import load_plugins
from st_ito_b200.utils import make_synthetic_param_model
from tests.signals import test_signal

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
chain = sys.argv[4] if len(sys.argv) > 4 else "mastering-pb"
L = int(seconds * 48000)
with contextlib.redirect_stdout(io.StringIO()):
    plugins, D, _ = load_plugins(effects.make_chain(chain))
model = make_synthetic_param_model(seed=3, conv_gain=2.0)
eng = model.stito_engine(0)
desc, _ = compile_chain(plugins, 48000)
eng.set_chain(desc)
x = test_signal(2, L, seed=0)
x = x / np.abs(x).max()
eng.set_input(x)
eng.set_target(x)
W = np.random.RandomState(0).rand(P, D)
for _ in range(2):  # warm-up: calibration pass + steady state
    eng.eval_population(W, 0, L)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for r in range(reps):
    fit, _, _ = eng.eval_population(W, 0, L)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
t = eng.timing()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items() if k != "ms_conv"}, "fit[0..3]", fit[:4].tolist())
