"""Developer driver for ncu: one population render (P candidates, 10 s stereo) through the mastering chain."""
import sys, os, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from st_ito_b200 import effects
from st_ito_b200.engine import Engine, compile_chain
from st_ito_b200.style_transfer import load_plugins
from tests.signals import test_signal

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with contextlib.redirect_stdout(io.StringIO()):
    plugins, D, _ = load_plugins(effects.make_chain("mastering-pb"))
desc, _ = compile_chain(plugins, 48000)
eng = Engine(model=None, device=0, chain=desc)
x = test_signal(2, 480000, seed=0)
W = np.random.RandomState(0).rand(P, D)
y = torch.empty((P, 2, 480000), dtype=torch.float32, device="cuda")
from st_ito_b200 import _lib
import time
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    _lib.check(_lib.lib().stito_process(eng._h, _lib.ptr(x), 2, 480000, _lib.ptr(W), P, D, 1, _lib.ptr(y), None))
    torch.cuda.synchronize(); print("render %d: %.2f ms" % (r, 1e3 * (time.time() - t0)))
print("peak", float(y.abs().max()))
