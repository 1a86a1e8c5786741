"""Developer probe: wall-clock per generation vs device time (host-side overhead between generations)."""
import os, sys, time, contextlib, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from st_ito_b200 import effects
from st_ito_b200.engine import compile_chain
from st_ito_b200.style_transfer import load_plugins, process_audio
from st_ito_b200.utils import make_synthetic_param_model
x = bench.make_workload(10.0)
with contextlib.redirect_stdout(io.StringIO()):
    plugins, D, _ = load_plugins(effects.make_chain("mastering-pb"))
model = make_synthetic_param_model(seed=3)
eng = model.stito_engine(0)
desc, _ = compile_chain(plugins, 48000); eng.set_chain(desc)
eng.set_target(process_audio(x, np.random.RandomState(1234).rand(D), 48000, plugins)); eng.set_input(x)
Ws = [np.random.RandomState(i).rand(64, D) for i in range(12)]
for W in Ws[:3]: eng.eval_population(W, 0, 480000)
torch.cuda.synchronize()
t0 = time.perf_counter(); dev = 0.0
for W in Ws[3:]:
    eng.eval_population(W, 0, 480000)
    dev += eng.timing()["ms_total"]
wall = (time.perf_counter() - t0) * 1e3
print(f"wall {wall/9:.3f} ms/gen, device ms_total {dev/9:.3f} ms/gen, overhead {(wall-dev)/9:.3f} ms/gen")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for W in Ws[3:]: eng.eval_population(W, 0, 480000)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
