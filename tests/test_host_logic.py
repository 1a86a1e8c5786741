"""CPU tests (no GPU, no compute calls into libstito): the C-ABI library loads and exports every symbol the
header declares with matching struct layouts, the host-side chain compilation / CMA-ES / sharding logic, and the
N > 1 host loop under a world_size-2 gloo group (the CUDA engine is replaced by a deterministic fake)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------- C ABI
def _header_functions():
    src = open(os.path.join(ROOT, "include", "stito.h")).read()
    return sorted(set(re.findall(r"STITO_API\s+[\w\s\*]+?\b(stito_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from st_ito_b200 import _lib

    _lib.build()
    names = _header_functions()
    assert len(names) >= 15 and "stito_eval_population" in names
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"libstito.so does not export {n}"
    assert sorted(_lib.EXPORTS) == names, "st_ito_b200/_lib.py binds a different symbol set than include/stito.h"
    # nothing but the ABI is visible (the kernels are built with -fvisibility=hidden)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(e for e in exported if e.startswith("stito_")) == names
    assert L.stito_version() >= 100
    L.stito_last_error.restype = ctypes.c_char_p
    assert isinstance(L.stito_last_error(), bytes)


def test_ctypes_struct_layouts_match_the_header():
    """sizeof/offsetof of the ctypes mirrors against a C program compiled from include/stito.h."""
    from st_ito_b200 import _lib

    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "stito.h"
int main(void) {
    printf("%zu %zu %zu %zu ", sizeof(stito_fx_desc), sizeof(stito_chain_desc), sizeof(stito_encoder_weights), sizeof(stito_timing));
    printf("%zu %zu %zu ", offsetof(stito_fx_desc, fixed_raw), offsetof(stito_chain_desc, fx), offsetof(stito_chain_desc, sample_rate));
    printf("%zu %zu %zu\n", offsetof(stito_encoder_weights, conv_w), offsetof(stito_encoder_weights, mel_w), offsetof(stito_timing, ms_conv));
    return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "layout.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "layout")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    want = [ctypes.sizeof(_lib.FxDesc), ctypes.sizeof(_lib.ChainDesc), ctypes.sizeof(_lib.EncoderWeights),
            ctypes.sizeof(_lib.Timing), _lib.FxDesc.fixed_raw.offset, _lib.ChainDesc.fx.offset,
            _lib.ChainDesc.sample_rate.offset, _lib.EncoderWeights.conv_w.offset, _lib.EncoderWeights.mel_w.offset,
            _lib.Timing.ms_conv.offset]
    assert got == want


def test_product_path_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA the host mirror raises instead of computing elsewhere."""
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from st_ito_b200 import effects
    from st_ito_b200.utils import make_synthetic_param_model

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        effects.BasicParametricEQ().process(np.zeros((1, 64), dtype=np.float32), 48000)
    m = make_synthetic_param_model(seed=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 2, 48000))


# -------------------------------------------------------------------------------- chain compilation
def _chain(kinds=("eq", "comp", "reverb"), fixed=None):
    from st_ito_b200 import effects

    table = {"eq": ("ParametricEQ", effects.BasicParametricEQ, 1), "comp": ("Compressor", effects.BasicCompressor, 1),
             "dist": ("Distortion", effects.BasicDistortion, 1), "delay": ("Delay", effects.BasicDelay, 2),
             "reverb": ("Reverb", effects.BasicReverb, 2)}
    plugins = {}
    for k in kinds:
        name, cls, ch = table[k]
        plugins[name] = {"class_path": cls, "num_params": None, "num_channels": ch,
                         "fixed_parameters": dict((fixed or {}).get(k, {}))}
    return plugins


def test_load_plugins_and_compile_chain_index_bookkeeping(capsys):
    from st_ito_b200 import _lib
    from st_ito_b200.engine import compile_chain
    from st_ito_b200.style_transfer import load_plugins, parameters_to_dict

    plugins, D, init = load_plugins(_chain(fixed={"comp": {"ratio": 4.0}}))
    capsys.readouterr()
    assert D == 29 == len(init)  # (1+18) + (1+4) + (1+4): one dead our_bypass slot per plugin
    assert [p["parameter_names"][0] for p in plugins.values()] == ["our_bypass"] * 3
    desc, Dc = compile_chain(plugins, 48000)
    assert Dc == D and desc.num_fx == 3 and desc.num_w == 29
    eq, comp, rev = desc.fx[0], desc.fx[1], desc.fx[2]
    assert (eq.kind, comp.kind, rev.kind) == (_lib.FX_EQ, _lib.FX_COMPRESSOR, _lib.FX_REVERB)
    assert list(eq.w_index[:18]) == list(range(1, 19))  # slot 0 is our_bypass
    # the fixed ratio still consumes its w slot (style_transfer.py:80-85) but maps to the fixed raw value
    assert list(comp.w_index[:4]) == [20, -1, 22, 23]
    assert comp.fixed_raw[1] == pytest.approx((4.0 - 1.0) / 19.0)
    assert list(rev.w_index[:4]) == [25, 26, 27, 28]
    w = np.linspace(0, 1, D)
    d = parameters_to_dict(w, plugins)
    assert d["Compressor"]["ratio"] == pytest.approx(4.0) and d["ParametricEQ"]["our_bypass"] == w[0]
    assert d["Reverb"]["width"] == pytest.approx(w[28])
    with pytest.raises(AssertionError):  # Parameter.set_value range check (effects.py:792)
        compile_chain(load_plugins(_chain(fixed={"comp": {"ratio": 99.0}}))[0], 48000)
    with pytest.raises(ValueError, match="vst_filepath"):
        load_plugins({"x": {"num_channels": 1}})


def test_compile_chain_of_the_dasp_style_effects(capsys):
    """SURVEY row R2 as ES-path plugins: make_chain("mastering-dasp") -> chain kinds 0 / 6 / 5 with their integer options."""
    from st_ito_b200 import _lib, effects
    from st_ito_b200.engine import compile_chain
    from st_ito_b200.style_transfer import load_plugins

    plugins, D, init = load_plugins(effects.make_chain("mastering-dasp"))
    capsys.readouterr()
    assert D == (1 + 18) + (1 + 6) + (1 + 25) == len(init)
    desc, Dc = compile_chain(plugins, 48000)
    assert Dc == D and desc.num_fx == 3
    eq, comp, rev = desc.fx[0], desc.fx[1], desc.fx[2]
    assert (eq.kind, comp.kind, rev.kind) == (_lib.FX_EQ, _lib.FX_LTI_COMPRESSOR, _lib.FX_CONV_REVERB)
    assert comp.num_params == 6 and comp.num_channels == 2 and list(comp.w_index[:6]) == list(range(20, 26))
    assert list(comp.iopt) == [512, 0, 0, 0]          # lookahead_samples of effects.py:646
    assert rev.num_params == 25 and list(rev.iopt)[:2] == [96000, 0]  # 2 s impulse response, noise seed
    # the plugin's ranges are the reference's (effects.py:629-634): changing one makes it a foreign plugin
    inst = effects.BasicLTICompressor(lookahead_samples=0)
    assert effects.is_native_plugin(inst) and inst.stito_iopt == (0, 0, 0, 0)
    inst.parameters["knee_db"].max_value = 12.0
    assert not effects.is_native_plugin(inst)


def test_run_optim_style_loader_has_no_bypass_slots():
    """scripts/run_optim.py:410-437 records parameter_names without our_bypass: D = 18 for the EQ."""
    from st_ito_b200.engine import compile_chain

    plugins = _chain(("eq",))
    inst = plugins["ParametricEQ"]["class_path"]()
    plugins["ParametricEQ"].update(instance=inst, parameter_names=list(inst.parameters), num_params=18)
    desc, D = compile_chain(plugins, 48000)
    assert D == 18 and list(desc.fx[0].w_index[:18]) == list(range(18))


# -------------------------------------------------------------------------------------------- CMA
def test_cma_minimises_inside_the_box_and_is_seed_reproducible():
    from st_ito_b200 import cma

    target = np.linspace(0.1, 0.9, 12)

    def f(x):
        return float(np.sum((np.asarray(x) - target) ** 2))

    runs = []
    for _ in range(2):
        es = cma.CMAEvolutionStrategy(np.full(12, 0.5), 0.33, {"bounds": [0, 1], "popsize": 16, "seed": 5,
                                                               "verbose": -9})
        assert es.result[0] is None
        for _ in range(120):
            X = es.ask()
            assert len(X) == 16 and all(np.all((x >= 0) & (x <= 1)) for x in X)
            es.tell(X, [f(x) for x in X])
        runs.append((es.result[0].copy(), es.result[1]))
    assert runs[0][1] < 1e-6 and np.allclose(runs[0][0], target, atol=2e-3)
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1]


def test_shard_bounds_cover_the_population_once():
    from st_ito_b200.dist import shard_bounds

    for P, G in [(64, 8), (64, 1), (10, 4), (3, 8), (256, 8), (1, 2)]:
        seen = []
        for r in range(G):
            lo, hi, chunk = shard_bounds(P, G, r)
            assert 0 <= lo <= hi <= P and hi - lo <= chunk
            seen += list(range(lo, hi))
        assert seen == list(range(P))


# ------------------------------------------------------------------------------ world_size 2, gloo
_WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
rank = int(sys.argv[1]); world = int(sys.argv[2]); port = sys.argv[3]; out = sys.argv[4]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
if world > 1:
    dist.init_process_group("gloo", rank=rank, world_size=world)
from st_ito_b200 import dist as sdist, effects, style_transfer
from st_ito_b200.utils import get_param_embeds, make_synthetic_param_model

# helpers first
rows = torch.arange(5 * 3, dtype=torch.float32).reshape(5, 3)
lo, hi, _ = sdist.shard_bounds(5, world, rank)
g = sdist.all_gather_rows(rows[lo:hi], 5)
assert torch.equal(g, rows), g
b = sdist.broadcast_array(np.full(4, float(rank + 1)))
assert np.all(b == 1.0)
# more ranks than candidates: the trailing rank owns an empty shard and still takes part in the collective
lo1, hi1, _ = sdist.shard_bounds(1, world, rank)
g1 = sdist.all_gather_rows(rows[:1][lo1:hi1], 1)
assert torch.equal(g1, rows[:1]), g1

class FakeEngine:  # stands in for the libstito handle: fitness = distance of w to a hidden optimum
    calls = []
    def set_chain(self, d): pass
    def set_target_embeds(self, m, s): pass
    def set_input(self, x, min_len=0): return max(x.shape[-1], min_len)
    def eval_population(self, W, start, length, want_embeds=False, want_audio=False, in_chs=None, device_out=False):
        W = np.asarray(W); FakeEngine.calls.append((W.shape[0], start, length))
        wstar = np.linspace(0.2, 0.8, W.shape[1])
        fit = torch.tensor([float(np.sum((w - wstar) ** 2)) - 1.0 for w in W], dtype=torch.float32)
        emb = None
        if want_embeds:  # embeddings that depend on w, so that dropout changes the ranking
            base = torch.linspace(0.0, 1.0, 512)
            e = torch.stack([torch.cos(base * (1.0 + 7.0 * float(np.sum((w - wstar) ** 2)))) for w in W]) \
                if W.shape[0] else torch.zeros(0, 512)
            emb = torch.stack([e, e])
        return fit, emb, None

model = make_synthetic_param_model(seed=1)
model.stito_engine = lambda *a, **k: FakeEngine()
model.forward = lambda x: (torch.ones(x.shape[0], 512), torch.ones(x.shape[0], 512))
style_transfer.process_audio = lambda x, w, sr, plugins, normalize_stages=False: np.asarray(x, dtype=np.float32)
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    plugins, D, _ = style_transfer.load_plugins(effects.make_chain("eq"))
x = torch.randn(1, 1, 300000, generator=torch.Generator().manual_seed(0))
t = torch.randn(1, 1, 300000, generator=torch.Generator().manual_seed(1))
res = style_transfer.run_es(x, t, 48000, plugins, model, get_param_embeds, max_iters=6, popsize={popsize}, sigma0=0.33,
                            find_w0=True, seed={seed}, verbose=False, dropout={dropout}, random_crop={random_crop})
json.dump({{"fopt": res["fopt"], "wopt": list(map(float, res["wopt"])), "hist": [float(v) for v in res["fval_history"][1:]],
           "calls": FakeEngine.calls}}, open(out, "w"))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
'''


def _run_world(world, seed, tmp_path, port, popsize=10, dropout=0.0, random_crop=False):
    script = tmp_path / f"worker_{world}_{seed}.py"
    script.write_text(_WORKER.format(root=ROOT, seed=seed, popsize=popsize, dropout=dropout, random_crop=random_crop))
    outs = [tmp_path / f"out_{world}_{seed}_{r}.json" for r in range(world)]
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(port), str(outs[r])],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = [p.communicate(timeout=240)[0] for p in procs]
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    import json

    return [json.load(open(o)) for o in outs]


@pytest.mark.parametrize("seed", [0, None])
def test_run_es_population_sharding_world2_gloo(tmp_path, seed):
    """Two ranks (gloo, CPU) shard every population, all-gather the fitness and stay in lock step: identical
    results on both ranks, identical to the single-process run when the CMA-ES is seeded; with seed=None rank 0's
    population is broadcast instead (style_transfer.replicated)."""
    port = 29600 + (os.getpid() % 300) + (0 if seed is None else 1)
    two = _run_world(2, repr(seed), tmp_path, port)
    assert two[0]["fopt"] == two[1]["fopt"] and two[0]["wopt"] == two[1]["wopt"] and two[0]["hist"] == two[1]["hist"]
    # every rank evaluated half of each population of 10, on the full-length view (L > 262144, no crop)
    assert all(c == [5, 0, 300000] for c in two[0]["calls"]) and len(two[0]["calls"]) == 7
    if seed is not None:
        one = _run_world(1, repr(seed), tmp_path, port + 2)[0]
        assert one["fopt"] == two[0]["fopt"] and one["wopt"] == two[0]["wopt"]
        assert all(c == [10, 0, 300000] for c in one["calls"])
        assert one["fopt"] <= min(one["hist"]) and one["fopt"] < 0.5  # best-so-far of a distance-to-optimum objective


def test_run_es_sharding_uneven_population_dropout_and_crop_world2_gloo(tmp_path):
    """(a) a population that does not divide over the ranks (7 = 4 + 3); (b) dropout > 0 draws a different
    mask on every rank, so the fitness vector CMA-ES sees is rank 0's, broadcast; (c) the random crop is one draw for
    the whole population on every rank.  The two ranks must stay in lock step: identical trajectories."""
    port = 29950 + (os.getpid() % 40)
    two = _run_world(2, "3", tmp_path, port, popsize=7, dropout=0.3, random_crop=True)
    assert two[0]["wopt"] == two[1]["wopt"] and two[0]["hist"] == two[1]["hist"] and two[0]["fopt"] == two[1]["fopt"]
    assert [c[0] for c in two[0]["calls"]] == [4] * 7 and [c[0] for c in two[1]["calls"]] == [3] * 7
    starts0, starts1 = [c[1] for c in two[0]["calls"]], [c[1] for c in two[1]["calls"]]
    assert starts0 == starts1 and len(set(starts0)) > 1 and all(c[2] == 262144 for c in two[0]["calls"])


# ------------------------------------------------------------------------- checkpoint loading (CPU only)
def test_load_param_model_reads_a_lightning_checkpoint(tmp_path):
    """load_param_model (reference utils.py:511-551): Lightning ckpt with `encoder.`-prefixed keys next to a
    config.yaml whose class_path still says `lcap.models.panns.Cnn14`; other sub-modules' keys are dropped."""
    import yaml

    from st_ito_b200.models.panns import AFX_REP_ARGS, Cnn14
    from st_ito_b200.utils import load_param_model, make_synthetic_param_model

    src = make_synthetic_param_model(seed=11)
    sd = {f"encoder.{k}": v.clone() for k, v in src.state_dict().items()}
    sd["instance_estimator.0.weight"] = torch.zeros(4, 4)  # ParameterEstimator heads: ignored by the loader
    sd["preset_estimator.0.bias"] = torch.zeros(4)
    ckpt = tmp_path / "afx-rep.ckpt"
    torch.save({"state_dict": sd, "epoch": 3}, ckpt)
    cfg = {"model": {"class_path": "lcap.methods.param.ParameterEstimator",
                     "init_args": {"encoder": {"class_path": "lcap.models.panns.Cnn14", "init_args": dict(AFX_REP_ARGS)}}}}
    (tmp_path / "config.yaml").write_text(yaml.safe_dump(cfg))
    m = load_param_model(str(ckpt))
    assert isinstance(m, Cnn14) and not m.training
    got = m.state_dict()
    for k, v in src.state_dict().items():
        assert torch.equal(got[k], v), k
    assert got["spectrogram_extractor.stft.conv_real.weight"].shape == (1025, 1, 2048)
    assert got["logmel_extractor.melW"].shape == (1025, 128)
    with pytest.raises(FileNotFoundError, match="afx-rep.ckpt"):
        load_param_model(str(tmp_path / "missing" / "afx-rep.ckpt"))


# ------------------------------------------------------------------------- native CMA-ES (C ABI, host only)
def test_native_cma_matches_the_numpy_statement_of_the_algorithm():
    """stito_cma_* (cma_host.cpp) against st_ito_b200.cma.PyCMAEvolutionStrategy: same update equations, so feeding both
    the SAME populations and fitness values must give the same mean / sigma / covariance scale; the eigensolver is
    checked against numpy; ask() is feasible, deterministic per seed and different across seeds."""
    import ctypes

    from st_ito_b200 import _lib, cma

    L = _lib.lib()
    rng = np.random.RandomState(0)
    for n in (1, 2, 7, 29, 50):
        M = rng.randn(n, n)
        A = M @ M.T + 0.1 * np.eye(n)
        V, d = np.empty((n, n)), np.empty(n)
        assert L.stito_cma_eig(A.ctypes.data, n, V.ctypes.data, d.ctypes.data) == 0
        assert np.abs((V * d) @ V.T - A).max() <= 1e-12 * np.abs(A).max()
        assert np.abs(V.T @ V - np.eye(n)).max() <= 1e-12
        assert np.allclose(np.sort(d), np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-12)

    target = np.linspace(0.1, 0.9, 12)
    f = lambda x: float(np.sum((x - target) ** 2))
    opts = {"bounds": [0, 1], "popsize": 16, "seed": 5, "verbose": -9}
    nat = cma.CMAEvolutionStrategy(np.full(12, 0.5), 0.33, dict(opts))
    assert isinstance(nat, cma.NativeCMAEvolutionStrategy) and nat.result[0] is None
    py = cma.CMAEvolutionStrategy(np.full(12, 0.5), 0.33, dict(opts, implementation="numpy"))
    assert isinstance(py, cma.PyCMAEvolutionStrategy)
    for it in range(40):
        X = nat.ask()
        assert len(X) == 16 and all(np.all((x >= 0) & (x <= 1)) for x in X)
        fv = [f(x) for x in X]
        # drive the numpy implementation with the native one's genotypes: replace its sample, keep its update
        py.ask()
        py._geno = np.ascontiguousarray(_native_geno(nat, L))
        nat.tell(X, fv)
        py.tell(X, fv)
        assert np.allclose(py.result.xfavorite, nat.result.xfavorite, rtol=1e-9, atol=1e-12), it
        assert abs(py.sigma - nat.sigma) <= 1e-9 * py.sigma
        assert np.allclose(py.result.stds, nat.result.stds, rtol=1e-8)
    assert nat.result[1] == py.result[1] and np.array_equal(nat.result[0], py.result[0])
    # optimisation quality + determinism
    runs = []
    for seed in (5, 5, 6):
        es = cma.CMAEvolutionStrategy(np.full(12, 0.5), 0.33, dict(opts, seed=seed))
        for _ in range(120):
            X = es.ask()
            es.tell(X, [f(x) for x in X])
        runs.append((es.result[0].copy(), es.result[1]))
    assert runs[0][1] < 1e-6 and np.allclose(runs[0][0], target, atol=2e-3)
    assert np.array_equal(runs[0][0], runs[1][0]) and not np.array_equal(runs[0][0], runs[2][0])
    with pytest.raises(ValueError):
        nat.tell(X, fv)  # no preceding ask()


def _native_geno(es, L):
    """Search points (before the box map) of the native strategy's last ask(): stito_cma_geno."""
    G = np.empty((es.popsize, es.N))
    assert L.stito_cma_geno(es._h, G.ctypes.data) == 0
    return G
