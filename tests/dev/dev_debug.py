"""Developer diagnostics: where do embedding / fitness differences vs the oracle come from?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import contextlib, io
import numpy as np, torch
from oracle import cnn14, dsp
from st_ito_b200 import effects
from st_ito_b200.engine import compile_chain
from st_ito_b200.style_transfer import load_plugins, process_audio
from st_ito_b200.utils import make_synthetic_param_model
from tests.signals import test_signal

SR = 48000
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)

ours = make_synthetic_param_model(seed=3)
ref = cnn14.make_encoder(seed=3)
cnn14.centre_heads(ref)
with torch.no_grad():
    ours.fc_mid.bias.copy_(ref.fc_mid.bias); ours.fc_side.bias.copy_(ref.fc_side.bias)
eng = ours.stito_engine(0)

# 1. batch invariance
x = torch.from_numpy(np.stack([test_signal(2, 100000, seed=100 + b) for b in range(5)]))
x = x / x.abs().amax(dim=(1, 2), keepdim=True)
for prec in (0, 1):
    eng.set_precision(prec)
    m5, s5 = eng.embed(x)
    m1, s1 = eng.embed(x[2:3])
    print(f"[1] precision {prec}: item alone vs in batch: max abs diff mid {float((m5[2]-m1[0]).abs().max()):.3e} side {float((s5[2]-s1[0]).abs().max()):.3e}  |mid| {float(m1.norm()):.3f}")

# 2. chain + encoder error budget
with contextlib.redirect_stdout(io.StringIO()):
    plugins, D, _ = load_plugins(effects.make_chain("mastering-pb"))
oplugins, _, _ = dsp.load_plugins(dsp.make_plugins(["eq", "comp", "reverb"]))
L = 100000
xs = test_signal(2, L, seed=41); xs = xs / np.abs(xs).max()
rng = np.random.RandomState(77)
w_star, W = rng.rand(D), rng.rand(6, D)
tgt = dsp.process_audio(xs, w_star, SR, oplugins)
te = cnn14.get_param_embeds(torch.from_numpy(tgt[None].copy()), ref, SR)
xpad = np.pad(xs, ((0, 0), (0, 262144 - L)))
oa = np.stack([dsp.process_audio(xpad, w, SR, oplugins) for w in W])
oe = cnn14.get_param_embeds(torch.from_numpy(oa.copy()), ref, SR)
of = cnn14.fitness(oe, te).numpy()
desc, _ = compile_chain(plugins, SR)
eng.set_chain(desc); eng.set_input(xs, min_len=262144); eng.set_target(tgt)
for prec in (0, 1):
    eng.set_precision(prec)
    fit, emb, aud = eng.eval_population(W, 0, 262144, want_embeds=True, want_audio=True, in_chs=2)
    print(f"[2] precision {prec}: audio max abs diff vs oracle {np.abs(aud.numpy()-oa).max():.3e}")
    print("    fitness gpu   ", fit.numpy())
    print("    fitness oracle", of)
    print("    embed rel err mid %.3e side %.3e" % (rel(emb[0].numpy(), oe['mid'].numpy()), rel(emb[1].numpy(), oe['side'].numpy())))
    # encoder only: oracle encoder on the GPU-rendered audio
    ge = cnn14.get_param_embeds(aud.clone(), ref, SR)
    print("    encoder-only rel err (oracle enc on gpu audio vs gpu enc) mid %.3e side %.3e" % (rel(emb[0].numpy(), ge['mid'].numpy()), rel(emb[1].numpy(), ge['side'].numpy())))
    print("    dsp-only rel err (oracle enc: gpu audio vs oracle audio)  mid %.3e side %.3e" % (rel(ge['mid'].numpy(), oe['mid'].numpy()), rel(ge['side'].numpy(), oe['side'].numpy())))
    # per effect audio diff
for kinds in (["eq"], ["eq", "comp"], ["comp"], ["reverb"]):
    with contextlib.redirect_stdout(io.StringIO()):
        pl, Dk, _ = load_plugins({n: effects.make_chain("basic")[n] for n in [{"eq": "ParametricEQ", "comp": "Compressor", "reverb": "Reverb"}[k] for k in kinds]})
    opl, _, _ = dsp.load_plugins(dsp.make_plugins(kinds))
    w = np.random.RandomState(3).rand(Dk)
    y = process_audio(xpad, w, SR, pl); yo = dsp.process_audio(xpad, w, SR, opl)
    d = np.abs(y - yo)
    print(f"[3] chain {kinds}: max abs diff {d.max():.3e}  mean abs diff {d.mean():.3e}  exact-equal fraction {(y==yo).mean():.4f}")
# logmel diff
lm = eng.logmel(torch.from_numpy(oa[:2].copy()))
with torch.no_grad():
    lo = ref.logmel(torch.from_numpy(oa[:2].copy()))[:, 0]
print("[4] logmel max abs diff %.3e mean %.3e" % (float((lm - lo).abs().max()), float((lm - lo).abs().mean())))
