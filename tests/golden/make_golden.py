"""Generate tests/golden/*.npz by EXECUTING the reference's own code.

Run in the build container only (needs /root/reference; see oracle/ref_import.py):

    python tests/golden/make_golden.py

What executes from the reference, unmodified: ``biqaud``, ``parametric_eq``,
``Parameter``, ``BasicParametricEQ``, ``load_plugins``, ``process_audio``,
``parameters_to_dict`` (st_ito/effects.py, st_ito/style_transfer.py) and the
``Cnn14`` module (st_ito/models/panns.py) with this repo's torchlibrosa
restatement as its front-end.  The fixtures pin oracle/ to the reference; the
GPU tests then compare the CUDA path with oracle/ live.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import cnn14, ref_import  # noqa: E402
from tests.signals import eq_corner_vectors, test_signal  # noqa: E402

effects, st, panns = ref_import.load()
SR = 48000


def eq_plugins():
    p = {"ParametricEQ": {"class_path": effects.BasicParametricEQ, "num_params": None, "num_channels": 1,
                          "fixed_parameters": {}}}
    return st.load_plugins(p)


def golden_biquad():
    rows = []
    rng = np.random.RandomState(7)
    cases = []
    for kind in ("low_shelf", "peaking", "high_shelf"):
        for g in (-24.0, 24.0):
            for fc in (20.0, 18000.0):
                for q in (0.1, 4.0):
                    cases.append((g, fc, q, kind))
        for _ in range(32):
            cases.append((rng.uniform(-24, 24), rng.uniform(20, 18000), rng.uniform(0.1, 4), kind))
    for g, fc, q, kind in cases:
        b, a = effects.biqaud(g, fc, q, SR, kind)
        rows.append([g, fc, q, {"low_shelf": 0, "peaking": 1, "high_shelf": 2}[kind], *b, *a])
    np.savez_compressed(os.path.join(HERE, "biquad.npz"), table=np.array(rows, dtype=np.float64))


def golden_eq():
    plugins, D, init = eq_plugins()
    out = {"D": D, "init": np.array(init)}
    rng = np.random.RandomState(11)
    ws = [rng.rand(D) for _ in range(3)] + eq_corner_vectors(D)
    out["W"] = np.array(ws)
    for chs in (1, 2):
        for L in (4096, 262144, 480000):
            x = test_signal(chs, L, seed=L + chs)
            ys = [st.process_audio(x.copy(), w, SR, plugins) for w in ws]
            y = np.stack(ys)  # [n, chs, L], peak == 1
            if L == 4096:
                out[f"y_{chs}_{L}"] = y
            else:  # keep the fixture small: strided samples + the head + fp64 energy
                out[f"ys_{chs}_{L}"] = y[:, :, ::997].copy()
                out[f"yh_{chs}_{L}"] = y[:, :, :512].copy()
                out[f"ye_{chs}_{L}"] = (y.astype(np.float64) ** 2).sum(axis=-1)
    out["param_dict_last"] = np.array([v for v in st.parameters_to_dict(ws[0], plugins)["ParametricEQ"].values()])
    np.savez_compressed(os.path.join(HERE, "eq.npz"), **out)


def golden_cnn14():
    torch.manual_seed(0)
    out = {}
    for tag, bn in (("plain", False), ("bnstats", True)):
        ref = panns.Cnn14(**cnn14.AFX_REP_ARGS).eval()
        ref.load_state_dict(cnn14.make_encoder(seed=3, bn_stats=bn).state_dict())
        for chs in (1, 2):
            x = torch.from_numpy(np.stack([test_signal(chs, 40000, seed=100 + b) for b in range(2)]))
            x = x / x.abs().amax(dim=(1, 2), keepdim=True)
            with torch.no_grad():
                mid, side = ref(x)
            out[f"{tag}_mid_{chs}"] = mid.numpy()
            out[f"{tag}_side_{chs}"] = side.numpy()
    np.savez_compressed(os.path.join(HERE, "cnn14.npz"), **out)


def golden_fitness():
    """P=8 EQ-only population, L=40000 stereo (>= 32 frames needed by five 2x2 pools after padding),
    scored by the reference's Cnn14 with seeded weights (conv_gain=2) and centred heads
    (oracle.cnn14.centre_heads): a well-conditioned ranking, see oracle.cnn14.make_encoder."""
    plugins, D, _ = eq_plugins()
    ref = panns.Cnn14(**cnn14.AFX_REP_ARGS).eval()
    ref.load_state_dict(cnn14.make_encoder(seed=3, bn_stats=True, conv_gain=2.0).state_dict())
    cnn14.centre_heads(ref)
    x = test_signal(2, 40000, seed=5)
    x = x / np.abs(x).max()
    w_star = np.random.RandomState(1234).rand(D)
    tgt = st.process_audio(x.copy(), w_star, SR, plugins)
    W = np.random.RandomState(99).rand(8, D)
    outs = np.stack([st.process_audio(x.copy(), w, SR, plugins) for w in W])

    def embed(a):  # utils.get_param_embeds cannot be imported; its 4 steps, on reference Cnn14
        a = torch.from_numpy(a.copy())
        for b in range(a.shape[0]):
            a[b] /= a[b].abs().max().clamp(1e-8)
        with torch.no_grad():
            m, s = ref(a)
        return {"mid": torch.nn.functional.normalize(m, dim=-1), "side": torch.nn.functional.normalize(s, dim=-1)}

    te, oe = embed(tgt[None]), embed(outs)
    f = torch.stack([-torch.cosine_similarity(oe[k], te[k], dim=-1) for k in oe]).mean(0).numpy()
    np.savez_compressed(os.path.join(HERE, "fitness.npz"), W=W, w_star=w_star, fitness=f,
                        argsort=np.argsort(f, kind="stable"), mid=oe["mid"].numpy(), side=oe["side"].numpy(),
                        tgt_mid=te["mid"].numpy(), tgt_side=te["side"].numpy(),
                        bias_mid=ref.fc_mid.bias.detach().numpy(), bias_side=ref.fc_side.bias.detach().numpy())


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (golden_biquad, golden_eq, golden_cnn14, golden_fitness):
        if not only or fn.__name__.replace("golden_", "") in only:
            fn()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
