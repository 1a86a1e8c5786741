"""Developer probe (GPU): how much margin do the 1e-4 embedding / fitness gates have?  Prints the measured relative
errors of the tensor-core path (precision 1) and of the fp32 CUDA-core mode against the golden fixture / the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cnn14
from st_ito_b200.engine import compile_chain
from st_ito_b200.utils import get_param_embeds, make_synthetic_param_model
from tests.signals import test_signal
from tests.test_gpu_parity import native_plugins, rel_err

SR = 48000
ours = make_synthetic_param_model(seed=3, bn_stats=True, conv_gain=2.0)
ref = cnn14.make_encoder(seed=3, bn_stats=True, conv_gain=2.0)
cnn14.centre_heads(ref)
with torch.no_grad():
    ours.fc_mid.bias.copy_(ref.fc_mid.bias); ours.fc_side.bias.copy_(ref.fc_side.bias)
g = np.load(os.path.join(ROOT, "tests", "golden", "fitness.npz"))
eng = ours.stito_engine()
plugins, D, _ = native_plugins(["eq"])
x = test_signal(2, 40000, seed=5); x = x / np.abs(x).max()
desc, _ = compile_chain(plugins, SR); eng.set_chain(desc); eng.set_input(x)
eng.set_target_embeds(torch.from_numpy(g["tgt_mid"][0]), torch.from_numpy(g["tgt_side"][0]))
for prec in (0, 1):
    eng.set_precision(prec)
    fit, emb, _ = eng.eval_population(g["W"], 0, 40000, want_embeds=True)
    print(f"golden centred fixture, precision {prec}: mid {rel_err(emb[0].numpy(), g['mid']):.2e} side "
          f"{rel_err(emb[1].numpy(), g['side']):.2e} fitness {np.abs(fit.numpy() - g['fitness']).max() / np.abs(g['fitness']).max():.2e}")
xl = torch.from_numpy(np.stack([test_signal(2, 480000, seed=200 + b) for b in range(2)]))
want = cnn14.get_param_embeds(xl.clone(), ref, SR)
for prec in (0, 1):
    ours.stito_engine().set_precision(prec)
    got = get_param_embeds(xl.clone(), ours, SR)
    print(f"10 s stereo centred model, precision {prec}: mid {rel_err(got['mid'].numpy(), want['mid'].numpy()):.2e} side "
          f"{rel_err(got['side'].numpy(), want['side'].numpy()):.2e}")
