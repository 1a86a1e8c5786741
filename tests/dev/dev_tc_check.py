"""Developer check: tensor-core (fp16x3) encoder vs the fp32 CUDA-core mode on the same inputs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from st_ito_b200.utils import make_synthetic_param_model
from tests.signals import test_signal

L = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
from oracle import cnn14
m = make_synthetic_param_model(seed=3, conv_gain=2.0)
ref = cnn14.make_encoder(seed=3, conv_gain=2.0)
cnn14.centre_heads(ref)
with torch.no_grad():
    m.fc_mid.bias.copy_(ref.fc_mid.bias); m.fc_side.bias.copy_(ref.fc_side.bias)
eng = m.stito_engine(0)
x = torch.from_numpy(np.stack([test_signal(2, L, seed=100 + b) for b in range(B)]))
x = x / x.abs().amax(dim=(1, 2), keepdim=True)
out = {}
for prec in (0, 1):
    eng.set_precision(prec)
    t0 = time.time()
    mid, side = eng.embed(x)
    torch.cuda.synchronize()
    out[prec] = (mid.numpy(), side.numpy())
    print("precision", prec, "time %.3f s" % (time.time() - t0), "norms", np.linalg.norm(mid.numpy(), axis=1)[:2])
with torch.no_grad():
    rm, rs = ref(x)
for k, name in enumerate(("mid", "side")):
    a, b, o = out[1][k], out[0][k], (rm, rs)[k].numpy()
    print(name, "rel err tc vs simt: %.3e   tc vs oracle: %.3e   simt vs oracle: %.3e" % (
        np.linalg.norm(a - b) / np.linalg.norm(b), np.linalg.norm(a - o) / np.linalg.norm(o), np.linalg.norm(b - o) / np.linalg.norm(o)))
