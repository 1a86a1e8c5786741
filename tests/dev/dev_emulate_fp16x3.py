"""Developer probe: intrinsic error of the error-compensated fp16x3 operand split (exact accumulation)
against fp64 truth, on the golden fitness fixture (centred heads).  CPU only."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cnn14, dsp
from tests.signals import test_signal

torch.set_num_threads(8)
SR = 48000
ref = cnn14.make_encoder(seed=3, bn_stats=True, conv_gain=2.0)
cnn14.centre_heads(ref)
ref.eval()
g = np.load("tests/golden/fitness.npz")
dsp.build()
plugins, D, _ = dsp.load_plugins(dsp.make_plugins(["eq"]))
x = test_signal(2, 40000, seed=5); x = x / np.abs(x).max()
aud = torch.stack([torch.from_numpy(dsp.process_audio(x, w, SR, plugins)) for w in g["W"]])
with torch.no_grad():
    feats = ref.logmel(aud)

def split(t, scale):
    v = (t * scale).float()
    hi = v.half().float()
    lo = (v - hi).half().float()
    return hi.double(), lo.double()

def body(feats, mode):
    x = feats.double() if mode != "fp32" else feats
    bs, chs = 8, 2
    for i in range(6):
        blk = getattr(ref, f"conv_block{i+1}")
        for j, (conv, bn) in enumerate(((blk.conv1, blk.bn1), (blk.conv2, blk.bn2))):
            sc = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = conv.weight * sc[:, None, None, None]
            b = bn.bias - bn.running_mean * sc
            if mode == "fp32":
                x = F.relu(F.conv2d(x, w, b, padding=1))
            elif mode == "fp64":
                x = F.relu(F.conv2d(x, w.double(), b.double(), padding=1))
            elif mode in ("x3", "x1", "x2a", "x2w"):
                if i == 0 and j == 0:
                    x = F.relu(F.conv2d(x.float(), w, b, padding=1)).double()
                    continue
                mx = w.abs().max().item()
                shift = -int(np.ceil(np.log2(mx)))
                wh, wl = split(w, 2.0 ** shift)
                ah, al = split(x, 64.0)
                acc = F.conv2d(ah, wh, padding=1)
                if mode == "x3":
                    acc = acc + (F.conv2d(al, wh, padding=1) + F.conv2d(ah, wl, padding=1))
                elif mode == "x2a":
                    acc = acc + F.conv2d(al, wh, padding=1)
                elif mode == "x2w":
                    acc = acc + F.conv2d(ah, wl, padding=1)
                y = acc.float() * np.float32(2.0 ** -shift / 64.0) + b[None, :, None, None]
                x = F.relu(y).double()
        if i < 5:
            x = F.avg_pool2d(x, 2)
            if mode in ("x3", "x1", "x2a", "x2w"):
                x = x.float().double()
    x = x.mean(dim=3)
    x = x.max(dim=2).values + x.mean(dim=2)
    x = x.view(bs, chs, -1).to(ref.fc_mid.weight.dtype if mode == "fp32" else torch.float64)
    if mode == "fp32":
        mid = ref.fc_mid(x[:, 0]); side = ref.fc_side(x[:, 1])
    else:
        mid = x[:, 0] @ ref.fc_mid.weight.double().T + ref.fc_mid.bias.double()
        side = x[:, 1] @ ref.fc_side.weight.double().T + ref.fc_side.bias.double()
    mid = mid / mid.norm(dim=-1, keepdim=True); side = side / side.norm(dim=-1, keepdim=True)
    return mid.double().numpy(), side.double().numpy()

def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
with torch.no_grad():
    truth = body(feats, "fp64")
    for mode in sys.argv[1:] or ["fp32", "x3", "x1", "x2a", "x2w"]:
        r = body(feats, mode)
        print(mode, "mid %.3e side %.3e" % (rel(r[0], truth[0]), rel(r[1], truth[1])),
              "| vs golden mid %.3e" % rel(r[0], g["mid"]))
    print("golden vs truth mid %.3e side %.3e" % (rel(g["mid"], truth[0]), rel(g["side"], truth[1])))
