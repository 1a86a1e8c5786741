"""Developer probe (CPU): embedding error if the deep conv layers used Winograd F(2x2,3x3) with fp16x3-split
transformed operands (exact accumulation), on the centred-head golden fixture."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cnn14, dsp
from tests.signals import test_signal

torch.set_num_threads(8)
SR = 48000
ref = cnn14.make_encoder(seed=3, bn_stats=True, conv_gain=2.0); cnn14.centre_heads(ref); ref.eval()
g = np.load("tests/golden/fitness.npz")
dsp.build()
plugins, D, _ = dsp.load_plugins(dsp.make_plugins(["eq"]))
x = test_signal(2, 40000, seed=5); x = x / np.abs(x).max()
aud = torch.stack([torch.from_numpy(dsp.process_audio(x, w, SR, plugins)) for w in g["W"]])
with torch.no_grad():
    feats = ref.logmel(aud)

F4 = len(sys.argv) > 1 and sys.argv[1] == "f4"
if F4:   # F(4x4, 3x3)
    BT = torch.tensor([[4, 0, -5, 0, 1, 0], [0, -4, -4, 1, 1, 0], [0, 4, -4, -1, 1, 0], [0, -2, -1, 2, 1, 0],
                       [0, 2, -1, -2, 1, 0], [0, 4, 0, -5, 0, 1]], dtype=torch.float64)
    G = torch.tensor([[1 / 4, 0, 0], [-1 / 6, -1 / 6, -1 / 6], [-1 / 6, 1 / 6, -1 / 6], [1 / 24, 1 / 12, 1 / 6],
                      [1 / 24, -1 / 12, 1 / 6], [0, 0, 1]], dtype=torch.float64)
    AT = torch.tensor([[1, 1, 1, 1, 1, 0], [0, 1, -1, 2, -2, 0], [0, 1, 1, 4, 4, 0], [0, 1, -1, 8, -8, 1]], dtype=torch.float64)
else:    # F(2x2, 3x3)
    BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float64)
    G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
    AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float64)
OT = AT.shape[0]; IT = BT.shape[0]

def split(t, scale):
    v = (t * scale).float(); hi = v.half().float(); lo = (v - hi).half().float()
    return hi.double(), lo.double()

def winograd_conv(xin, w, mode):
    """xin [N,C,H,W] float64 (values as stored: fp32-exact), w [Co,Ci,3,3] fp32. Returns conv (pad 1) fp64."""
    N, C, H, W = xin.shape
    Hp, Wp = (H + OT - 1) // OT * OT, (W + OT - 1) // OT * OT
    xp = F.pad(xin, (1, 1 + Wp - W, 1, 1 + Hp - H))
    tiles = xp.unfold(2, IT, OT).unfold(3, IT, OT)        # [N,C,th,tw,IT,IT]
    V = torch.einsum("ij,nctujk,lk->nctuil", BT, tiles.double(), BT)   # B^T d B
    U = torch.einsum("ij,ocjk,lk->ocil", G, w.double(), G)             # G g G^T  [Co,Ci,4,4]
    if mode == "exact":
        M = torch.einsum("nctuil,ocil->notuil", V, U)
    else:
        V32 = V.float()                                   # transform done in fp32 on CUDA cores
        U32 = U.float()
        sV = 64.0 / (V32.abs().max().item() / xin.abs().max().item() if F4 else 1.0) if False else 64.0 / (16.0 if F4 else 1.0)
        mx = U32.abs().max().item(); shift = -int(np.ceil(np.log2(mx)))
        vh, vl = split(V32, sV); uh, ul = split(U32, 2.0 ** shift)
        M = torch.einsum("nctuil,ocil->notuil", vh, uh) + (torch.einsum("nctuil,ocil->notuil", vl, uh) + torch.einsum("nctuil,ocil->notuil", vh, ul))
        M = M.float().double() * (2.0 ** -shift / sV)     # fp32 accumulator
    Y = torch.einsum("ij,notujk,lk->notuil", AT, M, AT)   # [N,Co,th,tw,2,2]
    Y = Y.permute(0, 1, 2, 4, 3, 5).reshape(N, w.shape[0], Hp, Wp)[:, :, :H, :W]
    return Y

def body(feats, wino_from, mode):
    x = feats.double(); bs, chs = 8, 2
    li = 0
    for i in range(6):
        blk = getattr(ref, f"conv_block{i+1}")
        for j, (conv, bn) in enumerate(((blk.conv1, blk.bn1), (blk.conv2, blk.bn2))):
            sc = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = conv.weight * sc[:, None, None, None]; b = bn.bias - bn.running_mean * sc
            if li >= wino_from:
                y = winograd_conv(x, w, mode).float() + b[None, :, None, None]
                x = F.relu(y).double()
            else:
                x = F.relu(F.conv2d(x, w.double(), b.double(), padding=1))
                if mode != "truth": x = x.float().double()
            li += 1
        if i < 5: x = F.avg_pool2d(x, 2)
    x = x.mean(dim=3); x = x.max(dim=2).values + x.mean(dim=2); x = x.view(bs, chs, -1)
    mid = x[:, 0] @ ref.fc_mid.weight.double().T + ref.fc_mid.bias.double()
    side = x[:, 1] @ ref.fc_side.weight.double().T + ref.fc_side.bias.double()
    return (mid / mid.norm(dim=-1, keepdim=True)).numpy(), (side / side.norm(dim=-1, keepdim=True)).numpy()

rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
with torch.no_grad():
    truth = body(feats, 99, "truth")
    for wf in ((9,) if F4 else (8, 6, 2)):
        for mode in ("exact", "x3"):
            r = body(feats, wf, mode)
            print(f"winograd from layer {wf:2d} mode {mode:5s}: mid {rel(r[0], truth[0]):.3e} side {rel(r[1], truth[1]):.3e}")
