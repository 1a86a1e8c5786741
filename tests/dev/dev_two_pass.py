"""Developer probe (GPU): per-layer two-pass variants of the fp16x3 convolution.  For every conv layer
l = 1..11 the A_lo*B_hi pass (activation low parts) or the A_hi*B_lo pass (weight low parts) is dropped on that layer alone
(STITO_TC_DROP_ALO / STITO_TC_DROP_BLO are read per launch) and the embedding error against the fp32 oracle is measured on
the two fixtures of dev_margins2.py.  Prints one JSON line per variant."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cnn14, dsp
from st_ito_b200.utils import get_param_embeds, make_synthetic_param_model
from tests.signals import test_signal
from tests.dev.dev_margins2 import heavy_tail_

SR = 48000


def main():
    dsp.build()
    eq, D, _ = dsp.load_plugins(dsp.make_plugins(["eq"]))
    rng = np.random.RandomState(11)
    base = test_signal(2, 40000, seed=500)
    x_short = torch.from_numpy(np.stack([dsp.process_audio(base, rng.rand(D), SR, eq) for b in range(6)]))
    x_long = torch.from_numpy(np.stack([test_signal(2, 480000, seed=500 + b) for b in range(2)]))
    for fixture in ("xavier", "heavy"):
        ours = make_synthetic_param_model(seed=3, bn_stats=True, conv_gain=2.0)
        ref = cnn14.make_encoder(seed=3, bn_stats=True, conv_gain=2.0)
        if fixture == "heavy":
            heavy_tail_(ours)
            heavy_tail_(ref)
        cnn14.centre_heads(ref)
        with torch.no_grad():
            ours.fc_mid.bias.copy_(ref.fc_mid.bias)
            ours.fc_side.bias.copy_(ref.fc_side.bias)
        ours.stito_engine().set_precision(1)
        want = {"short": cnn14.get_param_embeds(x_short.clone(), ref, SR), "long": cnn14.get_param_embeds(x_long.clone(), ref, SR)}

        def measure():
            out = {}
            for tag, x in (("short", x_short), ("long", x_long)):
                got = get_param_embeds(x.clone(), ours, SR)
                e = []
                for k in ("mid", "side"):
                    a, b = got[k].numpy().astype(np.float64), want[tag][k].numpy().astype(np.float64)
                    e.append(float((np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max()))
                out[tag] = max(e)
            return out

        variants = [("none", 0, 0)] + [(f"alo_l{l}", 1 << l, 0) for l in range(1, 12)] + \
                   [(f"blo_l{l}", 0, 1 << l) for l in range(1, 12)] + [("alo_all", 0xFFE, 0), ("blo_all", 0, 0xFFE)]
        for name, ma, mb in variants:
            os.environ["STITO_TC_DROP_ALO"] = hex(ma)
            os.environ["STITO_TC_DROP_BLO"] = hex(mb)
            print(json.dumps({"fixture": fixture, "variant": name, **measure()}), flush=True)
        os.environ["STITO_TC_DROP_ALO"] = "0"
        os.environ["STITO_TC_DROP_BLO"] = "0"
        ours.stito_engine().close()


if __name__ == "__main__":
    main()
