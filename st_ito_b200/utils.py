"""Embedding API of the ES path, mirroring ``st_ito/utils.py:444-551`` of the reference:
``load_param_model`` and ``get_param_embeds``.  The encoder forward runs in libstito.
"""
from __future__ import annotations

import os
from importlib import import_module

import torch
import yaml

# -------- self-supervised parameter estimation model -------- #


def get_param_embeds(
    x: torch.Tensor,
    model: torch.nn.Module,
    sample_rate: float,
    requires_grad: bool = False,
    peak_normalize: bool = False,
    dropout: float = 0.0,
):
    """reference utils.py:444-508: x [bs, chs, seq_len] -> {"mid": [bs, E], "side": [bs, E]}, L2-normalised."""
    bs, chs, seq_len = x.shape
    x_device = x

    # move audio to model device
    x = x.type_as(next(model.parameters()))

    if sample_rate != 48000:
        import torchaudio

        x = torchaudio.functional.resample(x, sample_rate, 48000)

    # peak normalize each batch item (in place when dtype/device already match, like the reference)
    for batch_idx in range(bs):
        x[batch_idx, ...] /= x[batch_idx, ...].abs().max().clamp(1e-8)

    if requires_grad:
        raise NotImplementedError("the B200 encoder is inference-only (run_autodiff is outside this path)")
    with torch.no_grad():
        mid_embeddings, side_embeddings = model(x)

    if dropout > 0.0:
        mid_embeddings = torch.nn.functional.dropout(mid_embeddings, p=dropout, training=True)
        side_embeddings = torch.nn.functional.dropout(side_embeddings, p=dropout, training=True)

    # check for nan (if / elif exactly as the reference :492-497)
    if torch.isnan(mid_embeddings).any():
        print("Warning: NaNs found in mid_embeddings")
        mid_embeddings = torch.nan_to_num(mid_embeddings)
    elif torch.isnan(side_embeddings).any():
        print("Warning: NaNs found in side_embeddings")
        side_embeddings = torch.nan_to_num(side_embeddings)

    # l2 normalize
    mid_embeddings = torch.nn.functional.normalize(mid_embeddings, p=2, dim=-1)
    side_embeddings = torch.nn.functional.normalize(side_embeddings, p=2, dim=-1)

    return {
        "mid": mid_embeddings.type_as(x_device),
        "side": side_embeddings.type_as(x_device),
    }


def load_param_model(ckpt_path: str = None, use_gpu: bool = False):
    """reference utils.py:511-551: Lightning checkpoint + sibling config.yaml -> Cnn14 in eval mode.

    The reference downloads ``afx-rep.ckpt`` / ``config.yaml`` from HuggingFace when they are
    missing; this build runs offline, so a missing checkpoint is an error that says where to put it.
    """
    if ckpt_path is None:  # look in tmp directory
        ckpt_path = os.path.join(os.getcwd(), "tmp", "afx-rep.ckpt")
    if not os.path.isfile(ckpt_path):
        raise FileNotFoundError(
            f"{ckpt_path} not found.  Fetch afx-rep.ckpt and config.yaml from "
            "https://huggingface.co/csteinmetz1/afx-rep and place them side by side "
            "(or use st_ito_b200.utils.make_synthetic_param_model for seeded random weights).")

    config_path = os.path.join(os.path.dirname(ckpt_path), "config.yaml")
    with open(config_path) as f:
        config = yaml.safe_load(f)

    encoder_configs = config["model"]["init_args"]["encoder"]
    module_path, class_name = encoder_configs["class_path"].rsplit(".", 1)
    # lcap.models.panns.Cnn14 / st_ito.models.panns.Cnn14 -> st_ito_b200.models.panns.Cnn14
    module_path = module_path.replace("lcap", "st_ito_b200").replace("st_ito.", "st_ito_b200.")
    if module_path == "st_ito":
        module_path = "st_ito_b200"
    module = import_module(module_path)
    model = getattr(module, class_name)(**encoder_configs["init_args"])

    checkpoint = torch.load(ckpt_path, map_location="cpu", weights_only=False)

    state_dict = {}
    for k, v in checkpoint["state_dict"].items():
        if k.startswith("encoder"):
            state_dict[k.replace("encoder.", "", 1)] = v

    model.load_state_dict(state_dict)
    model.eval()

    if use_gpu:
        model.cuda()

    return model


def centre_heads(model, clip: torch.Tensor = None):
    """Set the head biases to ``b = -W @ mu`` where ``mu`` is the pooled feature vector of a calibration clip.

    Post-ReLU pooled features share a large common-mode component, so with random weights every embedding is nearly
    parallel to every other (cosine fitness -1 +- 1e-8: no ranking to speak of).  Cancelling the
    calibration mean spreads the fitness over O(1) the way a trained encoder does.  Needs the B200: the raw
    embedding of the clip with zero biases IS ``W @ mu`` (one encoder forward through libstito).
    """
    if clip is None:
        g = torch.Generator().manual_seed(777)
        t = torch.arange(40000, dtype=torch.float32) / 48000.0
        tone = sum(torch.sin(2 * torch.pi * 110.0 * (2 ** k) * t) / (k + 1) for k in range(5))
        clip = torch.stack([0.5 * tone + 0.1 * torch.randn(40000, generator=g),
                            0.3 * tone + 0.1 * torch.randn(40000, generator=g)])[None]
        clip = clip / clip.abs().max()
    with torch.no_grad():
        model.fc_mid.bias.zero_()
        model.fc_side.bias.zero_()
        mid, side = model(clip.clone())
        model.fc_mid.bias.copy_(-mid[0].to(model.fc_mid.bias))
        model.fc_side.bias.copy_(-side[0].to(model.fc_side.bias))
    return model


def make_synthetic_param_model(seed: int = 0, bn_stats: bool = True, use_gpu: bool = False, conv_gain: float = 1.0):
    """AFx-Rep architecture with seeded random weights (no checkpoint is obtainable offline).

    Convolutions / heads use the reference's Xavier-uniform initialisers (panns.py:10-22); with
    ``bn_stats`` every BatchNorm gets non-trivial seeded affine parameters and running statistics so
    that BatchNorm folding is exercised.  ``conv_gain`` scales every convolution weight (gain 2 keeps
    the input-dependent part of the activations alive through 12 ReLU layers, which Xavier gain 1 does
    not).  Same construction as oracle/cnn14.py:make_encoder, so the two produce identical state_dicts
    for the same arguments.
    """
    from .models.panns import AFX_REP_ARGS, Cnn14

    g = torch.Generator().manual_seed(seed)
    prev = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        m = Cnn14(**AFX_REP_ARGS)
    finally:
        torch.random.set_rng_state(prev)
    if conv_gain != 1.0:
        for mod in m.modules():
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.data.mul_(conv_gain)
    if bn_stats:
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                n = mod.num_features
                mod.weight.data = 0.6 + 0.8 * torch.rand(n, generator=g)
                mod.bias.data = 0.2 * torch.randn(n, generator=g)
                mod.running_mean.data = 0.1 * torch.randn(n, generator=g)
                mod.running_var.data = 0.5 + torch.rand(n, generator=g)
    m.eval()
    if use_gpu:
        m.cuda()
    return m
