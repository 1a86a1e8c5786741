"""Host-side wrapper around a libstito handle: chain compilation (plugins dict -> flat descriptor),
weight upload, and the population-evaluation / render / embed calls.  PyTorch is used here only
for tensors, streams and device selection; the arithmetic is in st_ito_b200/csrc.
"""
from __future__ import annotations

import os
from ctypes import byref, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import ChainDesc, EncoderWeights, Timing, check, ptr


def default_device() -> int:
    """One process per GPU: LOCAL_RANK selects the device under torchrun, else torch's current device."""
    if not torch.cuda.is_available():
        raise RuntimeError("st_ito_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count()
    return torch.cuda.current_device()


_CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the handle value 0 means "the library's own stream" in stito.h


def _stream_ptr(device: int):
    """torch's current stream as a cudaStream_t.  torch's default stream is the legacy default stream (handle 0);
    libstito reads NULL as "use my own non-blocking stream", which is not ordered against work torch has queued
    (e.g. get_param_embeds' in-place peak normalisation of a CUDA tensor), so 0 is passed as cudaStreamLegacy."""
    handle = int(torch.cuda.current_stream(device).cuda_stream)
    return c_void_p(handle if handle != 0 else _CUDA_STREAM_LEGACY)


def _wait_for_torch(t):
    """Setup calls (set_input / set_target) run on the handle's own stream: a CUDA tensor argument may still be
    being written by kernels queued on torch's stream."""
    if isinstance(t, torch.Tensor) and t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()


# ------------------------------------------------------------------------------------------------
# chain compilation
# ------------------------------------------------------------------------------------------------
def plugins_are_native(plugins: dict) -> bool:
    from .effects import is_native_plugin

    for entry in plugins.values():
        inst = entry.get("instance")
        if inst is None:
            cls = entry.get("class_path")
            if cls is None or not isinstance(cls, type):
                return False
            inst = cls()
        if not is_native_plugin(inst):
            return False
    return True


def compile_chain(plugins: dict, sample_rate: float, normalize_stages: bool = False):
    """Flatten the reference's ordered ``plugins`` dict (style_transfer.py:17-42, 76-92) into a
    stito_chain_desc.  Returns (desc, D).  The walk over ``parameter_names`` reproduces
    process_audio's index bookkeeping: every name, including ``our_bypass`` and fixed parameters,
    consumes one slot of ``w``; ``our_bypass`` maps to nothing (the reference never skips a plugin).
    """
    from .effects import is_native_plugin

    if len(plugins) > _lib.MAX_FX:
        raise ValueError(f"at most {_lib.MAX_FX} plugins per chain")
    desc = ChainDesc()
    desc.num_fx = len(plugins)
    desc.normalize_stages = int(bool(normalize_stages))
    desc.sample_rate = float(sample_rate)
    widx = 0
    for f, (plugin_name, plugin) in enumerate(plugins.items()):
        if "instance" not in plugin:
            if "vst_filepath" in plugin:
                raise ValueError(f"{plugin_name}: VST3 hosting is outside the B200 path (built-in plugins only)")
            elif "class_path" in plugin:
                plugin["instance"] = plugin["class_path"]()
            else:
                raise ValueError("Plugin must contain 'vst_filepath' or 'class_path'.")
        inst = plugin["instance"]
        if not is_native_plugin(inst):
            raise ValueError(f"{plugin_name}: {type(inst).__name__} is not one of the built-in Basic* plugins")
        if "parameter_names" not in plugin:
            plugin["parameter_names"] = list(inst.parameters.keys())
        fx = desc.fx[f]
        fx.kind = inst.stito_kind
        fx.num_channels = int(plugin["num_channels"])
        for k, v in enumerate(getattr(inst, "stito_iopt", (0, 0, 0, 0))):
            fx.iopt[k] = int(v)
        names = [s[0] for s in inst._spec]
        fx.num_params = len(names)
        for k in range(_lib.MAX_FX_PARAMS):
            fx.w_index[k] = -1
        # parameters never named in parameter_names keep the instance's current raw_value
        for k, n in enumerate(names):
            fx.fixed_raw[k] = float(inst.parameters[n].raw_value)
        fixed = plugin.get("fixed_parameters", {})
        for name in plugin["parameter_names"]:
            if name != "our_bypass":
                k = names.index(name)
                if name in fixed:
                    inst.parameters[name].set_value(fixed[name])  # asserts the range like the reference
                    fx.fixed_raw[k] = float(inst.parameters[name].raw_value)
                    fx.w_index[k] = -1
                else:
                    fx.w_index[k] = widx
            widx += 1
    desc.num_w = widx
    return desc, widx


def _single_plugin_chain(plugin, chs: int, sample_rate: float):
    desc = ChainDesc()
    desc.num_fx = 1
    desc.num_w = 0
    desc.sample_rate = float(sample_rate)
    fx = desc.fx[0]
    fx.kind = plugin.stito_kind
    fx.num_channels = 2 if plugin.stito_kind == _lib.FX_CONV_REVERB else chs  # always a stereo effect (mono is up-mixed)
    for k, v in enumerate(getattr(plugin, "stito_iopt", (0, 0, 0, 0))):
        fx.iopt[k] = int(v)
    fx.num_params = len(plugin._spec)
    for k, (n, *_rest) in enumerate(plugin._spec):
        fx.w_index[k] = -1
        fx.fixed_raw[k] = float(plugin.parameters[n].raw_value)
    return desc


# ------------------------------------------------------------------------------------------------
# handle wrapper
# ------------------------------------------------------------------------------------------------
def _np32(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())


class Engine:
    """One libstito handle = (effect chain, encoder weights) on one B200."""

    def __init__(self, model=None, device=None, chain: ChainDesc = None):
        self.device = default_device() if device is None else int(device)
        self.embed_dim = 512
        self.hop, self.n_mels = 1024, 128
        self._h = c_void_p()
        self._input_key = None
        if chain is None:
            chain = ChainDesc()
            chain.sample_rate = 48000.0
        weights = None
        keep = []
        if model is not None:
            sd = model.state_dict()
            weights = EncoderWeights()
            weights.n_fft, weights.hop, weights.n_mels = model.window_size, model.hop_size, model.mel_bins
            weights.embed_dim = model.embed_dim
            weights.bn_eps = float(model.conv_block1.bn1.eps)
            self.embed_dim = model.embed_dim
            self.hop, self.n_mels = int(model.hop_size), int(model.mel_bins)

            def put(field, idx, key):
                a = _np32(sd[key])
                keep.append(a)
                p = a.ctypes.data_as(_lib._f32p)
                if idx is None:
                    setattr(weights, field, p)
                else:
                    getattr(weights, field)[idx] = p

            for b in range(6):
                for j in (1, 2):
                    li = 2 * b + (j - 1)
                    pre = f"conv_block{b + 1}"
                    put("conv_w", li, f"{pre}.conv{j}.weight")
                    put("bn_weight", li, f"{pre}.bn{j}.weight")
                    put("bn_bias", li, f"{pre}.bn{j}.bias")
                    put("bn_mean", li, f"{pre}.bn{j}.running_mean")
                    put("bn_var", li, f"{pre}.bn{j}.running_var")
            put("fc_mid_w", None, "fc_mid.weight")
            put("fc_mid_b", None, "fc_mid.bias")
            put("fc_side_w", None, "fc_side.weight")
            put("fc_side_b", None, "fc_side.bias")
            put("mel_w", None, "logmel_extractor.melW")
        check(_lib.lib().stito_create(byref(chain), byref(weights) if weights is not None else None,
                                      self.device, byref(self._h)))
        del keep

    # -- lifetime ----------------------------------------------------------------------------
    def close(self):
        if self._h:
            _lib.lib().stito_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration -----------------------------------------------------------------------
    def set_chain(self, chain: ChainDesc):
        check(_lib.lib().stito_set_chain(self._h, byref(chain)))

    def set_precision(self, precision: int):
        check(_lib.lib().stito_set_precision(self._h, int(precision)))

    def set_input(self, x, min_len: int = 0):
        """x: [chs, L] float32 (numpy or torch, host or device)."""
        x = self._as_f32(x)
        _wait_for_torch(x)
        chs, L = x.shape
        check(_lib.lib().stito_set_input(self._h, ptr(x), chs, L, int(min_len)))
        return max(L, int(min_len))

    def set_target(self, target):
        t = self._as_f32(target)
        _wait_for_torch(t)
        chs, L = t.shape
        check(_lib.lib().stito_set_target(self._h, ptr(t), chs, L))

    def set_target_embeds(self, mid, side):
        mid = self._as_f32(mid).reshape(-1)
        side = self._as_f32(side).reshape(-1)
        _wait_for_torch(mid)
        _wait_for_torch(side)
        check(_lib.lib().stito_set_target_embeds(self._h, ptr(mid), ptr(side), int(mid.shape[0])))

    def out_channels(self, chs: int) -> int:
        return check(_lib.lib().stito_out_channels(self._h, int(chs)))

    # -- the hot path ------------------------------------------------------------------------
    def eval_population(self, W, start: int, length: int, want_embeds: bool = False, want_audio: bool = False,
                        in_chs: int = None, device_out: bool = False):
        """evaluate() of style_transfer.py:474-573 for the population W [P, D] (float64).

        Returns (fitness float32[P], embeds float32[2, P, E] or None, audio float32[P, chs', len] or None) as
        pinned host torch tensors, or -- with ``device_out`` -- as tensors on this engine's GPU (what the NCCL
        all-gather of a sharded population consumes: no host bounce).  P == 0 (an empty shard) returns empty
        tensors.
        """
        W = np.ascontiguousarray(np.asarray(W, dtype=np.float64))
        if W.ndim != 2:
            raise ValueError("W must be [P, D]")
        P, D = W.shape
        ochs = self.out_channels(in_chs if in_chs is not None else 2) if want_audio else 0
        if device_out:
            dev = torch.device("cuda", self.device)
            fit = torch.empty((P,), dtype=torch.float32, device=dev)
            emb = torch.empty((2, P, self.embed_dim), dtype=torch.float32, device=dev) if want_embeds else None
            aud = torch.empty((P, ochs, length), dtype=torch.float32, device=dev) if want_audio else None
        else:
            # pinned result buffers are cached per shape (cudaHostAlloc per generation is measurable next to a
            # 15 ms evaluation); results are cloned out so that callers own what they get
            fit = self._pinned("fit", (P,))
            emb = self._pinned("emb", (2, P, self.embed_dim)) if want_embeds else None
            aud = torch.empty((P, ochs, length), dtype=torch.float32, pin_memory=True) if want_audio else None
        check(_lib.lib().stito_eval_population(self._h, ptr(W), P, D, int(start), int(length), ptr(fit), ptr(emb),
                                               ptr(aud), _stream_ptr(self.device)))
        if device_out:
            return fit, emb, aud
        return fit.clone(), (emb.clone() if emb is not None else None), aud

    # -- multi-GPU: fitness all-gather over peer memory, fused into the fitness kernel ---------------
    def gather_setup(self, rank: int, world: int, capacity: int = 4096) -> bool:
        """Export this handle's gather block, exchange the CUDA IPC handles over torch.distributed and attach the peers'
        blocks (stito_gather_export / _attach).  Collective.  Returns False (on every rank) if any rank could not set it
        up -- e.g. peer access between the GPUs is not available -- in which case callers keep the NCCL all-gather."""
        import ctypes

        import torch.distributed as dist

        ok, mine = 1, bytes(64)
        try:
            buf = ctypes.create_string_buffer(64)
            check(_lib.lib().stito_gather_export(self._h, int(capacity), buf))
            mine = buf.raw
        except Exception:
            ok = 0
        handles = [None] * world
        dist.all_gather_object(handles, (ok, mine))
        if all(h[0] for h in handles):
            try:
                blob = b"".join(h[1] for h in handles)
                check(_lib.lib().stito_gather_attach(self._h, int(rank), int(world), blob))
            except Exception:
                ok = 0
        else:
            ok = 0
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        self._gather_ready = all(flags)
        return self._gather_ready

    def eval_population_gather(self, W, start: int, length: int, lo: int, P_total: int):
        """Score the shard W = candidates [lo, lo + len(W)) of a population of P_total and return ALL P_total fitness values
        (pinned host tensor): stito_eval_population_gather.  Collective over the ranks of gather_setup()."""
        W = np.ascontiguousarray(np.asarray(W, dtype=np.float64))
        P, D = (W.shape if W.ndim == 2 else (0, 0))
        fit = self._pinned("fit_all", (int(P_total),))
        check(_lib.lib().stito_eval_population_gather(self._h, ptr(W) if P > 0 else None, int(P), int(D), int(start), int(length),
                                                      int(lo), int(P_total), ptr(fit), _stream_ptr(self.device)))
        return fit.clone()

    def _pinned(self, key, shape):
        cache = self.__dict__.setdefault("_pinned_cache", {})
        k = (key, tuple(shape))
        if k not in cache:
            n = int(np.prod(shape))
            # pin_memory of a zero-element tensor is not portable across torch versions: an empty shard needs no pinning
            cache[k] = torch.empty(shape, dtype=torch.float32, pin_memory=n > 0)
        return cache[k]

    def process(self, x, W, final_normalize: bool = True) -> np.ndarray:
        """process_audio for P parameter vectors: x [chs, L] -> [P, chs', L] float32 (numpy)."""
        x = self._as_f32(x)
        chs, L = x.shape
        W = np.ascontiguousarray(np.asarray(W, dtype=np.float64))
        if W.ndim == 1:
            W = W[None]
        P, D = W.shape
        y = np.empty((P, self.out_channels(chs), L), dtype=np.float32)
        check(_lib.lib().stito_process(self._h, ptr(x), chs, L, ptr(W) if D > 0 else None, P, D,
                                       int(bool(final_normalize)), ptr(y), _stream_ptr(self.device)))
        return y

    def embed(self, x: torch.Tensor, peak_normalize: bool = False):
        """Cnn14.forward on x [B, chs, L]; returns raw (mid, side) [B, E] on x's device / dtype."""
        xf = x.detach()
        if xf.dtype != torch.float32 or not xf.is_contiguous():
            xf = xf.to(torch.float32).contiguous()
        if xf.is_cuda and xf.device.index != self.device:
            xf = xf.to(f"cuda:{self.device}")
        B, chs, L = xf.shape
        dev = xf.device
        mid = torch.empty((B, self.embed_dim), dtype=torch.float32, device=dev)
        side = torch.empty((B, self.embed_dim), dtype=torch.float32, device=dev)
        check(_lib.lib().stito_embed(self._h, ptr(xf), B, chs, L, int(bool(peak_normalize)), ptr(mid), ptr(side),
                                     _stream_ptr(self.device)))
        return mid.to(x.device, x.dtype), side.to(x.device, x.dtype)

    def logmel(self, x: torch.Tensor) -> torch.Tensor:
        xf = x.detach().to(torch.float32).contiguous()
        B, chs, L = xf.shape
        T = L // self.hop + 1  # the library's own frame count (stito_logmel: T = L / hop + 1)
        out = torch.empty((B * chs, T, self.n_mels), dtype=torch.float32, device=xf.device)
        check(_lib.lib().stito_logmel(self._h, ptr(xf), B, chs, L, ptr(out), _stream_ptr(self.device)))
        return out

    def timing(self) -> dict:
        t = Timing()
        check(_lib.lib().stito_get_timing(self._h, byref(t)))
        d = {k: getattr(t, k) for k, _ in Timing._fields_ if k != "ms_conv"}
        d["ms_conv"] = [float(v) for v in t.ms_conv]
        return d

    @staticmethod
    def _as_f32(a):
        if isinstance(a, torch.Tensor):
            a = a.detach()
            if a.dtype != torch.float32 or not a.is_contiguous():
                a = a.to(torch.float32).contiguous()
            return a
        return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# ------------------------------------------------------------------------------------------------
# effects-only engine shared by plugin.process / process_audio
# ------------------------------------------------------------------------------------------------
_fx_engines = {}


def fx_engine(device=None) -> Engine:
    dev = default_device() if device is None else int(device)
    if dev not in _fx_engines:
        _fx_engines[dev] = Engine(model=None, device=dev)
    return _fx_engines[dev]


def render_single_plugin(plugin, x: np.ndarray, sample_rate: float) -> np.ndarray:
    """plugin.process(x, sample_rate): one effect, current raw_values, no final normalisation."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    if x.ndim != 2 or x.shape[0] not in (1, 2):
        raise ValueError("audio must be [chs, L] with 1 or 2 channels")
    eng = fx_engine()
    eng.set_chain(_single_plugin_chain(plugin, x.shape[0], sample_rate))
    return eng.process(x, np.zeros((1, 0)), final_normalize=False)[0]
