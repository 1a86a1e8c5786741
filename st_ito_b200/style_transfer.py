"""Inference-time optimisation host loop -- the call surface of the reference's
``st_ito/style_transfer.py`` for the ES path (load_plugins :17-42, process_audio :45-115,
parameters_to_dict :324-359, savepop_to_disk :362-396, run_es :399-692), with the population
evaluation moved onto the B200.

What changed relative to the reference, deliberately:
  * ``evaluate`` renders and scores the WHOLE population in one libstito call
    (stito_eval_population) instead of a Python loop over candidates + a batched encoder call;
    candidate audio is only copied back when it is needed (``savepop`` or a content model).
  * ``parallel=True`` is accepted and keeps the reference's length policy for that branch (no
    pad / crop, style_transfer.py:499-502) but spawns no process pool.
  * with torch.distributed initialised (one process per GPU) the population is sharded across
    ranks and the fitness values are all-gathered (st_ito_b200/dist.py).
  * ``seed=`` (swallowed by the reference's **kwargs) seeds the built-in CMA-ES.
Quirks kept on purpose: ``our_bypass`` never bypasses; in-place peak
normalisation of the caller's tensors; histories appended before ``tell``.
"""
from __future__ import annotations

import os
import time
from typing import List

import numpy as np
import torch

from . import dist as sdist
from .engine import compile_chain, fx_engine, plugins_are_native

try:  # the reference imports pycma; it is optional here
    import cma  # type: ignore
    if not hasattr(cma, "CMAEvolutionStrategy"):
        raise ImportError
except Exception:  # pragma: no cover - pycma is absent offline
    from . import cma  # noqa: F401

# ------- audio processing methods -------


def load_plugins(plugins: dict):
    """reference style_transfer.py:17-42"""
    total_num_params = 0
    init_params = []
    for plugin_name, plugin in plugins.items():
        if "vst_filepath" in plugin:
            import pedalboard  # VST3 hosting is not part of the B200 path; needs the real package

            plugin_instance = pedalboard.load_plugin(plugin["vst_filepath"])
        elif "class_path" in plugin:
            plugin_instance = plugin["class_path"]()
        else:
            raise ValueError("Plugin must contain 'vst_filepath' or 'class_path'.")

        plugin["parameter_names"] = ["our_bypass"]
        init_params.append(0.0)
        num_params = 1
        for name, parameter in plugin_instance.parameters.items():
            num_params += 1
            print(f"{plugin_name}: {name} = {parameter.raw_value}")
            init_params.append(parameter.raw_value)
            plugin["parameter_names"].append(name)
        print()

        plugin["num_params"] = num_params
        plugin["instance"] = plugin_instance
        total_num_params += num_params

    return plugins, total_num_params, init_params


def _assign_parameters(plugin: dict, w, widx: int) -> int:
    """The parameter loop of process_audio (style_transfer.py:76-92): mutates the plugin instance."""
    for name in plugin["parameter_names"]:
        if not name == "our_bypass":
            parameter = plugin["instance"].parameters[name]
            if name in plugin["fixed_parameters"]:
                if "vst_filepath" in plugin:
                    parameter.raw_value = plugin["fixed_parameters"][name]
                else:
                    parameter.set_value(plugin["fixed_parameters"][name])
                widx += 1
            else:
                parameter.raw_value = w[widx]
                widx += 1
        else:
            widx += 1  # the slot is consumed; the plugin still runs (reference :88-92)
    return widx


def _instantiate(plugin: dict):
    if "instance" not in plugin:
        if "vst_filepath" in plugin:
            import pedalboard

            plugin["instance"] = pedalboard.load_plugin(plugin["vst_filepath"])
        elif "class_path" in plugin:
            plugin["instance"] = plugin["class_path"]()
        else:
            raise ValueError("Plugin must contain 'vst_filepath' or 'class_path'.")


def process_audio(x: np.ndarray, w: np.ndarray, sr: int, plugins: List[dict], normalize_stages: bool = False):
    """Process audio with plugins and provided parameters on [0, 1] (reference :45-115).

    Args:
        x (np.ndarray): Audio of shape (chs, num_samples)
        w (np.ndarray): Parameter vector of shape (num_params,)
        sr (int): Sample rate
        plugins (dict): ordered plugin dicts
        normalize_stages (bool): Normalize the output of each stage
    """
    for plugin in plugins.values():
        _instantiate(plugin)
    if plugins_are_native(plugins):
        # whole chain in one device pass: EQ -> ... -> final peak normalisation
        desc, D = compile_chain(plugins, sr, normalize_stages)
        w = np.asarray(w, dtype=np.float64).reshape(-1)
        if w.shape[0] < D:
            raise IndexError(f"parameter vector has {w.shape[0]} entries, chain needs {D}")
        widx = 0
        for plugin in plugins.values():  # keep the reference's side effect on the plugin objects
            widx = _assign_parameters(plugin, w, widx)
        eng = fx_engine()
        eng.set_chain(desc)
        return eng.process(x, w[None, :D], final_normalize=True)[0]

    # generic plugins (user classes, VSTs): the reference's loop, plugin by plugin
    widx = 0
    for plugin in plugins.values():
        widx = _assign_parameters(plugin, w, widx)
        if plugin["num_channels"] == 2 and x.shape[0] == 1:
            x = np.concatenate((x, x), axis=0)
        if plugin["num_channels"] == 1 and x.shape[0] == 2:
            x_l = plugin["instance"].process(x[0:1, :], sample_rate=sr)
            x_r = plugin["instance"].process(x[1:2, :], sample_rate=sr)
            x = np.concatenate((x_l, x_r), axis=0)
        else:
            x = plugin["instance"].process(x, sample_rate=sr)
        if normalize_stages:
            x /= np.clip(np.max(np.abs(x)), a_min=1e-8, a_max=None)
    x /= np.clip(np.max(np.abs(x)), a_min=1e-8, a_max=None)
    return x


# ----------- Evolutionary Strategies ------------


def parameters_to_dict(w: np.ndarray, plugins: List[dict]):
    """Convert parameter vector to dictionary (reference :324-359)."""
    widx = 0
    w_dict = {}
    for plugin_name, plugin in plugins.items():
        if plugin_name not in w_dict:
            w_dict[plugin_name] = {}
        for name in plugin["parameter_names"]:
            if name == "our_bypass":
                w_dict[plugin_name][name] = w[widx]
                widx += 1
                continue
            parameter = plugin["instance"].parameters[name]
            if name in plugin["fixed_parameters"]:
                if "vst_filepath" in plugin:
                    parameter.raw_value = plugin["fixed_parameters"][name]
                else:
                    parameter.set_value(plugin["fixed_parameters"][name])
                widx += 1
            else:
                parameter.raw_value = w[widx]
                widx += 1
            if hasattr(parameter, "get_value"):
                w_dict[plugin_name][name] = parameter.get_value()
            else:
                w_dict[plugin_name][name] = parameter.raw_value
    return w_dict


def _save_wav(path: str, audio: torch.Tensor, sample_rate: int):
    """float32 WAV via scipy (torchaudio.save needs TorchCodec/soundfile, absent offline)."""
    from scipy.io import wavfile

    a = audio.detach().cpu().numpy().astype(np.float32)
    wavfile.write(path, int(sample_rate), np.ascontiguousarray(a.T))


def savepop_to_disk(iteration, fvals, output_embeds, output_audios, run_dir: str, sample_rate: int):
    """reference :362-396: one wav per population member, sorted by fitness."""
    pop_dir = os.path.join(run_dir, f"pop_{iteration}")
    os.makedirs(pop_dir, exist_ok=True)
    members = sorted(zip(fvals, output_audios), key=lambda m: m[0])
    for idx, (fval, output_audio) in enumerate(members):
        path = os.path.join(pop_dir, f"output_audio_pop_{idx}_fval_{fval:0.4e}.wav")
        output_audio = output_audio / torch.max(torch.abs(output_audio)).clamp(min=1e-8)
        _save_wav(path, output_audio, sample_rate)


def _is_fused(plugins, model, embed_func, content_model) -> bool:
    from .models.panns import Cnn14
    from .utils import get_param_embeds

    return (content_model is None and embed_func is get_param_embeds and isinstance(model, Cnn14)
            and plugins_are_native(plugins))


class FusedEvaluator:
    """``evaluate`` of the reference (style_transfer.py:474-573) for a chain of built-in plugins scored by the
    AFx-Rep encoder: ONE stito_eval_population call per population (per rank).

    Holds what is constant over a run -- the compiled chain, the target embeddings and the input waveform, all
    resident on the GPU -- and applies evaluate()'s length policy, the population sharding over
    torch.distributed ranks and the fitness all-gather.  ``run_es`` builds one; bench.py drives the same object.
    """

    crop_len = 262144

    def __init__(self, engine, plugins, sample_rate, target_embed, input_audio, random_crop=False, rng=np.random,
                 normalize_stages=False):
        self.engine = engine
        self.random_crop = random_crop
        self.rng = rng
        self.rank, self.world_size = sdist.world()
        desc, self.D = compile_chain(plugins, sample_rate, normalize_stages)
        engine.set_chain(desc)
        self.target_embed = target_embed
        engine.set_target_embeds(target_embed["mid"][0], target_embed["side"][0])
        x = input_audio[0] if input_audio.dim() == 3 else input_audio
        self.in_chs, self.x_len = int(x.shape[0]), int(x.shape[-1])
        engine.set_input(x, min_len=self.crop_len)
        # multi-GPU: the per-generation fitness all-gather runs over NVLink peer memory, fused into the fitness kernel
        # (stito_eval_population_gather); NCCL stays as the fall-back (and for embeddings / audio, which are not hot)
        self.peer_gather = False
        if self.world_size > 1 and sdist.backend() == "nccl" and os.environ.get("STITO_PEER_GATHER", "1") != "0":
            ready = getattr(engine, "_gather_ready", None)
            self.peer_gather = ready if ready is not None else engine.gather_setup(self.rank, self.world_size)

    def view_for(self, x_len: int, parallel: bool):
        """Length policy of evaluate (reference :499-518): (start, length) into the padded input."""
        if parallel:
            return 0, x_len
        if self.random_crop and (x_len - self.crop_len) > 16384:
            # the reference draws from numpy's global RNG; a seeded run draws from its own RandomState
            start_idx = int(self.rng.randint(16384, x_len - self.crop_len))
            if self.world_size > 1:  # one crop for the whole population, also across ranks
                start_idx = int(sdist.broadcast_array(np.array([float(start_idx)]))[0])
        else:
            start_idx = 0
        if x_len > self.crop_len:
            return (start_idx, self.crop_len) if self.random_crop else (0, x_len)
        return 0, self.crop_len

    def __call__(self, W, parallel=False, dropout=0.0, want_audio=False, want_embeds=False):
        """Returns (fvals list[P], {"mid","side"} [P, E] or None, audio [P, chs', len] or None)."""
        W = np.asarray(W, dtype=np.float64)
        P = W.shape[0]
        start, length = self.view_for(self.x_len, parallel)
        want_embeds = want_embeds or dropout > 0.0
        if self.world_size == 1:
            fit, emb, aud = self.engine.eval_population(W, start, length, want_embeds=want_embeds,
                                                        want_audio=want_audio, in_chs=self.in_chs)
        else:
            # rank r scores the slice [lo, hi) -- possibly empty when there are more ranks than candidates -- and
            # the rows are all-gathered; on NCCL the results stay on the GPU until after the collective
            lo, hi, _ = sdist.shard_bounds(P, self.world_size, self.rank)
            if self.peer_gather and not want_embeds and not want_audio and P <= 4096:
                fit = self.engine.eval_population_gather(W[lo:hi], start, length, lo, P)
                return fit.tolist(), None, None
            on_gpu = sdist.backend() == "nccl"
            fit, emb, aud = self.engine.eval_population(W[lo:hi], start, length, want_embeds=want_embeds,
                                                        want_audio=want_audio, in_chs=self.in_chs,
                                                        device_out=on_gpu)
            fit = sdist.all_gather_rows(fit, P).cpu()
            if emb is not None:
                emb = sdist.all_gather_rows(emb.transpose(0, 1).contiguous(), P).cpu().transpose(0, 1)
            if aud is not None:
                aud = sdist.all_gather_rows(aud, P).cpu()
        output_embeds = {"mid": emb[0], "side": emb[1]} if emb is not None else None
        if dropout > 0.0:  # stochastic regulariser of the reference (:550-551), on the tiny embeddings
            dists = []
            for name in ("mid", "side"):
                oe = torch.nn.functional.dropout(output_embeds[name], p=dropout)
                dists.append(-torch.cosine_similarity(oe, self.target_embed[name].cpu().float(), dim=-1))
            fit = torch.stack(dists, dim=0).mean(dim=0)
            if self.world_size > 1:
                # every rank drew its own mask; CMA-ES must see ONE fitness vector or the ranks' populations diverge
                fit = torch.from_numpy(sdist.broadcast_array(fit.double().numpy())).float()
        return fit.tolist(), output_embeds, aud


def run_es(
    input_audio: torch.Tensor,
    target_audio: torch.Tensor,
    sample_rate: int,
    plugins: List[dict],
    model: torch.nn.Module,
    embed_func: callable,
    content_model: torch.nn.Module = None,
    content_embed_func: callable = None,
    max_iters: int = 100,
    w0: torch.Tensor = None,
    find_w0: bool = True,
    sigma0: float = 0.1,
    distance: str = "cosine",
    random_crop: bool = False,
    popsize: int = 32,
    parallel: bool = False,
    dropout: float = 0.0,
    savepop: bool = False,
    run_dir: str = ".",
    *args,
    **kwargs,
):
    """Run CMA-ES optimization to find the best parameters (reference :399-692).

    Same arguments and result dict as the reference.  Extra keyword arguments the reference swallows
    in **kwargs: ``seed`` (CMA-ES / find_w0 / random_crop RNG), ``verbose`` and ``iteration_times`` (a list that
    receives the wall-clock seconds of every ask -> evaluate -> tell generation) are honoured.  ``normalize_stages``
    is swallowed exactly as in the reference, whose evaluate() and final render call process_audio
    without it (style_transfer.py:519-521, 676-678), so run_optim's --normalize-stages has no effect
    on the ES path there either.
    """
    seed = kwargs.get("seed", None)
    verbose = kwargs.get("verbose", True)
    normalize_stages = False
    rng = np.random.RandomState(seed) if seed is not None else np.random

    def log(*a):
        if verbose:
            print(*a)

    total_num_params = sum([plugin["num_params"] for plugin in plugins.values()])
    bs, chs, seq_len = input_audio.shape

    # peak normalize (in place, like the reference :452-453)
    input_audio /= torch.max(torch.abs(input_audio)).clamp(min=1e-8)
    target_audio /= torch.max(torch.abs(target_audio)).clamp(min=1e-8)

    # compute target embedding (only once)
    target_embed = embed_func(target_audio, model, sample_rate)

    if content_model is not None:
        target_content_embeds = content_embed_func(target_audio, content_model, sample_rate)
    else:
        target_content_embeds = None

    for plugin in plugins.values():
        _instantiate(plugin)
    fused = _is_fused(plugins, model, embed_func, content_model) and sample_rate == 48000
    rank, world_size = sdist.world()
    crop_len = 262144

    fused_eval = None
    if fused:
        if compile_chain(plugins, sample_rate, normalize_stages)[1] != total_num_params:
            raise ValueError(f"plugins declare {total_num_params} parameters but the chain walk consumes a different count")
        fused_eval = FusedEvaluator(model.stito_engine(), plugins, sample_rate, target_embed, input_audio,
                                    random_crop=random_crop, rng=rng, normalize_stages=normalize_stages)

    def view_for(x_len: int, parallel: bool):
        """Length policy of evaluate (reference :499-518) for the generic path."""
        if parallel:
            return 0, x_len
        if random_crop and (x_len - crop_len) > 16384:
            start_idx = int(rng.randint(16384, x_len - crop_len))
        else:
            start_idx = 0
        if x_len > crop_len:
            return (start_idx, crop_len) if random_crop else (0, x_len)
        return 0, crop_len

    def evaluate(W, x, sample_rate, plugins, target_embeds, target_content_embeds=None, parallel=False,
                 dropout=0.0):
        """Evaluate the current population (reference :474-573).

        Returns (fvals list[P], output_embeds {"mid","side"} [P, E], output_audios [P, chs', L] or None).
        """
        W = np.asarray(W, dtype=np.float64)
        want_audio = savepop or content_model is not None
        if fused:
            return fused_eval(W, parallel=parallel, dropout=dropout, want_audio=want_audio, want_embeds=savepop)

        # generic path: arbitrary plugins / embedding functions, candidate by candidate (reference loop)
        output_audios = []
        if parallel:
            xs = x
        else:
            start, length = view_for(x.shape[-1], False)
            if x.shape[-1] > crop_len:
                xs = x[:, :, start:start + length]
            else:
                xs = torch.nn.functional.pad(x, (0, crop_len - x.shape[-1]))
        for w in W:
            output_audios.append(torch.from_numpy(
                process_audio(xs.squeeze(0).numpy(), w, sample_rate, plugins, normalize_stages)))
        output_audios = torch.stack(output_audios, dim=0)
        output_embeds = embed_func(output_audios, model, sample_rate)
        if content_model is not None:
            output_content_embeds = content_embed_func(output_audios, content_model, sample_rate)
        dists = []
        for embed_name, output_embed in output_embeds.items():
            tgt = target_embeds[embed_name]
            if dropout > 0.0:
                output_embed = torch.nn.functional.dropout(output_embed, p=dropout)
            with torch.no_grad():
                dists.append(-torch.cosine_similarity(output_embed, tgt, dim=-1))
        if target_content_embeds is not None:
            for embed_name, oce in output_content_embeds.items():
                d = -torch.cosine_similarity(oce, target_content_embeds[embed_name], dim=-1)
                dists.append(2 * d)
        dist_ = torch.stack(dists, dim=0).mean(dim=0)
        return dist_.tolist(), output_embeds, output_audios

    def replicated(W):
        """All ranks must evaluate the same population: broadcast unless a shared seed makes it so."""
        W = np.asarray(W, dtype=np.float64)
        if world_size > 1 and seed is None:
            W = sdist.broadcast_array(W)
        return W

    # setup CMA-ES
    if find_w0:
        log("Finding the best w0...")
        tmp_w0s = replicated(np.stack([rng.rand(total_num_params) for _ in range(popsize)]))
        fvals, output_embeds, output_audios = evaluate(
            tmp_w0s, input_audio, sample_rate, plugins, target_embed,
            target_content_embeds=target_content_embeds, parallel=parallel, dropout=dropout)
        log(fvals)
        w0 = tmp_w0s[int(np.argmin(fvals))]
        if savepop and rank == 0:
            savepop_to_disk(-1, fvals, output_embeds, output_audios, run_dir, sample_rate)
    else:
        if w0 is None:
            w0 = np.ones(total_num_params) * 0.5
        else:
            w0 = w0.numpy() if isinstance(w0, torch.Tensor) else np.asarray(w0)

    init_param_dict = parameters_to_dict(w0, plugins)
    log(init_param_dict)

    opts = {"bounds": [0, 1], "popsize": popsize}
    if seed is not None:
        opts["seed"] = int(seed) + 1
    if not verbose:
        opts["verbose"] = -9
    es = cma.CMAEvolutionStrategy(w0, sigma0, opts)

    fval_history = []
    wopt_history = []
    iters_without_improvement = 0
    x = input_audio

    iteration_times = kwargs.get("iteration_times", None)  # optional list: wall seconds of every ask -> tell

    for iteration in range(max_iters):
        t_iter = time.perf_counter()
        x = input_audio.clone() if not fused else input_audio  # the fused path never mutates x
        W = replicated(es.ask())
        fvals, output_embeds, output_audios = evaluate(
            W, x, sample_rate, plugins, target_embed, target_content_embeds=target_content_embeds,
            parallel=parallel,
            dropout=(dropout if (iteration + 1) < max_iters else 0.0))  # no dropout on the last iteration

        # save best (before tell, like the reference :639-640)
        wopt_history.append(es.result[0])
        fval_history.append(es.result[1])

        if savepop and rank == 0:
            savepop_to_disk(iteration, fvals, output_embeds, output_audios, run_dir, sample_rate)
        es.tell(list(W), fvals)
        if iteration_times is not None:
            iteration_times.append(time.perf_counter() - t_iter)
        if verbose:
            es.disp()

        if iteration > 0:
            fval_delta = min(fvals) - min(fval_history)
        else:
            fval_delta = -0.02

        if fval_delta > -0.01:
            iters_without_improvement += 1
            log(f"Solution has not improved for {iters_without_improvement} iterations.")
        else:
            iters_without_improvement = 0

        if iters_without_improvement > 10:
            log("Stopping early due to no improvement.")
            break

    wopt = es.result[0]
    fopt = es.result[1]

    output_audio = torch.from_numpy(
        process_audio(x.squeeze(0).numpy(), wopt, sample_rate, plugins, normalize_stages))
    param_dict = parameters_to_dict(wopt, plugins)

    return {
        "output_audio": output_audio,
        "params": param_dict,
        "fopt": fopt,
        "wopt": wopt,
        "fval_history": fval_history,
        "wopt_history": wopt_history,
    }
