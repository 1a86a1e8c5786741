#!/usr/bin/env python
"""ES inference-time optimisation driver -- same command line as the reference's
``scripts/run_optim.py`` (:300-322), with the population evaluation on the B200.

    python scripts/run_optim.py input.wav target.wav --effect-type basic --max-iters 25 --popsize 64

Kept from the reference: positional ``input`` / ``target``; ``--max-iters --popsize --max-length --staged
--savepop --normalize-stages --use-gpu --parallel --effect-type --algorithm --dropout --metric``; the output
layout ``output/optim/<input>_to_<target>_<algorithm>/{input_audio,target_audio,output_audio_sigma=0.33}.wav``,
``parameters_sigma=0.33.json`` and (if matplotlib is installed) ``plot.png``; sigma0 = 0.33, find_w0=True.

Additive flags (defaults reproduce the reference): ``--chain`` picks a built-in chain preset
(``basic`` = the reference's EQ->Comp->Dist->Delay->Reverb literal at run_optim.py:376-407; ``eq`` and
``mastering-pb`` are BASELINE configs 1 and 2), ``--ckpt`` points at afx-rep.ckpt, ``--synthetic-weights``
uses seeded random AFx-Rep weights (no checkpoint can be downloaded offline), ``--seed`` seeds the CMA-ES.

Not on this path (errors, like an unknown algorithm in the reference): ``--effect-type vst`` (VST3 hosting via
pedalboard), ``--algorithm autodiff`` (dasp-pytorch), ``--metric clap``, ``--staged`` (dead code upstream,
SURVEY Appendix C.3).  Under ``torchrun`` the population is sharded over the ranks (one process per GPU).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_wav(path: str):
    """[chs, L] float32 in [-1, 1] + sample rate (scipy: torchaudio.load needs TorchCodec offline)."""
    from scipy.io import wavfile

    sr, a = wavfile.read(path)
    if a.dtype == np.int16:
        a = a.astype(np.float32) / 32768.0
    elif a.dtype == np.int32:
        a = a.astype(np.float32) / 2147483648.0
    elif a.dtype == np.uint8:
        a = (a.astype(np.float32) - 128.0) / 128.0
    else:
        a = a.astype(np.float32)
    if a.ndim == 1:
        a = a[:, None]
    return torch.from_numpy(np.ascontiguousarray(a.T)), int(sr)


def save_wav(path: str, audio: torch.Tensor, sr: int):
    from scipy.io import wavfile

    wavfile.write(path, int(sr), np.ascontiguousarray(audio.detach().cpu().numpy().astype(np.float32).T))


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("input", type=str)
    parser.add_argument("target", type=str)
    parser.add_argument("--max-iters", type=int, default=300)
    parser.add_argument("--popsize", type=int, default=32)
    parser.add_argument("--max-length", type=int, default=262144)
    parser.add_argument("--staged", action="store_true")
    parser.add_argument("--savepop", action="store_true")
    parser.add_argument("--normalize-stages", action="store_true")
    parser.add_argument("--use-gpu", action="store_true")
    parser.add_argument("--parallel", action="store_true")
    parser.add_argument("--effect-type", type=str, default="vst", choices=["vst", "basic"])
    parser.add_argument("--algorithm", type=str, default="es", choices=["es", "autodiff"])
    parser.add_argument("--dropout", type=float, default=0.0)
    parser.add_argument("--metric", type=str, default="param", choices=["param", "clap"])
    # additive
    parser.add_argument("--chain", type=str, default=None, choices=["basic", "eq", "mastering-pb"])
    parser.add_argument("--ckpt", type=str, default=None)
    parser.add_argument("--synthetic-weights", action="store_true")
    parser.add_argument("--seed", type=int, default=None)
    parser.add_argument("--output-dir", type=str, default=os.path.join("output", "optim"))
    args = parser.parse_args(argv)

    from st_ito_b200 import dist as sdist
    from st_ito_b200 import effects
    from st_ito_b200.style_transfer import run_es
    from st_ito_b200.utils import get_param_embeds, load_param_model, make_synthetic_param_model

    sample_rate = 48000
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    rank, _ = sdist.world()

    if args.algorithm != "es":
        raise ValueError(f"Unknown algorithm: {args.algorithm} (only the ES path is implemented on the B200)")
    if args.staged:
        raise ValueError("--staged is dead code in the reference (SURVEY Appendix C.3) and not implemented")
    chain = args.chain if args.chain is not None else ("basic" if args.effect_type == "basic" else None)
    if chain is None:
        raise ValueError("--effect-type vst needs pedalboard VST3 hosting, which is outside the B200 path; "
                         "use --effect-type basic (or --chain)")
    plugins = effects.make_chain(chain)

    # the reference's own loader (run_optim.py:410-437): NO our_bypass slot, w0 = current raw values
    total_num_params, init_params = 0, []
    for plugin_name, plugin in plugins.items():
        inst = plugin["class_path"]()
        num_params = 0
        for name, parameter in inst.parameters.items():
            num_params += 1
            if rank == 0:
                print(f"{plugin_name}: {name} = {parameter.raw_value}")
            init_params.append(parameter.raw_value)
        plugin["num_params"] = num_params
        plugin["instance"] = inst
        plugin["parameter_names"] = [name for name, _ in inst.parameters.items()]
        total_num_params += num_params
    w0 = torch.zeros(total_num_params)
    for idx, param in enumerate(init_params):
        w0[idx] = param

    input_audio, input_sr = load_wav(args.input)
    target_audio, target_sr = load_wav(args.target)
    if input_sr != sample_rate or target_sr != sample_rate:
        import torchaudio

        if input_sr != sample_rate:
            input_audio = torchaudio.functional.resample(input_audio, input_sr, sample_rate)
        if target_sr != sample_rate:
            target_audio = torchaudio.functional.resample(target_audio, target_sr, sample_rate)
    input_name = os.path.basename(args.input).replace(".wav", "")
    target_name = os.path.basename(args.target).replace(".wav", "")

    # crop to max length
    input_audio = input_audio[:, : args.max_length].contiguous()
    target_audio = target_audio[:, : args.max_length].contiguous()

    run_name = f"{input_name}_to_{target_name}_{args.algorithm}"
    run_dir = os.path.join(args.output_dir, run_name)
    os.makedirs(run_dir, exist_ok=True)

    if args.metric != "param":
        raise ValueError(f"Unknown metric: {args.metric} (only the AFx-Rep 'param' metric is on the B200 path)")
    if args.synthetic_weights:
        model = make_synthetic_param_model(seed=3, conv_gain=2.0, use_gpu=args.use_gpu)
    else:
        model = load_param_model(args.ckpt, use_gpu=args.use_gpu)
    embed_func = get_param_embeds

    if rank == 0:
        save_wav(os.path.join(run_dir, "input_audio.wav"), input_audio, sample_rate)
    target_audio /= torch.max(torch.abs(target_audio)).clamp(min=1e-8)
    if rank == 0:
        save_wav(os.path.join(run_dir, "target_audio.wav"), target_audio, sample_rate)

    input_audio = input_audio.unsqueeze(0)
    target_audio = target_audio.unsqueeze(0)
    sigma0 = 0.33
    if rank == 0:
        print(f"Running ES with sigma0 = {sigma0}")
    result = run_es(input_audio, target_audio, sample_rate, plugins, model, embed_func,
                    max_iters=args.max_iters, popsize=args.popsize, w0=w0, find_w0=True, sigma0=sigma0,
                    distance="cosine", parallel=args.parallel, dropout=args.dropout, savepop=args.savepop,
                    normalize_stages=args.normalize_stages, run_dir=run_dir, seed=args.seed, verbose=rank == 0)
    if rank != 0:
        return result

    output_audio = result["output_audio"]
    fval_history = result["fval_history"]
    try:  # plotting is optional (matplotlib is not part of the offline image)
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt

        fig, axs = plt.subplots(1, 2, figsize=(10, 5))
        axs[0].plot(fval_history, label=f"sigma0={sigma0:0.2f}")
        axs[0].set_xlabel("Iteration")
        axs[0].set_ylabel("Distance")
        axs[0].legend()
        plt.savefig(os.path.join(run_dir, "plot.png"), dpi=300)
    except ImportError:
        pass

    output_audio = output_audio / torch.max(torch.abs(output_audio)).clamp(min=1e-8)
    save_wav(os.path.join(run_dir, f"output_audio_sigma={sigma0:0.2f}.wav"), output_audio.squeeze(0), sample_rate)
    with open(os.path.join(run_dir, f"parameters_sigma={sigma0:0.2f}.json"), "w") as f:
        json.dump(result["params"], f, indent=4)
    print(f"fopt = {result['fopt']:.6f}; results in {run_dir}")
    return result


if __name__ == "__main__":
    main()
