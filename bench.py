#!/usr/bin/env python
"""bench.py -- ES candidates/sec of the population-evaluation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pop P] [--seconds S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one CMA-ES generation: evaluate(W) for a population of P candidates: render
each candidate through the effect chain (EQ -> Compressor -> Reverb), log-mel, AFx-Rep encoder, cosine
fitness against the target's mid/side embeddings.  Workload = BASELINE config 2 (10 s stereo 48 kHz,
P = 64, "mastering-pb" chain) per GPU; with N GPUs every rank evaluates its own P candidates (weak
scaling, the population is sharded) and the fitness values are all-gathered over NCCL each generation.

value  : candidates/s with the input waveform resident in HBM (the per-generation parameter block W,
         P*D doubles, is the only H2D traffic inside the timed region).
e2e    : candidates/s through the host-buffer API -- every step uploads the input waveform and W from
         pinned host memory and reads fitness + embeddings back.
The reference arm (--impl reference) and cpu_baseline time the CPU oracle (a restatement of the
reference's CPU path: the reference itself is pure Python over scipy/pedalboard/torch and cannot be
installed offline) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 48000
METRIC = "ES candidates/sec (pop=64, 10s@48kHz stereo)"
UNIT = "candidates/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(seconds, chs=2):
    from tests.signals import test_signal

    L = int(seconds * SR)
    x = test_signal(chs, L, seed=0)
    return x / np.abs(x).max()


# ------------------------------------------------------------------------------------- CPU arm
def cpu_candidates(x, W, plugins, model, target_embeds):
    """The reference's CPU path for a slice of the population: serial process_audio per candidate
    (style_transfer.py:512-521) then one batched encoder call with all host threads."""
    import torch

    from oracle import cnn14, dsp

    import copy
    from concurrent.futures import ThreadPoolExecutor

    # DSP: one candidate per host thread (the reference's `parallel=True` pool, style_transfer.py:499-502; the
    # oracle's C kernels release the GIL).  Plugin objects are stateful, so every worker gets its own copy.
    workers = max(1, min(len(W), os.cpu_count() or 1))
    pool_plugins = [copy.deepcopy(plugins) for _ in range(workers)]

    def render(job):
        k, w = job
        return dsp.process_audio(x, w, SR, pool_plugins[k % workers])

    fits = []
    for i in range(0, len(W), 8):  # encoder in batches of 8 candidates: bounded activation memory on the host
        with ThreadPoolExecutor(max_workers=workers) as ex:
            rendered = list(ex.map(render, [(k, w) for k, w in enumerate(W[i:i + 8])]))
        audios = torch.stack([torch.from_numpy(a) for a in rendered])
        emb = cnn14.get_param_embeds(audios, model, SR)
        fits.append(cnn14.fitness(emb, target_embeds))
    return torch.cat(fits)


def cpu_setup(x, chain_kinds):
    import torch

    from oracle import cnn14, dsp

    # the CPU arm uses every host thread it can: torchrun exports OMP_NUM_THREADS=1, which would cripple the baseline
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    dsp.build()
    plugins, D, _ = dsp.load_plugins(dsp.make_plugins(chain_kinds))
    model = cnn14.make_encoder(seed=3)
    w_star = np.random.RandomState(1234).rand(D)
    tgt = dsp.process_audio(x, w_star, SR, plugins)
    te = cnn14.get_param_embeds(torch.from_numpy(tgt[None].copy()), model, SR)
    return plugins, D, model, te


def run_reference(args, rank):
    """--impl reference: the CPU path on this box's host cores; rank 0 only."""
    if rank != 0:
        return
    import torch

    x = make_workload(args.seconds)
    plugins, D, model, te = cpu_setup(x, ["eq", "comp", "reverb"])
    sample = args.ref_sample
    times = []
    for step in range(args.warmup + args.steps):
        W = np.random.RandomState(100 + step).rand(sample, D)
        t0 = time.perf_counter()
        cpu_candidates(x, W, plugins, model, te)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config2: {args.seconds:g}s stereo 48kHz, EQ+Compressor+Reverb, pop={args.pop}",
                   "sample": f"{sample} candidates of the population per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} candidates/step x {len(times)} steps, oracle DSP (C, serial per "
                                   f"candidate) + torch CPU Cnn14 ({cores} threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from st_ito_b200 import effects
    from st_ito_b200.engine import compile_chain
    from st_ito_b200.style_transfer import load_plugins, process_audio
    from st_ito_b200.utils import make_synthetic_param_model

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peaks = load_peaks()
    P, L = args.pop, int(args.seconds * SR)
    x = make_workload(args.seconds)

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        plugins, D, _ = load_plugins(effects.make_chain("mastering-pb"))
    model = make_synthetic_param_model(seed=3)
    eng = model.stito_engine(local_rank)
    if args.precision is not None:
        eng.set_precision(args.precision)
    desc, _ = compile_chain(plugins, SR)
    eng.set_chain(desc)
    w_star = np.random.RandomState(1234).rand(D)
    eng.set_target(process_audio(x, w_star, SR, plugins))
    eng.set_input(x)
    x_pinned = torch.from_numpy(x).pin_memory()

    # seeded, sampler-independent populations, distinct per rank and per step; drawn before the timed
    # region (they are the synthetic inputs of the steps; uploading them IS timed)
    _pops = {s: np.random.RandomState(1000 * (rank + 1) + s).rand(P, D) for s in range(args.warmup + args.steps)}

    def population(step):
        return _pops[step]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fit_all = torch.empty(world * P, dtype=torch.float32, device=dev) if world > 1 else None

    def generation(step, e2e):
        W = population(step)
        if e2e:
            eng.set_input(x_pinned)
        fit, emb, _ = eng.eval_population(W, 0, L, want_embeds=e2e)
        if world > 1:  # one all-gather of the scalar fitness values per generation
            dist.all_gather_into_tensor(fit_all, fit.to(dev, non_blocking=True))
        return fit

    def timed(e2e):
        for s in range(args.warmup):
            generation(s, e2e)
        launches, stage = 0, {"ms_dsp": 0.0, "ms_frontend": 0.0, "ms_encoder": 0.0, "ms_fitness": 0.0}
        conv_ms = np.zeros(12)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for s in range(args.steps):
            generation(args.warmup + s, e2e)
            t = eng.timing()
            launches += t["launches"]
            for k in stage:
                stage[k] += t[k]
            conv_ms += np.array(t["ms_conv"])
            last = t
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, wall, launches, stage, conv_ms, last

    sampler = ClockSampler(local_rank) if rank == 0 else None
    stream = torch.cuda.Stream(dev)  # libstito launches on torch's current stream; events are recorded on it
    with torch.cuda.stream(stream):
        if sampler:
            sampler.start()
        ms, wall, launches, stage, conv_ms, last = timed(e2e=False)
        clocks = sampler.stop() if sampler else None
        ms_e2e, _, _, _, _, _ = timed(e2e=True)

    if rank != 0:
        return
    K = args.steps
    value = world * P * K / (ms / 1e3)
    e2e_value = world * P * K / (ms_e2e / 1e3)
    tc = last["precision"] == 1
    # dominant kernel: the tcgen05 implicit-GEMM convolution (11 launches per generation: conv layers 2..12; layer 1,
    # Cin = 1 / K = 9, is a separate memory-bound CUDA-core kernel).  Algorithmic FLOPs of those launches / their
    # CUDA-event time on the launch stream (per-layer events recorded inside libstito).
    T = L // 1024 + 1
    ch = [1, 64, 128, 256, 512, 1024, 2048]
    layer_flop, hh, ww = [], T, 128
    for b in range(6):
        layer_flop += [2.0 * hh * ww * ch[b + 1] * 9 * ch[b], 2.0 * hh * ww * ch[b + 1] * 9 * ch[b + 1]]
        hh, ww = hh // 2, ww // 2
    layer_flop = [f * 2 * P for f in layer_flop]  # 2 signals (mid, side) per stereo candidate
    ms_layers = [float(v) / K for v in conv_ms]
    tc_flop, tc_ms = sum(layer_flop[1:]), sum(ms_layers[1:])
    achieved = tc_flop / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    # tensor-pipe work actually executed: 3 MMAs per MAC (fp16x3); layers with Cin*Cout/(Cin+Cout) >= 340 (conv 9..12)
    # run as Winograd F(2x2,3x3) GEMMs with 4/9 of the direct MACs unless STITO_TC_WINOGRAD=0
    wino = os.environ.get("STITO_TC_WINOGRAD", "1") != "0"
    wthr = int(os.environ.get("STITO_TC_WINO_MIN", "340"))
    cin = [1, 64, 64, 128, 128, 256, 256, 512, 512, 1024, 1024, 2048]
    cout = [64, 64, 128, 128, 256, 256, 512, 512, 1024, 1024, 2048, 2048]
    executed = sum(f * 3.0 * ((4.0 / 9.0) if (wino and ci >= 512 and ci * co >= wthr * (ci + co)) else 1.0)
                   for f, ci, co in zip(layer_flop[1:], cin[1:], cout[1:]))
    executed_tflops = executed / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    peak = peaks["tflops_sustained"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01f_conv_traffic.json")
    if tc and P == 64 and abs(args.seconds - 10.0) < 1e-9 and os.path.isfile(tpath):
        with open(tpath) as f:
            tj = json.load(f)
            traffic = tj["dram_bytes_per_generation"] / tj["launches"]  # per launch, like `achieved`
    roofline = {
        "kernel": ("conv3x3_tc_kernel / conv3x3_c64_kernel (tcgen05 implicit-GEMM 3x3 conv / Winograd GEMMs, fp16x3 split "
                   "precision), conv layers 2..12 of every generation") if tc else "conv3x3 fp32 CUDA-core kernel (precision 0)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_unit": "bytes per launch (mean of the 11 launches; ncu dram read+write, "
                                            "profiles/r01f_conv_traffic.json)",
        "peak_source": f"{peaks['source']} bf16 sustained (cuBLAS, MEASURED_PEAKS.json)",
        "note": "achieved = algorithmic (direct-convolution) 2*MAC of conv layers 2..12 (%.2f GFLOP per stereo candidate) / "
                "CUDA-event time of those layers; the fp16x3 scheme executes 3 tensor-core MACs per MAC and the four "
                "deepest layers run as Winograd GEMMs (4/9 of the MACs): see tensor_pipe_executed_tflops" % (tc_flop / P / 1e9),
        "tensor_pipe_executed_tflops": executed_tflops if tc else None,
        "ms_per_layer": ms_layers,
        "stages_ms": {k: v / K for k, v in stage.items()},
        "hbm": {"dsp_GBps": last["dsp_bytes"] / max(stage["ms_dsp"] / K, 1e-9) / 1e6,
                "frontend_GBps": last["frontend_bytes"] / max(stage["ms_frontend"] / K, 1e-9) / 1e6,
                "peak_GBps": peaks["hbm_gbs"]},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16x3+f32acc (encoder), f64 (EQ), f32 (comp/reverb/FFT)" if tc else "f32 (encoder), f64 (EQ)",
        "data": "synthetic",
        "config": {"workload": f"config2: {args.seconds:g}s stereo 48kHz, EQ+Compressor+Reverb (mastering-pb), "
                               f"pop={P} per GPU, D={D}, AFx-Rep Cnn14 seeded random weights",
                   "l2": "per-step working set (activations > 1 GB) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"population sharded, {world} x {P} candidates, fitness all-gather"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes + P * D * 8),
                "d2h_bytes_per_step": int(P * 4 + 2 * P * 512 * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "wall_s": wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        import torch as _t

        plugins_o, D_o, model_o, te = cpu_setup(x, ["eq", "comp", "reverb"])
        n = args.cpu_sample
        Wc = np.random.RandomState(7).rand(n, D_o)
        t0 = time.perf_counter()
        cpu_candidates(x, Wc, plugins_o, model_o, te)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port",
                                "sample": f"{n} candidates of the same workload: oracle DSP (C, one candidate per host thread) "
                                          f"+ torch CPU Cnn14, {dt:.1f} s"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pop", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--precision", type=int, default=None, help="0 = fp32 CUDA cores, 1 = fp16x3 tcgen05")
    ap.add_argument("--cpu-sample", type=int, default=96,
                    help="candidates timed for cpu_baseline in the default arm (about 10-30 s of host work)")
    ap.add_argument("--ref-sample", type=int, default=16,
                    help="candidates per step of the --impl reference arm (bounded sample of the P=64 population)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:  # plain `python bench.py --gpus N`: spawn the ranks
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    if world > 1:
        import torch
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
