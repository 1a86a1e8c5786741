#!/usr/bin/env python
"""bench.py -- ES candidates/sec of the population-evaluation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C] [--pop P] [--weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (default = BASELINE config 2): 10 s stereo 48 kHz input, "mastering-pb" chain (EQ -> Compressor -> Reverb,
D = 29), CMA-ES with a population of 64, AFx-Rep (Cnn14) embedding metric, 25 generations per ES run.

A "step" is one ES run segment of `--iters` generations (25) driven by the real host loop:
    W = es.ask()  ->  evaluate(W) [render every candidate, log-mel, encoder, cosine fitness]  ->  es.tell(W, fvals)
with a fresh, seeded CMA-ES per step (find_w0 and the final render are outside the metric).  K steps are timed with CUDA
events on the launch stream (max over ranks); the CMA-ES host time therefore counts (it shows up as device idle time).

value  : candidates/s with the input waveform resident in HBM (per generation only the parameter block travels).
e2e    : the same loop through the host-buffer API: EVERY generation re-uploads the input waveform from pinned host
         memory and reads fitness + embeddings back.
--gpus N (torchrun): STRONG scaling by default -- the population of P candidates is sharded over the N ranks
         (P/N each) and the fitness values are all-gathered over NCCL every generation; --weak gives every rank its
         own P candidates instead (population N*P).
The reference arm (--impl reference) and cpu_baseline time the CPU oracle -- a restatement of the reference's CPU path;
the reference itself is pure Python over scipy/pedalboard/torch and cannot be installed offline -- on this
box's host cores, one generation (P candidates) per step.
"""
import argparse
import json
import math
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
UNIT = "candidates/s"

# BASELINE.json configs: (seconds, channels, chain preset, population, generations per ES run)
CONFIGS = {
    1: dict(seconds=5.0, chs=1, chain="eq", pop=8, iters=5, label="EQ-only"),
    2: dict(seconds=10.0, chs=2, chain="mastering-pb", pop=64, iters=25, label="EQ+Compressor+Reverb (mastering-pb)"),
    3: dict(seconds=10.0, chs=2, chain="mastering-pb", pop=256, iters=50, label="EQ+Compressor+Reverb (mastering-pb)"),
    4: dict(seconds=30.0, chs=2, chain="mastering-conv", pop=128, iters=25,
            label="EQ+Compressor+2s-IR convolution reverb (mastering-conv)"),
}
ORACLE_KINDS = {"eq": ["eq"], "mastering-pb": ["eq", "comp", "reverb"], "basic": ["eq", "comp", "dist", "delay", "reverb"],
                "mastering-conv": ["eq", "comp", "convreverb2s"], "mastering-dasp": ["eq", "lticomp", "convreverb2s"]}


def resolve(args):
    c = dict(CONFIGS[args.config])
    if args.seconds is not None:
        c["seconds"] = args.seconds
    if args.pop is not None:
        c["pop"] = args.pop
    if args.iters is not None:
        c["iters"] = args.iters
    if args.chain is not None:
        c["chain"] = args.chain
        c["label"] = args.chain
    return c


def metric_name(c, pop_total):
    return f"ES candidates/sec (pop={pop_total}, {c['seconds']:g}s@48kHz {'stereo' if c['chs'] == 2 else 'mono'})"


def workload_label(args, c, pop_total, D):
    """Identical in both arms (the driver compares it)."""
    return (f"config{args.config}: {c['seconds']:g}s {'stereo' if c['chs'] == 2 else 'mono'} 48kHz, {c['label']}, "
            f"pop={pop_total}, D={D}, {c['iters']} generations per ES run, AFx-Rep Cnn14 (seeded synthetic weights, "
            f"centred heads)")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""

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
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
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
def cpu_candidates(x, W, plugins, model, target_embeds, serial=False):
    """The reference's CPU path for a population: process_audio per candidate -- on a pool of host threads like the
    reference's `parallel=True` branch (style_transfer.py:499-502; the oracle's C kernels release the GIL), or one
    after the other like its default branch (:512-521) with serial=True -- then the batched encoder call on all cores."""
    import copy
    from concurrent.futures import ThreadPoolExecutor

    import torch

    from oracle import cnn14, dsp

    if x.shape[-1] <= 262144:  # evaluate()'s length policy (style_transfer.py:505-518): right-zero-pad to 262144
        x = np.pad(x, ((0, 0), (0, 262144 - x.shape[-1])))
    workers = 1 if serial else max(1, min(len(W), os.cpu_count() or 1))
    pool_plugins = [copy.deepcopy(plugins) for _ in range(workers)]  # plugin objects are stateful

    def render(job):
        k, w = job
        return dsp.process_audio(x, w, SR, pool_plugins[k % workers])

    fits = []
    for i in range(0, len(W), 8):  # encoder in batches of 8 candidates: bounded activation memory on the host
        jobs = [(k, w) for k, w in enumerate(W[i:i + 8])]
        if serial:
            rendered = [render(j) for j in jobs]
        else:
            with ThreadPoolExecutor(max_workers=workers) as ex:
                rendered = list(ex.map(render, jobs))
        audios = torch.stack([torch.from_numpy(a) for a in rendered])
        emb = cnn14.get_param_embeds(audios, model, SR)
        fits.append(cnn14.fitness(emb, target_embeds))
    return torch.cat(fits)


def cpu_setup(x, chain, head_bias=None):
    """Oracle plugins / encoder for the CPU legs.  `head_bias` = (fc_mid.bias, fc_side.bias) of the GPU arm's model, so
    that both score with the same weights; without it the heads are centred by the oracle itself."""
    import torch

    from oracle import cnn14, dsp

    # the CPU arm uses every host thread it can: torchrun exports OMP_NUM_THREADS=1, which would cripple the baseline
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    dsp.build()
    plugins, D, _ = dsp.load_plugins(dsp.make_plugins(ORACLE_KINDS[chain]))
    model = cnn14.make_encoder(seed=3, conv_gain=2.0)
    if head_bias is None:
        cnn14.centre_heads(model)
    else:
        with torch.no_grad():
            model.fc_mid.bias.copy_(head_bias[0])
            model.fc_side.bias.copy_(head_bias[1])
    w_star = np.random.RandomState(1234).rand(D)
    tgt = dsp.process_audio(x, w_star, SR, plugins)
    te = cnn14.get_param_embeds(torch.from_numpy(tgt[None].copy()), model, SR)
    return plugins, D, model, te


def run_reference(args, rank):
    """--impl reference: the CPU path on this box's host cores; rank 0 only.  One generation (P candidates) per step."""
    if rank != 0:
        return
    import torch

    c = resolve(args)
    P = c["pop"]
    x = make_workload(c["seconds"], c["chs"])
    plugins, D, model, te = cpu_setup(x, c["chain"])
    times = []
    for step in range(args.warmup + args.steps):
        W = np.random.RandomState(100 + step).rand(P, D)
        t0 = time.perf_counter()
        cpu_candidates(x, W, plugins, model, te)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = P * len(times) / total
    cores = torch.get_num_threads()
    ns = min(P, 16)  # the reference's default branch renders the candidates one after the other
    t0 = time.perf_counter()
    cpu_candidates(x, np.random.RandomState(99).rand(ns, D), plugins, model, te, serial=True)
    value_serial = ns / (time.perf_counter() - t0)
    line = {
        "impl": "reference", "metric": metric_name(c, P), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (encoder), f64 (EQ)",
        "data": "synthetic",
        "config": {"workload": workload_label(args, c, P, D),
                   "sample": f"one generation ({P} candidates) of the {c['iters']}-generation ES run per step"},
        "cpu_baseline": {"value": value, "value_serial": value_serial, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{P} candidates/step x {len(times)} steps: oracle DSP (C, one candidate per host "
                                   f"thread) + torch CPU Cnn14 ({cores} threads); value_serial: {ns} candidates rendered "
                                   f"one after the other (the reference's default branch, style_transfer.py:512-521)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import contextlib
    import io

    import torch
    import torch.distributed as dist

    from st_ito_b200 import cma, effects
    from st_ito_b200.style_transfer import FusedEvaluator, load_plugins, process_audio
    from st_ito_b200.utils import centre_heads, get_param_embeds, make_synthetic_param_model

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peaks = load_peaks()
    c = resolve(args)
    P_total = c["pop"] * (world if args.weak else 1)
    L, iters = int(c["seconds"] * SR), c["iters"]
    x = make_workload(c["seconds"], c["chs"])

    with contextlib.redirect_stdout(io.StringIO()):
        plugins, D, _ = load_plugins(effects.make_chain(c["chain"]))
    model = centre_heads(make_synthetic_param_model(seed=3, conv_gain=2.0))
    eng = model.stito_engine(local_rank)
    if args.precision is not None:
        eng.set_precision(args.precision)
    w_star = np.random.RandomState(1234).rand(D)
    tgt = process_audio(x, w_star, SR, plugins)
    te = get_param_embeds(torch.from_numpy(tgt[None].copy()), model, SR)
    x_t = torch.from_numpy(x)
    x_pinned = x_t.pin_memory()
    ev = FusedEvaluator(eng, plugins, SR, te, x_t[None])
    start, length = ev.view_for(L, False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    record = {}  # (step, generation) -> (W, fitness): the populations the CPU leg re-scores for the parity object

    def es_step(step, e2e, acc=None, keep=0):
        """One ES run segment: `iters` generations of ask -> evaluate -> tell with a fresh seeded CMA-ES."""
        es = cma.CMAEvolutionStrategy(np.full(D, 0.5), 0.33, {"bounds": [0, 1], "popsize": P_total, "seed": 1 + step,
                                                              "verbose": -9})
        for g in range(iters):
            t0 = time.perf_counter()
            W = np.asarray(es.ask())
            t1 = time.perf_counter()
            if e2e:
                eng.set_input(x_pinned, min_len=ev.crop_len)
            fvals, emb, _ = ev(W, want_embeds=e2e)
            t2 = time.perf_counter()
            es.tell(list(W), fvals)
            t3 = time.perf_counter()
            if acc is not None:
                acc["host_cma_ms"] += 1e3 * ((t1 - t0) + (t3 - t2))
                if g % 5 == 0:  # per-stage / per-layer event times: sampled every 5th generation (the query costs host time)
                    t = eng.timing()
                    acc["sampled"] += 1
                    acc["launches"] += t["launches"]
                    for k in ("ms_dsp", "ms_frontend", "ms_encoder", "ms_fitness"):
                        acc[k] += t[k]
                    acc["conv"] += np.array(t["ms_conv"])
                    acc["comp_fallbacks"] += t["comp_fallbacks"]
                    acc["act_overflow"] += t["act_overflow"]
                    acc["last"] = t
            if g < keep:
                record[(step, g)] = (W.copy(), np.array(fvals, dtype=np.float32))

    def timed(e2e):
        for s in range(args.warmup):
            es_step(s, e2e)
        acc = {"launches": 0, "ms_dsp": 0.0, "ms_frontend": 0.0, "ms_encoder": 0.0, "ms_fitness": 0.0,
               "conv": np.zeros(12), "host_cma_ms": 0.0, "comp_fallbacks": 0, "act_overflow": 0, "last": None, "sampled": 0}
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for s in range(args.steps):
            es_step(args.warmup + s, e2e, acc, keep=(args.cpu_sample if (s == 0 and not e2e) else 0))
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
            ll = torch.tensor([float(acc["launches"])], device=dev)
            dist.all_reduce(ll, op=dist.ReduceOp.SUM)
            acc["launches"] = int(ll.item())
        return ms, wall, acc

    sampler = ClockSampler(local_rank) if rank == 0 else None
    stream = torch.cuda.Stream(dev)  # libstito launches on torch's current stream; events are recorded on it
    with torch.cuda.stream(stream):
        if sampler:
            sampler.start()
        ms, wall, acc = timed(e2e=False)
        clocks = sampler.stop() if sampler else None
        ms_e2e, _, _ = timed(e2e=True)
        shard_check = None
        if world > 1:
            # sharded evaluation + all-gather must equal one rank scoring the whole population, bit for bit
            # (the single rank scores the same shards one after the other: which DSP kernels run depends on the number of
            # candidates per call; against ONE call over the whole population the difference is float32 rounding noise)
            from st_ito_b200 import dist as sdist

            Wc = np.random.RandomState(4242).rand(P_total, D)
            gathered = np.array(ev(Wc)[0], dtype=np.float32)
            parts = []
            for r in range(world):
                lo, hi, _ = sdist.shard_bounds(P_total, world, r)
                if hi > lo:
                    parts.append(eng.eval_population(Wc[lo:hi], start, length)[0].numpy())
            shardwise = np.concatenate(parts)
            whole = eng.eval_population(Wc, start, length)[0].numpy()
            same = torch.tensor([1.0 if np.array_equal(gathered, shardwise) else 0.0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            shard_check = {"population": P_total, "ranks": world, "gathered_equals_single_rank_bitwise": bool(same.item() == 1.0),
                           "max_abs_diff_vs_one_call_over_the_whole_population": float(np.abs(gathered - whole).max())}

    if rank != 0:
        return
    K = args.steps
    gens = K * iters
    value = P_total * gens / (ms / 1e3)
    e2e_value = P_total * gens / (ms_e2e / 1e3)
    last = acc["last"]
    tc = last["precision"] == 1
    P_rank = P_total // world if world > 1 else P_total  # candidates per rank (rank 0's shard)
    # dominant kernel: the tcgen05 implicit-GEMM convolution (conv layers 2..12; layer 1, Cin = 1 / K = 9, is a separate
    # memory-bound CUDA-core kernel).  Algorithmic FLOPs of those launches / their CUDA-event time on the launch stream
    # (per-layer events recorded inside libstito), on rank 0's shard.
    T = length // 1024 + 1
    ochs = 2 if (c["chs"] == 2 or c["chain"] != "eq") else 1
    ch = [1, 64, 128, 256, 512, 1024, 2048]
    layer_flop, hh, ww = [], T, 128
    for b in range(6):
        layer_flop += [2.0 * hh * ww * ch[b + 1] * 9 * ch[b], 2.0 * hh * ww * ch[b + 1] * 9 * ch[b + 1]]
        hh, ww = hh // 2, ww // 2
    layer_flop = [f * ochs * P_rank for f in layer_flop]  # one log-mel image per output channel (mid, side)
    ns = max(acc["sampled"], 1)   # generations whose stage / layer events were read
    ms_layers = [float(v) / ns for v in acc["conv"]]
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
    traffic, traffic_src = None, None
    for name in ("r02_conv_traffic.json", "r01f_conv_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if tc and world == 1 and args.config == 2 and P_total == 64 and os.path.isfile(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_generation"] / tj["launches"]  # per launch, like `achieved`
            traffic_src = f"profiles/{name}: ncu dram read+write of the conv-stack kernels, mean over its {tj['launches']} launches per generation"
            break
    stages = {k: acc[k] / ns for k in ("ms_dsp", "ms_frontend", "ms_encoder", "ms_fitness")}
    roofline = {
        "kernel": ("conv3x3_tc_kernel / conv3x3_c64_kernel (tcgen05 implicit-GEMM 3x3 conv / Winograd GEMMs, fp16x3 split "
                   "precision), conv layers 2..12 of every generation") if tc else "conv3x3 fp32 CUDA-core kernel (precision 0)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_unit": traffic_src,
        "peak_source": f"{peaks['source']} bf16 sustained (cuBLAS, MEASURED_PEAKS.json); the timed region is {ms / 1e3:.1f} s",
        "note": "achieved = algorithmic (direct-convolution) 2*MAC of conv layers 2..12 (%.2f GFLOP per candidate) / "
                "CUDA-event time of those layers; the fp16x3 scheme executes 3 tensor-core MACs per MAC and the four "
                "deepest layers run as Winograd GEMMs (4/9 of the MACs): see tensor_pipe_executed_tflops" % (tc_flop / max(P_rank, 1) / 1e9),
        "tensor_pipe_executed_tflops": executed_tflops if tc else None,
        "ms_per_layer": ms_layers,
        "stages_ms_per_generation": stages,
        "hbm": {"dsp_GBps": last["dsp_bytes"] / max(stages["ms_dsp"], 1e-9) / 1e6,
                "frontend_GBps": last["frontend_bytes"] / max(stages["ms_frontend"], 1e-9) / 1e6,
                "peak_GBps": peaks["hbm_gbs"],
                "note": "algorithmic bytes (SURVEY 8d) / CUDA-event time of the DSP chain and of the log-mel kernel"},
    }
    h2d = int(x.nbytes + (P_rank if world > 1 else P_total) * D * 8)
    d2h = int(P_rank * 4 + 2 * P_rank * 512 * 4)
    line = {
        "metric": metric_name(c, P_total), "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f16x3+f32acc (encoder), f64 (EQ), f32 (comp/reverb/FFT)" if tc else "f32 (encoder), f64 (EQ)",
        "data": "synthetic",
        "config": {"workload": workload_label(args, c, P_total, D),
                   "step": f"one ES run segment = {iters} generations of ask -> evaluate -> tell (fresh seeded CMA-ES, "
                           f"sigma0 = 0.33, bounds [0, 1]); {gens} generations timed",
                   "l2": "per-generation working set (activations > 1 GB at 64 candidates) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": (f"population of {P_total} sharded over {world} ranks ({P_rank} candidates each), one "
                                   f"fitness all-gather per generation: "
                                   + ("fused into the fitness kernel over NVLink peer memory (stito_eval_population_gather)"
                                      if ev.peer_gather else "NCCL all_gather_into_tensor")) if world > 1 else "single GPU"},
        "ms_per_generation": ms / gens,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * iters, "d2h_bytes_per_step": d2h * iters,
                "ms_per_generation": ms_e2e / gens,
                "what": "same loop; every generation re-uploads the input waveform + W from pinned host memory and reads "
                        "fitness + embeddings back (per rank)"},
        "gpu_launches": int(round(acc["launches"] * gens / ns)),
        "clocks": clocks,
        "roofline": roofline,
        "host_cma_ms_per_generation": acc["host_cma_ms"] / gens,
        "comp_fallbacks_sampled": int(acc["comp_fallbacks"]), "act_overflow_sampled": int(acc["act_overflow"]),
        "stage_times_sampled_generations": int(ns),
        "wall_s": wall,
    }
    if shard_check is not None:
        line["shard_check"] = shard_check
    if world == 1 and not args.no_cpu_baseline and args.cpu_sample > 0:
        import torch as _t

        bias = (model.fc_mid.bias.detach().cpu().clone(), model.fc_side.bias.detach().cpu().clone())
        plugins_o, D_o, model_o, te_o = cpu_setup(x, c["chain"], head_bias=bias)
        keys = sorted(record)
        n, dt, max_rel, argsort_equal, near_ties = 0, 0.0, 0.0, True, 0
        for k in keys:  # the CPU leg re-scores populations the GPU evaluated inside the timed region
            Wc, f_gpu = record[k]
            t0 = time.perf_counter()
            f_cpu = cpu_candidates(x, Wc, plugins_o, model_o, te_o).numpy()
            dt += time.perf_counter() - t0
            n += len(Wc)
            max_rel = max(max_rel, float(np.max(np.abs(f_gpu - f_cpu) / np.maximum(np.abs(f_cpu), 1e-3))))
            argsort_equal &= bool(np.array_equal(np.argsort(f_gpu, kind="stable"), np.argsort(f_cpu, kind="stable")))
            near_ties += int(np.sum(np.diff(np.sort(f_cpu.astype(np.float64))) < 1e-6))
        ns = min(len(record[keys[0]][0]), 16)
        t0 = time.perf_counter()
        cpu_candidates(x, record[keys[0]][0][:ns], plugins_o, model_o, te_o, serial=True)
        value_serial = ns / (time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": n / dt, "value_serial": value_serial, "unit": UNIT,
                                "cores": _t.get_num_threads(), "kind": "port",
                                "sample": f"{n} candidates = the first {len(keys)} populations of the timed region: oracle DSP "
                                          f"(C, one candidate per host thread) + torch CPU Cnn14, {dt:.1f} s; value_serial: "
                                          f"{ns} candidates rendered one after the other (reference default branch)"}
        line["parity"] = {"candidates": n, "max_rel_err": max_rel, "argsort_equal": argsort_equal,
                          "near_ties_below_1e-6": near_ties,
                          "what": "fitness of the populations timed on the GPU vs the CPU oracle on the same W; "
                                  "rel err = |df| / max(|f|, 1e-3), gate 1e-4; argsort per population"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (default 2)")
    ap.add_argument("--pop", type=int, default=None, help="total population (default: the config's)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --pop candidates PER GPU (default: sharded)")
    ap.add_argument("--seconds", type=float, default=None)
    ap.add_argument("--iters", type=int, default=None, help="generations per step (default: the config's ES length)")
    ap.add_argument("--chain", default=None, choices=sorted(ORACLE_KINDS))
    ap.add_argument("--precision", type=int, default=None, help="0 = fp32 CUDA cores, 1 = fp16x3 tcgen05")
    ap.add_argument("--cpu-sample", type=int, default=2,
                    help="populations of the timed region re-scored by the CPU oracle (cpu_baseline + parity)")
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
