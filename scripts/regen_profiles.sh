#!/bin/bash
# Turn the outputs of scripts/gpu_session_final.sh (gpurun_out/f_*) into the tracked summaries under profiles/ (developer tool).
cd "$(dirname "$0")/.."
T=gpurun_out/f
# the LTI-compressor kernels were rewritten after the evidence session: their files come from scripts/gpu_session_lti.sh
for f in prof_lti_raw.csv launches_dasp.csv bench_dasp.jsonl; do [ -f gpurun_out/l1_$f ] && cp gpurun_out/l1_$f ${T}_$f; done
n_of() { python - "$1" <<'PY'
import csv, sys
print(len(list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith('==')]))))
PY
}
python scripts/make_profile_summary.py r02 ${T}_launches_p64.csv ${T}_prof_conv_raw.csv $(n_of ${T}_launches_p64.csv)
gen() {  # tag-suffix launches.csv "header"
python - "$1" "$2" "$3" <<'PY'
import csv, sys
suffix, path, head = sys.argv[1:4]
rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
out = [f"# r02: per-kernel device time of ONE generation, {head}; ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)", "kernel,grid,block,us"]
tot = 0.0
for r in rows:
    us = float(r["Metric Value"]) / 1e3
    tot += us
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    out.append(f"{name},{r['Grid Size'].replace(',', ' ')},{r['Block Size'].replace(',', ' ')},{us:.1f}")
out.append(f"TOTAL,,,{tot:.1f}")
open(f"profiles/r02_generation_kernels_{suffix}.csv", "w").write("\n".join(out) + "\n")
PY
}
gen p8 ${T}_launches_p8.csv "P=8 (pop = 64 sharded over 8 GPUs), 10 s stereo, EQ+Comp+Reverb: streaming pair + cluster-split Freeverb"
gen c4 ${T}_launches_c4.csv "config 4 chain at P=16: 30 s stereo, EQ + Compressor + 2 s-IR convolution reverb"
gen dasp ${T}_launches_dasp.csv "dasp-style chain (SURVEY row R2) at P=16: 10 s stereo, EQ + compressor with LTI smoothing + 2 s-IR convolution reverb"
python scripts/make_hbm_summary.py ${T}_prof_dsp_raw.csv profiles/r02_dsp_frontend_ncu_full_summary.csv "r02: ncu --set full --clock-control none of the HBM-bound kernels of ONE generation (P=64, 10 s stereo, EQ+Comp+Reverb): python scripts/dev_generation.py 64 1"
python scripts/make_hbm_summary.py ${T}_prof_crv_raw.csv profiles/r02_convreverb_ncu_full_summary.csv "r02: ncu --set full --clock-control none of the convolution-reverb kernels (config 4 chain, P=16, 30 s stereo, 96000-tap IR; quarter-circle twiddle table in shared memory): python scripts/dev_generation.py 16 1 30 mastering-conv"
python scripts/make_hbm_summary.py ${T}_prof_p8_raw.csv profiles/r02_small_population_ncu_full_summary.csv "r02: ncu --set full --clock-control none of the small-population pair (P=8, 10 s stereo): streaming compressor + cluster-split Freeverb: python scripts/dev_generation.py 8 1"
python scripts/make_hbm_summary.py ${T}_prof_lti_raw.csv profiles/r02_lticomp_ncu_full_summary.csv "r02: ncu --set full --clock-control none of the compressor with LTI gain smoothing (P=16, 10 s stereo, stereo-linked): python scripts/dev_generation.py 16 1 10 mastering-dasp"
cp ${T}_bench.jsonl profiles/r02_bench_config2.jsonl
cp ${T}_bench_c1.jsonl profiles/r02_bench_config1.jsonl
cp ${T}_bench_c4.jsonl profiles/r02_bench_config4.jsonl
cp ${T}_bench_dasp.jsonl profiles/r02_bench_dasp_chain.jsonl
cp ${T}_bench_ref.jsonl profiles/r02_bench_reference_arm.jsonl
python - <<'PY'
import json
out = ["# r02: population sweep (BASELINE config 5) at fixed 10 s stereo 48 kHz, EQ+Comp+Reverb, 1 x B200: python bench.py --pop P --steps 1 --warmup 1 --iters 8 --no-cpu-baseline",
       "# real host loop (ask -> evaluate -> tell); P <= 32 uses the streaming compressor->reverb pair, P <= 14 the cluster-split Freeverb",
       "pop,candidates_per_s,e2e_candidates_per_s,ms_per_generation,conv_roofline_frac_algorithmic,ms_dsp,ms_logmel,ms_encoder,ms_host_cma"]
for ln in open("gpurun_out/f_pop_sweep.jsonl"):
    if not ln.startswith("{"): continue
    d = json.loads(ln); r = d["roofline"]; s = r["stages_ms_per_generation"]
    pop = int(d["config"]["workload"].split("pop=")[1].split(",")[0])
    out.append(f"{pop},{d['value']:.0f},{d['e2e']['value']:.0f},{d['ms_per_generation']:.3f},{r['frac']:.3f},{s['ms_dsp']:.3f},{s['ms_frontend']:.3f},{s['ms_encoder']:.3f},{d['host_cma_ms_per_generation']:.3f}")
open("profiles/r02_pop_sweep_1gpu.csv", "w").write("\n".join(out) + "\n")
PY
python scripts/sass_census.py > profiles/r02_sass_census.csv
