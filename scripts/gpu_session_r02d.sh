#!/bin/bash
# round 2, GPU session D: split-reverb fix check, ncu source-level profile of it, two-pass sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "config2_full or single_plugin or ragged or other_sample_rates or properties_at_full" > gpurun_out/d_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_gpu_tests.log
for p in 8 16; do
  timeout 300 python bench.py --pop $p --steps 2 --warmup 1 --no-cpu-baseline >> gpurun_out/d_pop_sweep.jsonl 2>> gpurun_out/d_bench.err
done
STITO_REVERB_SPLIT=0 timeout 300 python bench.py --pop 8 --steps 2 --warmup 1 --no-cpu-baseline >> gpurun_out/d_pop8_nosplit.jsonl 2>> gpurun_out/d_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/d_launches_p8.csv python scripts/dev_generation.py 8 1 > gpurun_out/d_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"reverb_split|compressor" -o gpurun_out/d_prof_split -f python scripts/dev_generation.py 8 1 > gpurun_out/d_ncu2.log 2>&1
ncu -i gpurun_out/d_prof_split.ncu-rep --page source --csv -k regex:reverb_split > gpurun_out/d_split_src.csv 2>/dev/null
ncu -i gpurun_out/d_prof_split.ncu-rep --page raw --csv > gpurun_out/d_prof_split_raw.csv 2>/dev/null
timeout 900 python tests/dev/dev_two_pass.py > gpurun_out/d_two_pass.jsonl 2> gpurun_out/d_two_pass.err
for m in 0 0xFFE; do
  STITO_TC_DROP_ALO=$m timeout 300 python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu-baseline >> gpurun_out/d_drop_alo.jsonl 2>> gpurun_out/d_bench.err
done
STITO_TC_DROP_BLO=0xFFE timeout 300 python bench.py --steps 1 --warmup 1 --iters 5 --no-cpu-baseline >> gpurun_out/d_drop_blo.jsonl 2>> gpurun_out/d_bench.err
grep -E "passed|failed|FAILED|rc=" gpurun_out/d_gpu_tests.log | tail -5
python - <<'PY'
import json
for f in ['d_pop_sweep.jsonl','d_pop8_nosplit.jsonl','d_drop_alo.jsonl','d_drop_blo.jsonl']:
    for ln in open('gpurun_out/'+f):
        if not ln.startswith('{'): continue
        d=json.loads(ln); r=d['roofline']
        print(f, d['metric'][:38], 'ms/gen %.3f'%d['ms_per_generation'], {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()}, [round(x,3) for x in r['ms_per_layer']])
PY
python scripts/launch_summary.py gpurun_out/d_launches_p8.csv 34 | head -8
python scripts/ncu_top.py gpurun_out/d_split_src.csv 30 | cut -c1-250
head -c 3000 gpurun_out/d_two_pass.jsonl
