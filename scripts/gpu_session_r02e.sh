#!/bin/bash
# round 2, GPU session E: split reverb with bulk-copy DSMEM hand-off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; md5sum st_ito_b200/libstito.so
timeout 900 python -m pytest tests -m gpu -x -q -k "config2_full or single_plugin or ragged or other_sample_rates or properties_at_full" > gpurun_out/e_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_gpu_tests.log
for p in 8 12; do
  timeout 300 python bench.py --pop $p --steps 2 --warmup 1 --no-cpu-baseline >> gpurun_out/e_pop_sweep.jsonl 2>> gpurun_out/e_bench.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/e_launches_p8.csv python scripts/dev_generation.py 8 1 > gpurun_out/e_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"reverb_split" -o gpurun_out/e_prof_split -f python scripts/dev_generation.py 8 1 > gpurun_out/e_ncu2.log 2>&1
ncu -i gpurun_out/e_prof_split.ncu-rep --page source --csv > gpurun_out/e_split_src.csv 2>/dev/null
grep -E "passed|failed|FAILED|rc=" gpurun_out/e_gpu_tests.log | tail -5
python - <<'PY'
import json
for f in ['e_pop_sweep.jsonl']:
    for ln in open('gpurun_out/'+f):
        if not ln.startswith('{'): continue
        d=json.loads(ln); r=d['roofline']
        print(f, d['metric'][:38], 'ms/gen %.3f'%d['ms_per_generation'], {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()})
PY
python scripts/launch_summary.py gpurun_out/e_launches_p8.csv 34 2>/dev/null | head -6
python scripts/ncu_top.py gpurun_out/e_split_src.csv 24 | cut -c1-220
tail -3 gpurun_out/e_bench.err
