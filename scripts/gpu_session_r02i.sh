#!/bin/bash
# round 2, session I: conv-reverb FFT with shared-memory twiddles; chunk128 = 4 default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; md5sum st_ito_b200/libstito.so
timeout 1500 python -m pytest tests -m gpu -q -k "noise_shaped or conv_reverb or config4 or second_weight" > gpurun_out/i_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_gpu_tests.log
timeout 900 python bench.py --config 4 --steps 1 --warmup 1 --iters 5 --cpu-sample 1 >> gpurun_out/i_bench_c4.jsonl 2>> gpurun_out/i_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/i_launches_c4.csv python scripts/dev_generation.py 16 1 30 mastering-conv > gpurun_out/i_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"crv_" -o gpurun_out/i_prof_crv -f python scripts/dev_generation.py 16 1 30 mastering-conv > gpurun_out/i_ncu6.log 2>&1
ncu -i gpurun_out/i_prof_crv.ncu-rep --page raw --csv > gpurun_out/i_prof_crv_raw.csv 2>/dev/null; rm -f gpurun_out/i_prof_crv.ncu-rep
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline >> gpurun_out/i_bench.jsonl 2>> gpurun_out/i_bench.err
grep -E "passed|failed|FAILED|rc=" gpurun_out/i_gpu_tests.log | tail -6 | cut -c1-300
python scripts/launch_summary.py gpurun_out/i_launches_c4.csv 34 2>/dev/null | head -9
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/i_*.jsonl')):
    for ln in open(f):
        if not ln.startswith('{'): continue
        d=json.loads(ln); r=d['roofline']
        print(f.split('/')[-1], d['metric'][:36], 'value %.0f ms/gen %.3f'%(d['value'],d['ms_per_generation']), {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()}, 'frac %.3f'%r['frac'], d.get('parity',{}).get('max_rel_err'))
PY
