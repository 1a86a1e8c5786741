#!/bin/bash
# quick check of a DSP change: reverb-related parity tests, small-population bench, default bench, launch list at P=8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; T=gpurun_out/${1:-s}
timeout 900 python -m pytest tests -m gpu -q -k "config2_full or single_plugin or ragged or other_sample_rates or properties_at_full or split or stream or reverb or lti" > ${T}_tests.log 2>&1; echo "pytest rc=$?" >> ${T}_tests.log
for p in 8 16; do
  timeout 600 python bench.py --pop $p --steps 2 --warmup 1 --iters 10 --no-cpu-baseline >> ${T}_pop.jsonl 2>> ${T}_bench.err
done
timeout 600 python bench.py --steps 3 --warmup 1 --cpu-sample 2 > ${T}_bench.jsonl 2>> ${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${T}_launches_p8.csv python scripts/dev_generation.py ${2:-8} 1 > ${T}_ncu.log 2>&1
tail -3 ${T}_tests.log; tail -3 ${T}_bench.err
python - "$T" <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'_*.jsonl')):
    for ln in open(f):
        if not ln.startswith('{'): continue
        d=json.loads(ln); r=d['roofline']
        print(f.split('/')[-1], 'value %.0f e2e %.0f ms/gen %.3f'%(d['value'],d['e2e']['value'],d['ms_per_generation']), {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()}, d.get('parity',{}).get('max_rel_err'))
PY
grep -E "reverb|compressor|eq_" ${T}_launches_p8.csv | awk -F'","' '{print $5, $NF}' | head -12
