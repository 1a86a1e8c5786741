#!/bin/bash
# round 2, evidence session: full GPU suite, smoke, default bench + reference arm, population sweep (config 5), launch lists and
# ncu captures behind profiles/r02_*
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; T=gpurun_out/f
md5sum st_ito_b200/libstito.so > ${T}_md5.txt
timeout 2400 python -m pytest tests -m gpu -q > ${T}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> ${T}_gpu_tests.log
timeout 600 python __graft_entry__.py smoke > ${T}_smoke.log 2>&1; echo "smoke rc=$?" >> ${T}_smoke.log
timeout 900 python bench.py > ${T}_bench.jsonl 2> ${T}_bench.err; echo "bench rc=$?" >> ${T}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > ${T}_bench_ref.jsonl 2>> ${T}_bench.err
for p in 8 16 32 64 128 256 512 1024; do
  timeout 600 python bench.py --pop $p --steps 1 --warmup 1 --iters 8 --no-cpu-baseline >> ${T}_pop_sweep.jsonl 2>> ${T}_bench.err
done
timeout 900 python bench.py --config 4 --steps 1 --warmup 1 --iters 5 --cpu-sample 1 >> ${T}_bench_c4.jsonl 2>> ${T}_bench.err
timeout 900 python bench.py --config 1 --steps 3 --warmup 1 --cpu-sample 2 >> ${T}_bench_c1.jsonl 2>> ${T}_bench.err
timeout 900 python bench.py --chain mastering-dasp --steps 2 --warmup 1 --iters 10 --cpu-sample 2 >> ${T}_bench_dasp.jsonl 2>> ${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${T}_launches_p64.csv python scripts/dev_generation.py 64 1 > ${T}_ncu1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${T}_launches_p8.csv python scripts/dev_generation.py 8 1 > ${T}_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${T}_launches_c4.csv python scripts/dev_generation.py 16 1 30 mastering-conv > ${T}_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv3x3|wino" -o ${T}_prof_conv -f python scripts/dev_generation.py 64 1 > ${T}_ncu4.log 2>&1
ncu -i ${T}_prof_conv.ncu-rep --page raw --csv > ${T}_prof_conv_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"eq_|compressor|reverb|logmel|tc_conv_first" -o ${T}_prof_dsp -f python scripts/dev_generation.py 64 1 > ${T}_ncu5.log 2>&1
ncu -i ${T}_prof_dsp.ncu-rep --page raw --csv > ${T}_prof_dsp_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"crv_" -o ${T}_prof_crv -f python scripts/dev_generation.py 16 1 30 mastering-conv > ${T}_ncu6.log 2>&1
ncu -i ${T}_prof_crv.ncu-rep --page raw --csv > ${T}_prof_crv_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"lti_" -o ${T}_prof_lti -f python scripts/dev_generation.py 16 1 10 mastering-dasp > ${T}_ncu8.log 2>&1
ncu -i ${T}_prof_lti.ncu-rep --page raw --csv > ${T}_prof_lti_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${T}_launches_dasp.csv python scripts/dev_generation.py 16 1 10 mastering-dasp > ${T}_ncu9.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"reverb_split|compressor" -o ${T}_prof_p8 -f python scripts/dev_generation.py 8 1 > ${T}_ncu7.log 2>&1
ncu -i ${T}_prof_p8.ncu-rep --page raw --csv > ${T}_prof_p8_raw.csv 2>/dev/null
rm -f ${T}_prof_conv.ncu-rep ${T}_prof_dsp.ncu-rep ${T}_prof_crv.ncu-rep ${T}_prof_p8.ncu-rep ${T}_prof_lti.ncu-rep
grep -E "passed|failed|FAILED|rc=" ${T}_gpu_tests.log | tail -5; tail -2 ${T}_smoke.log; tail -3 ${T}_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f_*.jsonl')):
    for ln in open(f):
        if not ln.startswith('{'): continue
        d=json.loads(ln)
        if d.get('impl')=='reference': print(f, 'REF', d['value'], d['cpu_baseline'].get('value_serial')); continue
        r=d['roofline']
        print(f.split('/')[-1], d['metric'][:40], 'value %.0f e2e %.0f ms/gen %.3f'%(d['value'],d['e2e']['value'],d['ms_per_generation']), {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()}, 'cma %.3f frac %.3f'%(d['host_cma_ms_per_generation'], r['frac']), d.get('parity',{}).get('max_rel_err'), d.get('parity',{}).get('argsort_equal'))
PY
