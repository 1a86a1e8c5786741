#!/bin/bash
# round 2, GPU session C: full GPU suite (weight scaling fix, chunk 2, native CMA, comp->reverb streaming), pop sweep, streaming on/off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/c_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_gpu_tests.log
for p in 8 16 32 64; do
  timeout 300 python bench.py --pop $p --steps 2 --warmup 1 --no-cpu-baseline >> gpurun_out/c_pop_sweep.jsonl 2>> gpurun_out/c_bench.err
done
STITO_DSP_STREAMING=0 timeout 300 python bench.py --pop 8 --steps 2 --warmup 1 --no-cpu-baseline >> gpurun_out/c_pop_sweep_nostream.jsonl 2>> gpurun_out/c_bench.err
STITO_TC_CHUNK=4 timeout 300 python bench.py --pop 64 --steps 2 --warmup 1 --no-cpu-baseline >> gpurun_out/c_pop64_chunk4.jsonl 2>> gpurun_out/c_bench.err
grep -E "passed|failed|FAILED|rc=" gpurun_out/c_gpu_tests.log | tail -12
python - <<'PY'
import json
for f in ['gpurun_out/c_pop_sweep.jsonl','gpurun_out/c_pop_sweep_nostream.jsonl','gpurun_out/c_pop64_chunk4.jsonl']:
    for ln in open(f):
        if not ln.startswith('{'): continue
        d=json.loads(ln); r=d['roofline']
        print(f.split('/')[-1], d['metric'], 'value %.0f ms/gen %.3f'%(d['value'],d['ms_per_generation']), {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()}, 'cma %.3f'%d['host_cma_ms_per_generation'], 'frac %.3f'%r['frac'])
PY
