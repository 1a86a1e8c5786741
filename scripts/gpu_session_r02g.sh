#!/bin/bash
# round 2, session G: per-layer chunk lengths, amax published once per kernel, lazy timing, multi-microbatch fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; md5sum st_ito_b200/libstito.so
timeout 1200 python -m pytest tests -m gpu -x -q -s -k "second_weight or many_microbatches or config2_full or fp16_range" > gpurun_out/g_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline >> gpurun_out/g_bench.jsonl 2>> gpurun_out/g_bench.err
STITO_TC_CHUNK128=2 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline >> gpurun_out/g_bench_c128_2.jsonl 2>> gpurun_out/g_bench.err
STITO_TC_CHUNK128=4 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline >> gpurun_out/g_bench_c128_4.jsonl 2>> gpurun_out/g_bench.err
timeout 300 python bench.py --pop 8 --steps 3 --warmup 2 --no-cpu-baseline >> gpurun_out/g_bench_p8.jsonl 2>> gpurun_out/g_bench.err
grep -E "passed|failed|FAILED|rc=|fixture" gpurun_out/g_gpu_tests.log | tail -12 | cut -c1-330
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/g_*.jsonl')):
    for ln in open(f):
        if not ln.startswith('{'): continue
        d=json.loads(ln); r=d['roofline']
        print(f.split('/')[-1], d['metric'][:36], 'value %.0f ms/gen %.3f'%(d['value'],d['ms_per_generation']), {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()}, 'cma %.3f frac %.3f'%(d['host_cma_ms_per_generation'], r['frac']), [round(x,3) for x in r['ms_per_layer']])
PY
tail -3 gpurun_out/g_bench.err
