#!/bin/bash
# minimal multi-GPU confirmation: BASELINE's metric only (pop = 64 sharded over N ranks, peer-memory gather).  usage: N [tag]
cd "$(dirname "$0")/.."
N=${1:-8}; TAG=${2:-mm}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 4 --warmup 2 --no-cpu-baseline 2>> gpurun_out/${TAG}_n${N}.err | grep '^{' >> gpurun_out/${TAG}_n${N}.jsonl
python - <<PY
import json
for ln in open('gpurun_out/${TAG}_n${N}.jsonl'):
    d=json.loads(ln); r=d['roofline']
    print(d['config']['parallelism'][-60:], '|', d['scaling'], 'n_gpus', d['n_gpus'], 'value %.0f e2e %.0f ms/gen %.3f'%(d['value'], d['e2e']['value'], d['ms_per_generation']), {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()}, 'cma %.3f'%d['host_cma_ms_per_generation'], d.get('shard_check'))
PY
tail -3 gpurun_out/${TAG}_n${N}.err
