#!/bin/bash
# multi-GPU session: strong scaling of BASELINE config 2 (pop = 64 sharded over N ranks) and config 3 (pop = 256), weak scaling for
# reference.  usage: bash scripts/gpu_session_multi.sh N [tag]
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" 2>> gpurun_out/${TAG}_n${N}.err | grep '^{' >> gpurun_out/${TAG}_n${N}.jsonl; }
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_n${N}_gpus.txt
run --steps 4 --warmup 2 --no-cpu-baseline
run --config 3 --iters 10 --steps 3 --warmup 1 --no-cpu-baseline
run --weak --steps 2 --warmup 1 --iters 10 --no-cpu-baseline
STITO_PEER_GATHER=0 run --steps 4 --warmup 2 --no-cpu-baseline
python - <<PY
import json
for ln in open('gpurun_out/${TAG}_n${N}.jsonl'):
    d=json.loads(ln); r=d['roofline']
    print(d['config']['parallelism'][-60:], '|', d['metric'], d['scaling'], 'n_gpus', d['n_gpus'], 'value %.0f e2e %.0f ms/gen %.3f'%(d['value'], d['e2e']['value'], d['ms_per_generation']), {k:round(v,3) for k,v in r['stages_ms_per_generation'].items()}, 'cma %.3f'%d['host_cma_ms_per_generation'], d.get('shard_check'))
PY
tail -5 gpurun_out/${TAG}_n${N}.err
