#!/bin/bash
# where does the cluster-split Freeverb stop paying?  populations 10 / 12 / 14 with and without it (real host loop, 8 generations)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; T=gpurun_out/${1:-th}
for p in 10 12 14; do
  timeout 300 python bench.py --pop $p --steps 1 --warmup 1 --iters 8 --no-cpu-baseline >> ${T}_split.jsonl 2>> ${T}.err
  STITO_REVERB_SPLIT=0 timeout 300 python bench.py --pop $p --steps 1 --warmup 1 --iters 8 --no-cpu-baseline >> ${T}_core.jsonl 2>> ${T}.err
done
python - "$T" <<'PY'
import json, sys
for k in ("split", "core"):
    for ln in open(f"{sys.argv[1]}_{k}.jsonl"):
        if ln.startswith("{"):
            d = json.loads(ln); s = d["roofline"]["stages_ms_per_generation"]
            print(k, d["config"]["workload"].split("pop=")[1].split(",")[0], "value %.0f ms/gen %.3f dsp %.3f" % (d["value"], d["ms_per_generation"], s["ms_dsp"]))
PY
