#!/bin/bash
# check of the LTI-compressor kernels: parity tests, ncu --set full of the two kernels, launch list of the dasp-style chain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; T=gpurun_out/${1:-l}
timeout 900 python -m pytest tests -m gpu -q -k "lti" > ${T}_tests.log 2>&1; echo "pytest rc=$?" >> ${T}_tests.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"lti_" -o ${T}_prof_lti -f python scripts/dev_generation.py 16 1 10 mastering-dasp > ${T}_ncu1.log 2>&1
ncu -i ${T}_prof_lti.ncu-rep --page raw --csv > ${T}_prof_lti_raw.csv 2>/dev/null
rm -f ${T}_prof_lti.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${T}_launches_dasp.csv python scripts/dev_generation.py 16 1 10 mastering-dasp > ${T}_ncu2.log 2>&1
timeout 900 python bench.py --chain mastering-dasp --steps 2 --warmup 1 --iters 10 --cpu-sample 2 > ${T}_bench_dasp.jsonl 2>> ${T}_bench.err
tail -3 ${T}_tests.log; grep -E "lti_" ${T}_launches_dasp.csv | awk -F'","' '{print $5, $NF}'
