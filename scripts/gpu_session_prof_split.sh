#!/bin/bash
# source-level ncu capture of the cluster-split Freeverb at P = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; T=gpurun_out/${1:-sp}
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"reverb_split" -o ${T}_prof -f python scripts/dev_generation.py 8 1 > ${T}_ncu.log 2>&1
ncu -i ${T}_prof.ncu-rep --page source --csv > ${T}_source.csv 2>/dev/null
ncu -i ${T}_prof.ncu-rep --page raw --csv > ${T}_raw.csv 2>/dev/null
rm -f ${T}_prof.ncu-rep
python scripts/ncu_top.py ${T}_source.csv 60 | tail -70
