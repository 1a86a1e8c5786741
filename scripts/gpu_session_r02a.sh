#!/bin/bash
# round 2, GPU session A: full GPU test suite, default bench, small-population floor, launch list, ncu on the DSP / front-end kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_gpu_tests.log
timeout 600 python bench.py > gpurun_out/a_bench.log 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
for p in 8 16 32; do
  timeout 300 python bench.py --pop $p --steps 2 --warmup 1 --no-cpu-baseline >> gpurun_out/a_pop_sweep.jsonl 2>> gpurun_out/a_bench.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/a_launches.csv python scripts/dev_generation.py 64 1 > gpurun_out/a_ncu1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/a_launches_p8.csv python scripts/dev_generation.py 8 1 > gpurun_out/a_ncu1b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"eq_|compressor|reverb|logmel|tc_conv_first" -o gpurun_out/a_prof_dsp -f python scripts/dev_generation.py 64 1 > gpurun_out/a_ncu2.log 2>&1
ncu -i gpurun_out/a_prof_dsp.ncu-rep --page raw --csv > gpurun_out/a_prof_dsp_raw.csv 2>/dev/null
tail -5 gpurun_out/a_gpu_tests.log; tail -2 gpurun_out/a_bench.log | cut -c1-1500; cat gpurun_out/a_pop_sweep.jsonl | cut -c1-400
