#!/bin/bash
# round 2, GPU session B: new tests (conv reverb, margins), margins probe matrix, config-4 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "second_weight or noise_shaped or conv_reverb or config4 or fp16_range or config2_full" > gpurun_out/b_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_gpu_tests.log
for fx in xavier heavy; do for comp in 0 0.25; do for chunk in 4 2; do
  STITO_TC_COMP=$comp STITO_TC_CHUNK=$chunk timeout 300 python tests/dev/dev_margins2.py $fx 2>/dev/null | tail -1 >> gpurun_out/b_margins.jsonl
done; done; done
DEV_PRECISION=0 timeout 300 python tests/dev/dev_margins2.py heavy 2>/dev/null | tail -1 >> gpurun_out/b_margins.jsonl
timeout 900 python bench.py --config 4 --steps 2 --warmup 1 --iters 5 --cpu-sample 1 > gpurun_out/b_bench_c4.log 2> gpurun_out/b_bench_c4.err; echo "rc=$?" >> gpurun_out/b_bench_c4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/b_launches_c4.csv python scripts/dev_generation.py 16 1 30 mastering-conv > gpurun_out/b_ncu1.log 2>&1
tail -5 gpurun_out/b_gpu_tests.log; cat gpurun_out/b_margins.jsonl; tail -c 1500 gpurun_out/b_bench_c4.log; tail -3 gpurun_out/b_bench_c4.err
