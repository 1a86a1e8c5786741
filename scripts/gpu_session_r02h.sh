#!/bin/bash
# round 2, session H: margins vs chunk length of the N=128 layers; remaining tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; md5sum st_ito_b200/libstito.so
timeout 1200 python -m pytest tests -m gpu -q -s -k "second_weight or many_microbatches or config2_full or fp16_range" > gpurun_out/h_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_gpu_tests.log
for c128 in 2 3 4; do for comp in 0 0.25; do for fx in xavier heavy; do
  STITO_TC_COMP=$comp STITO_TC_CHUNK128=$c128 timeout 300 python tests/dev/dev_margins2.py $fx 2>/dev/null | tail -1 | sed "s/^{/{\"chunk128\": $c128, /" >> gpurun_out/h_margins.jsonl
done; done; done
grep -E "passed|failed|FAILED|rc=" gpurun_out/h_gpu_tests.log | tail -6 | cut -c1-300
cat gpurun_out/h_margins.jsonl | cut -c1-330
