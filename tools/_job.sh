#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_topk.py -x -q > gpurun_out/pytest_s55.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_s55.log | cut -c1-300
for i in 1 0; do AGCF_STAGE2_IMPL=$i timeout 200 python tools/eval_bench.py 2>&1 | sed -n 1,1p; done
