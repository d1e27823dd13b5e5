#!/bin/bash
# scratch job for gpurun (overwritten per call)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_contrast.py tests/test_gpu_contrast_engine.py -x -q > gpurun_out/pytest_s23a.log 2>&1
tail -25 gpurun_out/pytest_s23a.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_contrast.py --deselect tests/test_gpu_contrast_engine.py > gpurun_out/pytest_s23b.log 2>&1
tail -15 gpurun_out/pytest_s23b.log
timeout 900 python tools/contrast_bench.py yelp2018 100 XSimGCL,SimGCL > gpurun_out/contrast_s23.jsonl 2> gpurun_out/contrast_s23.err
cat gpurun_out/contrast_s23.jsonl; tail -5 gpurun_out/contrast_s23.err
