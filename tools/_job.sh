#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_contrast_engine.py tests/test_gpu_propagate.py -q > gpurun_out/pytest_s28.log 2>&1
tail -8 gpurun_out/pytest_s28.log
CONTRAST_STEPS=500 timeout 600 python tools/contrast_bench.py yelp2018 10 none > gpurun_out/contrast_s28.jsonl 2> gpurun_out/contrast_s28.err
cat gpurun_out/contrast_s28.jsonl; tail -3 gpurun_out/contrast_s28.err
