#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_contrast_engine.py -q > gpurun_out/pytest_s43.log 2>&1
tail -2 gpurun_out/pytest_s43.log | cut -c1-300
timeout 600 python tools/whitebox_bench.py ml-1m 2>/dev/null | tail -1
