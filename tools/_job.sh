#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_contrast_engine.py -q -s > gpurun_out/pytest_s26.log 2>&1
grep -v Warn gpurun_out/pytest_s26.log | grep "engine vs\|passed\|failed\|Error\|assert" | head -20
