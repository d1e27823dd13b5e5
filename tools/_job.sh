#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_fullsize.py -x -q > gpurun_out/pytest_s61.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_s61.log | cut -c1-300
for i in 1 0; do AGCF_STAGE2_IMPL=$i timeout 200 python tools/eval_bench.py 2>&1 | sed -n '1p;3p'; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/launches_eval_s61.csv python tools/eval_bench.py > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_eval_s61.csv 2>/dev/null | head -9
