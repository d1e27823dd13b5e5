#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_topk.py -x -q > gpurun_out/pytest_s54.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_s54.log | cut -c1-300
timeout 200 python tools/eval_bench.py 2>&1 | sed -n 1,1p
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv -k regex:group_max_tc -s 2 -c 2 python tools/eval_bench.py 2>/dev/null | grep "group_max_tc" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | head -4
