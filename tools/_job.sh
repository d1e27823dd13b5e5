#!/bin/bash
# scratch job for gpurun (overwritten per call)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_contrast.py tests/test_gpu_train.py tests/test_gpu_fused_step.py -x -q > gpurun_out/pytest_s22.log 2>&1
tail -15 gpurun_out/pytest_s22.log
for seg in 16 32 64 128; do
  ARLIB_B200_WL_SEGMENT=$seg timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s22_wl$seg.json 2> gpurun_out/bench_s22_wl$seg.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_s22_wl$seg.json").read().strip().splitlines()[-1])
    r=l["roofline"]
    print("wl_segment $seg", "ms/step %.4f"%l["ms_per_step"], "value %.3fM"%(l["value"]/1e6), "e2e %.3fM"%(l["e2e"]["value"]/1e6), "full %.4f"%r["avg_launch_ms"], r["batch_sparse_launch_ms"])
except Exception as e:
    print("cfg $seg failed", e)
PY
done
timeout 900 python tools/contrast_bench.py yelp2018 100 > gpurun_out/contrast_s22.jsonl 2> gpurun_out/contrast_s22.err
cat gpurun_out/contrast_s22.jsonl; tail -5 gpurun_out/contrast_s22.err
