#!/bin/bash
# scratch job for gpurun (overwritten per call)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_step.py -x -q > gpurun_out/pytest_s20_fused.log 2>&1
tail -5 gpurun_out/pytest_s20_fused.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fused_step.py > gpurun_out/pytest_s20.log 2>&1
tail -3 gpurun_out/pytest_s20.log
for cfg in "0 0 0" "1 0 0" "1 1 0" "1 1 1"; do
  set -- $cfg
  ARLIB_B200_WORKLISTS=$1 ARLIB_B200_FUSE_ADAM=$2 ARLIB_B200_PDL=$3 timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s20_w$1a$2p$3.json 2> gpurun_out/bench_s20_w$1a$2p$3.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_s20_w$1a$2p$3.json").read().strip().splitlines()[-1])
    r=l["roofline"]
    print("cfg $cfg", "ms/step %.4f"%l["ms_per_step"], "value %.3fM"%(l["value"]/1e6), "e2e %.3fM"%(l["e2e"]["value"]/1e6), "full %.4f"%r["avg_launch_ms"], r["batch_sparse_launch_ms"])
except Exception as e:
    print("cfg $cfg failed", e)
PY
done
