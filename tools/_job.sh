#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_step.py tests/test_gpu_propagate.py -x -q > gpurun_out/pytest_s31.log 2>&1
tail -3 gpurun_out/pytest_s31.log
for d in 64 32 16 8; do SPMM_D=$d timeout 300 python tools/spmm_variants.py 2>&1 | tail -1; done
timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s31_n1.json 2> gpurun_out/bench_s31_n1.err
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_s31_n1.json").read().strip().splitlines()[-1])
r=l["roofline"]
print("N=1", "ms/step %.4f"%l["ms_per_step"], "value %.3fM"%(l["value"]/1e6), "e2e %.3fM"%(l["e2e"]["value"]/1e6), "full %.4f (eager %.4f)"%(r["avg_launch_ms"], r["avg_launch_ms_eager_single"]), "frac %.3f"%r["frac"], r["batch_sparse_launch_ms"])
PY
