#!/bin/bash
mkdir -p gpurun_out
echo default_l16m4; timeout 300 python tools/spmm_variants.py amazon-book 2>&1 | tail -1
echo l16m5; ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_l16m5.so timeout 300 python tools/spmm_variants.py amazon-book 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_propagate.py tests/test_gpu_fused_step.py tests/test_gpu_fullsize.py -q 2>&1 | tail -3
timeout 900 python bench.py --workload amazon-book --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s36_amazon-book.json 2> gpurun_out/bench_s36_amazon-book.err
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_s36_amazon-book.json").read().strip().splitlines()[-1])
r=l["roofline"]
print("amazon-book", "ms/step %.4f"%l["ms_per_step"], "value %.3fM"%(l["value"]/1e6), "e2e %.3fM"%(l["e2e"]["value"]/1e6), "full %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], "l2frac %.3f"%r["l2_gather"]["frac"], r["batch_sparse_launch_ms"], "eval %.3gM users/s"%(l["eval"]["users_per_s"]/1e6))
PY
