#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s33.log 2>&1; tail -2 gpurun_out/smoke_s33.log
for wl in amazon-book yelp2018 ml-1m; do
timeout 900 python bench.py --workload $wl --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s33_$wl.json 2> gpurun_out/bench_s33_$wl.err
python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_s33_$wl.json").read().strip().splitlines()[-1])
    r=l["roofline"]
    print("$wl", "ms/step %.4f"%l["ms_per_step"], "value %.3fM"%(l["value"]/1e6), "e2e %.3fM"%(l["e2e"]["value"]/1e6), "full %.4f"%r["avg_launch_ms"], "frac %.3f"%r["frac"], "l2frac %.3f"%r["l2_gather"]["frac"], r["batch_sparse_launch_ms"], "eval %.3gM users/s"%(l["eval"]["users_per_s"]/1e6), "step_frac %.3f"%r["step_frac_of_hbm_roofline"])
except Exception as e:
    print("$wl failed", e)
PY
tail -2 gpurun_out/bench_s33_$wl.err
done
ARLIB_B200_ALPHA=0.8 timeout 600 python bench.py --alpha 0.8 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s33_gowalla_a08.json 2> gpurun_out/bench_s33_gowalla_a08.err
python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_s33_gowalla_a08.json").read().strip().splitlines()[-1])
    r=l["roofline"]
    print("gowalla a=0.8", "ms/step %.4f"%l["ms_per_step"], "value %.3fM"%(l["value"]/1e6), "full %.4f"%r["avg_launch_ms"], r["batch_sparse_launch_ms"])
except Exception as e:
    print("a08 failed", e)
PY
