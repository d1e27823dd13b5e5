#!/bin/bash
# scratch job for gpurun (overwritten per call)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s21.log 2>&1
tail -3 gpurun_out/pytest_s21.log
for cfg in "0" "1"; do
  ARLIB_B200_PERSISTENT=$cfg timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s21_p$cfg.json 2> gpurun_out/bench_s21_p$cfg.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_s21_p$cfg.json").read().strip().splitlines()[-1])
    r=l["roofline"]
    print("persistent $cfg", "ms/step %.4f"%l["ms_per_step"], "value %.3fM"%(l["value"]/1e6), "e2e %.3fM"%(l["e2e"]["value"]/1e6), "full %.4f"%r["avg_launch_ms"], r["batch_sparse_launch_ms"], "eval", l["eval"]["users_per_s"])
except Exception as e:
    print("cfg $cfg failed", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_train_s21.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_s21.log 2>&1
python tools/launch_summary.py gpurun_out/launches_train_s21.csv | head -30
