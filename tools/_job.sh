#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_s50.log 2>&1
tail -3 gpurun_out/pytest_s50.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_s50_n1.json 2> gpurun_out/bench_s50_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_s50_ref.json 2> gpurun_out/bench_s50_ref.err
python - <<PY
import json
for f in ("gpurun_out/bench_s50_n1.json", "gpurun_out/bench_s50_ref.json"):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        r=l.get("roofline") or {}
        print(f, "value %.4g"%l["value"], "ms/step %.4f"%l["ms_per_step"], "e2e", l.get("e2e",{}).get("value"), "frac", r.get("frac"), "full_ms", r.get("avg_launch_ms"), "eval", (l.get("eval") or {}).get("users_per_s"), "launches", l.get("gpu_launches"), "clocks", l.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
