#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_s63.log 2>&1
tail -3 gpurun_out/pytest_s63.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_s63_n1.json 2> gpurun_out/bench_s63_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_s63_ref.json 2> gpurun_out/bench_s63_ref.err
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_s63_n1.json").read().strip().splitlines()[-1])
r=l["roofline"]; e=l["eval"]
print("value %.4g"%l["value"], "ms/step %.4f"%l["ms_per_step"], "e2e %.4g"%l["e2e"]["value"], "frac %.3f"%r["frac"], "eval %.4g users/s %.3f ms"%(e["users_per_s"], e["ms"]), l["clocks"])
l=json.loads(open("gpurun_out/bench_s63_ref.json").read().strip().splitlines()[-1])
print("ref value %.4g"%l["value"], "eval", l["eval"]["users_per_s"])
PY
