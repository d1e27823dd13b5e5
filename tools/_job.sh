#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_s29.log 2>&1
tail -4 gpurun_out/pytest_s29.log
timeout 900 python bench.py > gpurun_out/bench_s29_n1.json 2> gpurun_out/bench_s29_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_s29_ref.json 2> gpurun_out/bench_s29_ref.err
python - <<PY
import json
for f in ("gpurun_out/bench_s29_n1.json", "gpurun_out/bench_s29_ref.json"):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g"%l["value"], "ms/step %.4f"%l["ms_per_step"], "e2e", l.get("e2e",{}).get("value"), "roofline", {k:l.get("roofline",{}).get(k) for k in ("achieved","frac","avg_launch_ms","batch_sparse_launch_ms")}, "eval", l.get("eval",{}).get("users_per_s"), "clocks", l.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 python tools/spmm_variants.py > gpurun_out/spmm_variants_s29.txt 2>&1; tail -1 gpurun_out/spmm_variants_s29.txt
timeout 600 python tools/whitebox_bench.py ml-1m > gpurun_out/whitebox_s29.jsonl 2> gpurun_out/whitebox_s29.err; cat gpurun_out/whitebox_s29.jsonl; tail -3 gpurun_out/whitebox_s29.err
CONTRAST_STEPS=500 timeout 900 python tools/contrast_bench.py yelp2018 100 XSimGCL,SimGCL > gpurun_out/contrast_s29.jsonl 2> gpurun_out/contrast_s29.err; cat gpurun_out/contrast_s29.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_train_s29.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_s29.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_colmask -s 20 -c 2 -o gpurun_out/colmask_s29 -f python tools/spmm_variants.py > gpurun_out/ncu_colmask_s29.log 2>&1
CONTRAST_STEPS=5 ncu --set full --clock-control none --import-source on -k regex:nce_bwd_kernel -s 8 -c 2 -o gpurun_out/nce_bwd_s29 -f python tools/contrast_bench.py yelp2018 10 none > gpurun_out/ncu_nce_s29.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 30 -c 2 -o gpurun_out/spmm_s29 -f python tools/spmm_variants.py > gpurun_out/ncu_spmm_s29.log 2>&1
for r in colmask_s29 nce_bwd_s29 spmm_s29; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; done
ls -la gpurun_out/*s29*
