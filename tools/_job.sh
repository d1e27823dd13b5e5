#!/bin/bash
mkdir -p gpurun_out
N=8
ARLIB_B200_DIST=dshard timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s44_n8_dshard.json 2> gpurun_out/bench_s44_n8_dshard.err
python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_s44_n8_dshard.json").read().strip().splitlines()[-1])
    r=l["roofline"]
    print("dshard N=8", "ms/step %.4f"%l["ms_per_step"], "value %.3fM"%(l["value"]/1e6), "e2e %.3fM"%(l["e2e"]["value"]/1e6), "full %.4f"%r["avg_launch_ms"], r["batch_sparse_launch_ms"], "eval", l["eval"]["users_per_s"])
except Exception as e:
    print("failed", e)
PY
tail -3 gpurun_out/bench_s44_n8_dshard.err
