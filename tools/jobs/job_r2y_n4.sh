# 4 GPUs: final C2 line (column-sharded layout)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2y; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29621 bench.py --gpus 4 --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_n4_dshard.json 2> $O/bench_n4_dshard.err; echo rc=$?
python -c "
import json;d=json.loads(open('$O/bench_n4_dshard.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['roofline']['avg_launch_ms'])"
