# 8 GPUs: final C2 and Amazon-book lines (column-sharded layout) + the 8-rank parity test
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2y; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29631 bench.py --gpus 8 --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_n8_dshard.json 2> $O/bench_n8_dshard.err; echo rc=$?
timeout 400 $TR --master-port 29632 bench.py --gpus 8 --workload amazon-book --steps 200 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_amazon_n8_dshard.json 2> $O/bench_amazon_n8_dshard.err; echo rc=$?
for f in bench_n8_dshard bench_amazon_n8_dshard; do python -c "
import json;d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['roofline']['avg_launch_ms'])"; done
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -s > $O/dist_test_8ranks.log 2>&1; tail -3 $O/dist_test_8ranks.log
