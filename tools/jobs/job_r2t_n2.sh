# 2 GPUs: multi-GPU parity + both layouts at C2 after the late kernel changes
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2t; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -s > $O/dist_test.log 2>&1; echo "dist rc=$?" >> $O/dist_test.log; tail -5 $O/dist_test.log
timeout 600 $TR --master-port 29611 bench.py --gpus 2 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n2_dshard.json 2> $O/bench_n2_dshard.err; echo rc=$?
ARLIB_B200_DIST=rows timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n2_rows.json 2> $O/bench_n2_rows.err; echo rc=$?
for f in bench_n2_dshard bench_n2_rows; do python -c "
import json;d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['roofline']['avg_launch_ms'])"; done
for f in $O/*.err; do tail -2 $f; done
