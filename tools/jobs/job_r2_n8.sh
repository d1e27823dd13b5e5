# 8 GPUs (one NVSwitch box): 8-rank parity, the C2 scaling points 8 / 4, C5a in both layouts, C5b row-partitioned
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2n8; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -s > $O/dist_test_8ranks.log 2>&1; echo "dist rc=$?" >> $O/dist_test_8ranks.log; tail -4 $O/dist_test_8ranks.log
grep -q "1 passed" $O/dist_test_8ranks.log || { echo "8-rank parity failed: not spending the box on the bench lines"; tail -40 $O/dist_test_8ranks.log; exit 1; }
timeout 500 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n8_dshard.json 2> $O/bench_n8_dshard.err; echo rc=$?
timeout 500 $TR --nproc-per-node 4 --master-port 29712 bench.py --gpus 4 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n4_dshard.json 2> $O/bench_n4_dshard.err; echo rc=$?
timeout 500 $TR --nproc-per-node 8 --master-port 29713 bench.py --gpus 8 --workload amazon-book --steps 200 --warmup 5 --no-cpu-baseline > $O/amazon_n8_dshard.json 2> $O/amazon_n8_dshard.err; echo rc=$?
ARLIB_B200_DIST=rows timeout 500 $TR --nproc-per-node 8 --master-port 29714 bench.py --gpus 8 --workload amazon-book --steps 200 --warmup 5 --no-cpu-baseline > $O/amazon_n8_rows.json 2> $O/amazon_n8_rows.err; echo rc=$?
timeout 900 $TR --nproc-per-node 8 --master-port 29715 bench.py --gpus 8 --workload c5b --steps 10 --warmup 3 > $O/c5b_n8_rows.json 2> $O/c5b_n8_rows.err; echo rc=$?
ARLIB_B200_SEGMENT=64 ARLIB_B200_WL_SEGMENT=64 timeout 500 $TR --nproc-per-node 8 --master-port 29717 bench.py --gpus 8 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n8_dshard_seg64.json 2> $O/bench_n8_dshard_seg64.err; echo rc=$?
ARLIB_B200_DIST=rows timeout 500 $TR --nproc-per-node 8 --master-port 29716 bench.py --gpus 8 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n8_rows.json 2> $O/bench_n8_rows.err; echo rc=$?
head -c 1200 $O/bench_n8_dshard.json; tail -2 $O/*.err
