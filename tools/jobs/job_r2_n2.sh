# 2 GPUs: multi-GPU parity + the two layouts at C2, small scale-stress in rows mode
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2n2; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -s > $O/dist_test.log 2>&1; echo "dist rc=$?" >> $O/dist_test.log; tail -5 $O/dist_test.log
timeout 900 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -4 $O/all_tests.log
timeout 600 $TR --master-port 29611 bench.py --gpus 2 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n2_dshard.json 2> $O/bench_n2_dshard.err; echo rc=$?
ARLIB_B200_DIST=rows timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n2_rows.json 2> $O/bench_n2_rows.err; echo rc=$?
AGCF_SPMM_COOP=0 timeout 600 $TR --master-port 29613 bench.py --gpus 2 --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n2_dshard_nocoop.json 2> $O/bench_n2_dshard_nocoop.err; echo rc=$?
timeout 900 $TR --master-port 29614 bench.py --gpus 2 --workload c5b-small --steps 10 --warmup 3 > $O/c5b_small_n2_rows.json 2> $O/c5b_small_n2_rows.err; echo rc=$?
timeout 600 $TR --master-port 29615 bench.py --gpus 2 --workload amazon-book --steps 200 --warmup 5 --no-cpu-baseline > $O/amazon_n2_dshard.json 2> $O/amazon_n2_dshard.err; echo rc=$?
head -c 1500 $O/bench_n2_dshard.json; tail -3 $O/*.err
