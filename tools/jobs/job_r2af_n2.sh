# 2 GPUs: parity test only (user-sharded evaluation goes through the per-chunk workspaces)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2af; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -s > $O/dist_test.log 2>&1; echo "dist rc=$?" >> $O/dist_test.log; tail -4 $O/dist_test.log
