# 1 GPU: warp-granular dynamic scheduler (ARLIB_B200_PERSISTENT=1) on narrow slices, partitions and the full-width launch; eval
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2m; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_propagate.py tests/test_gpu_topk.py -x -q -m gpu > $O/tests.log 2>&1; tail -2 $O/tests.log
for P in 0 1; do for D in 8 16 32; do for SEG in 32 64; do
  ARLIB_B200_PERSISTENT=$P SPMM_D=$D ARLIB_B200_SEGMENT=$SEG timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/warp-sched=$P seg=$SEG /" >> $O/spmm_warp_sched.txt
done; done
ARLIB_B200_PERSISTENT=$P timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/warp-sched=$P /" >> $O/spmm_warp_sched.txt
ARLIB_B200_PERSISTENT=$P SPMM_PART=8 timeout 300 python tools/spmm_variants.py 2>&1 | tail -8 | sed "s/^/warp-sched=$P /" >> $O/spmm_warp_sched.txt
done
cat $O/spmm_warp_sched.txt
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; head -3 $O/eval_bench.txt
