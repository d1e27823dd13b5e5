# 1 GPU: timing experiment -- vector index loads on (pretend-)aligned item starts for narrow slices; eval after the match.any revert
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2k; mkdir -p $O
export PYTHONUNBUFFERED=1
for LIB in libagcf_exp.so libagcf_exp3.so; do for D in 8 16 32; do for SEG in 32 64; do
  ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/$LIB SPMM_D=$D ARLIB_B200_SEGMENT=$SEG timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/$LIB seg=$SEG /" >> $O/spmm_vecidx_experiment.txt
done; done; done
cat $O/spmm_vecidx_experiment.txt
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; head -3 $O/eval_bench.txt
timeout 600 python -m pytest tests/test_gpu_topk.py -x -q -m gpu > $O/tests_topk.log 2>&1; tail -2 $O/tests_topk.log
