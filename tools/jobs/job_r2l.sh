# 1 GPU: narrow-slice SpMM -- segment length x persistent (dynamic block scheduler) x resident CTAs per SM
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2l; mkdir -p $O
export PYTHONUNBUFFERED=1
for LIB in libagcf.so csrc/build/libagcf_m5.so csrc/build/libagcf_m6.so; do for P in 0 1; do for D in 8 16; do for SEG in 16 32 64; do
  ARLIB_B200_PERSISTENT=$P ARLIB_B200_LIB=$PWD/arlib_b200/$LIB SPMM_D=$D ARLIB_B200_SEGMENT=$SEG timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/$(basename $LIB) persistent=$P seg=$SEG /" >> $O/spmm_narrow_matrix.txt
done; done; done; done
cat $O/spmm_narrow_matrix.txt
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; head -3 $O/eval_bench.txt
timeout 600 python -m pytest tests/test_gpu_topk.py -x -q -m gpu > $O/tests_topk.log 2>&1; tail -2 $O/tests_topk.log
