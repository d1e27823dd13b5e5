# 1 GPU: CTA size of the plain SpMM launches (finer CTA turnover at equal warps per SM)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q; mkdir -p $O
export PYTHONUNBUFFERED=1
for T in 256 128 64; do
  AGCF_SPMM_THREADS=$T timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/threads=$T /" >> $O/spmm_threads.txt
  AGCF_SPMM_THREADS=$T timeout 300 python tools/spmm_variants.py amazon-book 2>&1 | tail -1 | sed "s/^/threads=$T /" >> $O/spmm_threads.txt
  for D in 8 16 32; do
  AGCF_SPMM_THREADS=$T SPMM_D=$D ARLIB_B200_SEGMENT=64 timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/threads=$T seg=64 /" >> $O/spmm_threads.txt
  done
done
cat $O/spmm_threads.txt
