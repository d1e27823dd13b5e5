# 1 GPU: SpMM diagnostic -- all gathers redirected to a few hot rows (what does the kernel cost without L2 gather traffic?)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2p; mkdir -p $O
export PYTHONUNBUFFERED=1
for H in 1 64 4096; do for D in 64 32 16 8; do
  SPMM_HOTCOL=$H SPMM_D=$D timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/hotcol=$H /" >> $O/spmm_hotcol.txt
done; done
cat $O/spmm_hotcol.txt
