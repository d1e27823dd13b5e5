# 1 GPU: column-masked kernel with 4-warp CTAs (tuning build) vs shipped
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2x; mkdir -p $O
export PYTHONUNBUFFERED=1
for LIB in libagcf.so csrc/build/libagcf_cm128.so; do
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/$(basename $LIB) /" >> $O/spmm_cm128.txt
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 300 python tools/spmm_variants.py amazon-book 2>&1 | tail -1 | sed "s/^/$(basename $LIB) /" >> $O/spmm_cm128.txt
  for D in 16 32; do ARLIB_B200_LIB=$PWD/arlib_b200/$LIB SPMM_D=$D ARLIB_B200_SEGMENT=64 timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/$(basename $LIB) /" >> $O/spmm_cm128.txt; done
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_$(basename $LIB).json 2> $O/bench_$(basename $LIB).err; python -c "
import json;d=json.loads(open('$O/bench_$(basename $LIB).json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['batch_sparse_launch_ms'])"
done
cat $O/spmm_cm128.txt
