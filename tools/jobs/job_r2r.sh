# 1 GPU: (col, val) pairs broadcast through shared memory instead of shuffles (AGCF_SPMM_SMEM_BCAST build) vs shipped
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2r; mkdir -p $O
export PYTHONUNBUFFERED=1
for LIB in libagcf.so csrc/build/libagcf_sb.so; do
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/$(basename $LIB) /" >> $O/spmm_smem_bcast.txt
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 300 python tools/spmm_variants.py amazon-book 2>&1 | tail -1 | sed "s/^/$(basename $LIB) /" >> $O/spmm_smem_bcast.txt
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB SPMM_PART=8 timeout 300 python tools/spmm_variants.py 2>&1 | tail -8 | sed "s/^/$(basename $LIB) /" >> $O/spmm_smem_bcast.txt
done
cat $O/spmm_smem_bcast.txt
ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_sb.so timeout 600 python -m pytest tests/test_gpu_propagate.py tests/test_gpu_fused_step.py -x -q -m gpu > $O/tests_sb.log 2>&1; tail -2 $O/tests_sb.log
for LIB in libagcf.so csrc/build/libagcf_sb.so; do
ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_$(basename $LIB).json 2> $O/bench_$(basename $LIB).err; python -c "
import json;d=json.loads(open('$O/bench_$(basename $LIB).json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['batch_sparse_launch_ms'],d['eval']['ms'])"
done
