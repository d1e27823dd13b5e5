# 1 GPU: split (cols / vals) shared-memory broadcast layout, d = 64 and d = 128, vs shipped
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2v; mkdir -p $O
export PYTHONUNBUFFERED=1
ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_sp1.so timeout 600 python -m pytest tests/test_gpu_propagate.py tests/test_gpu_fused_step.py -x -q -m gpu > $O/tests_sp1.log 2>&1; tail -2 $O/tests_sp1.log
for LIB in libagcf.so csrc/build/libagcf_sp1.so; do
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/$(basename $LIB) /" >> $O/spmm_smem_split.txt
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 300 python tools/spmm_variants.py amazon-book 2>&1 | tail -1 | sed "s/^/$(basename $LIB) /" >> $O/spmm_smem_split.txt
done
cat $O/spmm_smem_split.txt
for LIB in libagcf.so csrc/build/libagcf_sp1.so; do
ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_$(basename $LIB).json 2> $O/bench_$(basename $LIB).err; python -c "
import json;d=json.loads(open('$O/bench_$(basename $LIB).json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['batch_sparse_launch_ms'])"
ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 900 python bench.py --workload amazon-book --steps 200 --warmup 5 --no-epoch-e2e --no-cpu-baseline > $O/bench_amazon_$(basename $LIB).json 2> $O/bench_amazon_$(basename $LIB).err
python -c "
import json;d=json.loads(open('$O/bench_amazon_$(basename $LIB).json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'])"
done
