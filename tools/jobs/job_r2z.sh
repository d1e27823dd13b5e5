# 1 GPU: Adam moments with evict-first loads / stores (tuning build) vs shipped
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2z; mkdir -p $O
export PYTHONUNBUFFERED=1
for rep in 1 2; do for LIB in libagcf.so csrc/build/libagcf_as.so; do
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_$(basename $LIB)_$rep.json 2> $O/bench_$(basename $LIB)_$rep.err; python -c "
import json;d=json.loads(open('$O/bench_$(basename $LIB)_$rep.json').read().strip().splitlines()[-1]);print('$LIB',d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['batch_sparse_launch_ms'])"
done; done
ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_as.so timeout 600 python -m pytest tests/test_gpu_fused_step.py -x -q -m gpu > $O/tests_as.log 2>&1; tail -2 $O/tests_as.log
