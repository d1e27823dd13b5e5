# 1 GPU: interleaved single-wave SpMM launches: parity tests, narrow-slice timings A/B, timeline, step bench
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2j; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_propagate.py tests/test_gpu_fused_step.py tests/test_gpu_train.py -x -q -m gpu > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log; tail -3 $O/tests.log
for IL in 0 1; do for D in 8 16 32; do for SEG in 32 64; do
  AGCF_SPMM_INTERLEAVE=$IL SPMM_D=$D ARLIB_B200_SEGMENT=$SEG timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/il=$IL seg=$SEG /" >> $O/spmm_narrow_interleave.txt
done; done; done
for IL in 0 2; do
AGCF_SPMM_INTERLEAVE=$IL timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/il=$IL /" >> $O/spmm_narrow_interleave.txt
done
for W in 4 8; do for IL in 0 1; do
AGCF_SPMM_INTERLEAVE=$IL SPMM_PART=$W timeout 300 python tools/spmm_variants.py 2>&1 | tail -$W | sed "s/^/il=$IL /" >> $O/spmm_partition_interleave.txt
done; done
cat $O/spmm_narrow_interleave.txt $O/spmm_partition_interleave.txt
for SEG in 32 64; do
  ARLIB_B200_SEGMENT=$SEG ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_trace.so TRACE_D=8 timeout 300 python tools/spmm_trace.py > $O/spmm_cta_timeline_d8_seg${SEG}_il.txt 2>&1; head -3 $O/spmm_cta_timeline_d8_seg${SEG}_il.txt; tail -7 $O/spmm_cta_timeline_d8_seg${SEG}_il.txt
done
for IL in 0 1; do
AGCF_SPMM_INTERLEAVE=$IL timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_n1_il$IL.json 2> $O/bench_n1_il$IL.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/bench_n1_il$IL.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['batch_sparse_launch_ms'],d['eval']['ms'],d['eval']['users_per_s'])"
done
