# 1 GPU: TMA gather4 micro-benchmark, narrow-slice SpMM at HEAD, d = 8 CTA timeline, evaluation after the host-sync fix
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2i; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 120 tools/tma_gather_peak > $O/tma_gather_peak.txt 2>&1; cat $O/tma_gather_peak.txt
timeout 120 tools/l2_gather_peak > $O/l2_gather_peak.txt 2>&1; tail -8 $O/l2_gather_peak.txt
for D in 8 16 32; do for SEG in 32 64; do
  SPMM_D=$D ARLIB_B200_SEGMENT=$SEG timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/seg=$SEG /" >> $O/spmm_narrow_segments.txt
done; done
timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 >> $O/spmm_narrow_segments.txt
cat $O/spmm_narrow_segments.txt
for SEG in 32 64; do
  ARLIB_B200_SEGMENT=$SEG ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_trace.so TRACE_D=8 timeout 300 python tools/spmm_trace.py > $O/spmm_cta_timeline_d8_seg$SEG.txt 2>&1; head -8 $O/spmm_cta_timeline_d8_seg$SEG.txt; tail -7 $O/spmm_cta_timeline_d8_seg$SEG.txt
done
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; head -3 $O/eval_bench.txt
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['eval']['ms'],d['eval']['users_per_s'])"
