# 1 GPU: d = 8 CTA timeline, eval chunk size, then the profiling / sanitizer pass (job_r2b.sh)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2f; mkdir -p $O
for D in 8 16; do
  ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_trace.so TRACE_D=$D timeout 300 python tools/spmm_trace.py > $O/spmm_cta_timeline_d$D.txt 2>&1; head -12 $O/spmm_cta_timeline_d$D.txt; tail -7 $O/spmm_cta_timeline_d$D.txt
done
ARLIB_B200_EVAL_CHUNK=32768 timeout 300 python tools/eval_bench.py > $O/eval_bench_chunk32768.txt 2>&1; head -1 $O/eval_bench_chunk32768.txt
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; head -1 $O/eval_bench.txt
bash tools/jobs/job_r2b.sh
