# 1 GPU: state check of HEAD (all GPU tests, bench, reference arm), then profiler evidence + compute-sanitizer
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2h; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log
tail -5 $O/all_tests.log
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
head -c 2500 $O/bench_n1.json
# launch lists (ncu time-only pass; cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_train.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-epoch-e2e > $O/launches_train.json 2> $O/launches_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_eval.csv python tools/eval_bench.py > $O/launches_eval.log 2>&1
# full captures: the shipped scoring GEMM, stage 2, the full-width SpMM
timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_max_tc_kernel -c 2 -o $O/eval_gemm python tools/eval_bench.py > $O/eval_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2_ -c 6 -o $O/eval_stage2 python tools/eval_bench.py > $O/eval_stage2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 20 -c 2 -o $O/spmm_d64 python tools/spmm_variants.py > $O/spmm_d64.log 2>&1
for f in eval_gemm eval_stage2 spmm_d64; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
done
# compute-sanitizer: memcheck + racecheck over the kernel tests that combine segments / re-zero G / exchange
for tool in memcheck racecheck; do
  timeout 540 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $O/sanitizer_$tool.log \
     python -m pytest tests/test_gpu_propagate.py tests/test_gpu_fused_step.py tests/test_gpu_bpr.py -q -m gpu -k "not golden" > $O/sanitizer_$tool.out 2>&1
  echo "sanitizer $tool rc=$?" >> $O/sanitizer_$tool.out
  tail -3 $O/sanitizer_$tool.out; tail -3 $O/sanitizer_$tool.log
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
rm -f $O/*.ncu-rep.tmp
ls -la $O | head -50
