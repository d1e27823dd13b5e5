# 1 GPU: profiler evidence + sanitizer + side benches (run after job_r2.sh is green)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b; mkdir -p $O
export PYTHONUNBUFFERED=1
# launch lists (ncu time-only pass; cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_train.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-epoch-e2e > $O/launches_train.json 2> $O/launches_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_eval.csv python tools/eval_bench.py > $O/launches_eval.log 2>&1
# full captures: the shipped scoring GEMM, the full-width SpMM, the narrow cooperative SpMM (d = 8)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_max_tc_kernel -c 2 -o $O/eval_gemm python tools/eval_bench.py > $O/eval_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2_ -c 8 -o $O/eval_stage2 python tools/eval_bench.py > $O/eval_stage2.log 2>&1
SPMM_D=8 ARLIB_B200_SEGMENT=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 20 -c 2 -o $O/spmm_d8 python tools/spmm_variants.py > $O/spmm_d8.log 2>&1
AGCF_SPMM_COOP=1 SPMM_D=8 ARLIB_B200_SEGMENT=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_coop_kernel -s 20 -c 2 -o $O/spmm_coop_d8 python tools/spmm_variants.py > $O/spmm_coop_d8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 20 -c 2 -o $O/spmm_d64 python tools/spmm_variants.py > $O/spmm_d64.log 2>&1
for f in eval_gemm eval_stage2 spmm_d8 spmm_coop_d8 spmm_d64; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
done
# compute-sanitizer: memcheck + racecheck over the kernel tests that combine segments / re-zero G / exchange
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $O/sanitizer_$tool.log \
     python -m pytest tests/test_gpu_propagate.py tests/test_gpu_fused_step.py tests/test_gpu_bpr.py -x -q -m gpu -k "not golden" > $O/sanitizer_$tool.out 2>&1
  echo "sanitizer $tool rc=$?" >> $O/sanitizer_$tool.out
  tail -3 $O/sanitizer_$tool.out; tail -3 $O/sanitizer_$tool.log
done
# side benches
timeout 600 python tools/contrast_bench.py > $O/contrast_yelp2018.jsonl 2> $O/contrast.err
timeout 600 python tools/ngcf_bench.py > $O/ngcf_gowalla.json 2> $O/ngcf.err
timeout 600 python tools/whitebox_bench.py > $O/whitebox_ml1m.jsonl 2> $O/whitebox.err
timeout 900 python bench.py --workload c5b-small --steps 10 --warmup 3 > $O/c5b_small_n1.json 2> $O/c5b_small_n1.err
timeout 900 python bench.py --workload amazon-book --steps 200 --warmup 5 --no-epoch-e2e > $O/amazon_n1.json 2> $O/amazon_n1.err
ls -la $O | head -50
