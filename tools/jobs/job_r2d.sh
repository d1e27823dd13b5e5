set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -12 $O/all_tests.log
timeout 300 python -m pytest tests/test_gpu_golden_models.py tests/test_gpu_ngcf.py -q -m gpu -s 2>&1 | grep -i "vs reference\|vs the\|vs oracle\|elements further" > $O/golden_errors.txt; cat $O/golden_errors.txt
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; cat $O/eval_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_eval.csv python tools/eval_bench.py > $O/launches_eval.log 2>&1
# narrow SpMM: segment length sweep (work-item granularity vs the latency chain)
for D in 8 16; do for SEG in 16 32 64 128; do
  SPMM_D=$D ARLIB_B200_SEGMENT=$SEG timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/seg=$SEG /" >> $O/spmm_narrow_segments.txt
done; done
cat $O/spmm_narrow_segments.txt
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --workload c5b-small --steps 10 --warmup 3 > $O/c5b_small_n1.json 2> $O/c5b_small_n1.err; echo "c5b-small rc=$?"; tail -5 $O/c5b_small_n1.err; head -c 1500 $O/c5b_small_n1.json
