set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -25 $O/all_tests.log
timeout 300 python -m pytest tests/test_gpu_golden_models.py tests/test_gpu_ngcf.py -q -m gpu -s 2>&1 | grep -i "vs reference\|vs the\|engine vs" > $O/golden_errors.txt; cat $O/golden_errors.txt
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; cat $O/eval_bench.txt
AGCF_STAGE2_IMPL=1 timeout 300 python tools/eval_bench.py > $O/eval_bench_impl1.txt 2>&1; head -2 $O/eval_bench_impl1.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_eval.csv python tools/eval_bench.py > $O/launches_eval.log 2>&1
SPMM_D=8 ARLIB_B200_SEGMENT=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 20 -c 2 -o $O/spmm_d8 python tools/spmm_variants.py > $O/spmm_d8.log 2>&1
ncu -i $O/spmm_d8.ncu-rep --page raw --csv > $O/spmm_d8_raw.csv 2>/dev/null
timeout 600 python tools/ngcf_bench.py > $O/ngcf_gowalla.json 2> $O/ngcf.err; cat $O/ngcf_gowalla.json; tail -3 $O/ngcf.err
timeout 600 python bench.py --steps 200 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
