set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python tools/ngcf_diag.py > $O/ngcf_diag.txt 2>&1; cat $O/ngcf_diag.txt | grep -v Warn | tail -12
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_ngcf.py::test_ngcf_class_takes_the_fused_path_and_keeps_weight_views > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -5 $O/all_tests.log
for D in 8 16; do for SEG in 32 64 128; do
  SPMM_D=$D ARLIB_B200_SEGMENT=$SEG timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/seg=$SEG /" >> $O/spmm_narrow_segments.txt
done; done
timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 >> $O/spmm_narrow_segments.txt
cat $O/spmm_narrow_segments.txt
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; head -2 $O/eval_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_eval.csv python tools/eval_bench.py > $O/launches_eval.log 2>&1
timeout 600 python tools/ngcf_bench.py > $O/ngcf_gowalla.json 2> $O/ngcf.err; cat $O/ngcf_gowalla.json
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
