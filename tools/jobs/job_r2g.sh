set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -5 $O/all_tests.log
for D in 8 16 32; do for SEG in 16 32 64; do
  SPMM_D=$D ARLIB_B200_SEGMENT=$SEG timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/seg=$SEG /" >> $O/spmm_narrow_segments.txt
done; done
timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 >> $O/spmm_narrow_segments.txt
cat $O/spmm_narrow_segments.txt
for D in 8; do
  ARLIB_B200_SEGMENT=32 ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_trace.so TRACE_D=$D timeout 300 python tools/spmm_trace.py > $O/spmm_cta_timeline_d${D}_seg32.txt 2>&1; head -8 $O/spmm_cta_timeline_d${D}_seg32.txt; tail -7 $O/spmm_cta_timeline_d${D}_seg32.txt
done
timeout 300 python tools/eval_bench.py > $O/eval_bench.txt 2>&1; head -2 $O/eval_bench.txt
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
timeout 600 python tools/contrast_bench.py yelp2018 100 XSimGCL,SimGCL > $O/contrast_yelp2018.jsonl 2> $O/contrast.err; cat $O/contrast_yelp2018.jsonl | cut -c1-250
