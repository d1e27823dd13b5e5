# 1 GPU: software-pipelined re-scoring kernel -- top-K tests (bit-identity across stage-2 variants, goldens), A/B
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ac; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_fullsize.py tests/test_gpu_golden_models.py -x -q -m gpu > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log; tail -3 $O/tests.log
for rep in 1 2; do
  AGCF_S2RI_PIPE=0 timeout 300 python tools/eval_bench.py 2>&1 | head -1 | sed "s/^/pipe=0 /" >> $O/eval_pipe.txt
  AGCF_S2RI_PIPE=1 timeout 300 python tools/eval_bench.py 2>&1 | head -1 | sed "s/^/pipe=1 slices=2 /" >> $O/eval_pipe.txt
  for S in 1 4; do ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_sl$S.so timeout 300 python tools/eval_bench.py 2>&1 | head -1 | sed "s/^/pipe=1 slices=$S /" >> $O/eval_pipe.txt; done
done
cat $O/eval_pipe.txt
timeout 300 python tools/eval_bench.py amazon-book 2>&1 | head -1 | sed "s/^/amazon pipe=1 /" >> $O/eval_pipe.txt
AGCF_S2RI_PIPE=0 timeout 300 python tools/eval_bench.py amazon-book 2>&1 | head -1 | sed "s/^/amazon pipe=0 /" >> $O/eval_pipe.txt
tail -2 $O/eval_pipe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_eval.csv python tools/eval_bench.py > $O/launches_eval.log 2>&1
