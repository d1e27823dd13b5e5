# 1 GPU: evaluation -- candidates row staged in smem on/off x user chunk x re-score CTAs/SM, 20 reps; top-K tests
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2o; mkdir -p $O; rm -f $O/eval_tuning.txt
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv > $O/smi.txt
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/tests.log 2>&1; tail -2 $O/tests.log
for rep in 1 2; do for LIB in libagcf.so csrc/build/libagcf_ri3.so; do for ST in 0 1; do for CH in 16384 32768; do
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB AGCF_S2_CAND_STAGE=$ST ARLIB_B200_EVAL_CHUNK=$CH timeout 300 python tools/eval_bench.py 2>&1 | head -1 | sed "s/^/$(basename $LIB) stage=$ST chunk=$CH /" >> $O/eval_tuning.txt
done; done; done; done
cat $O/eval_tuning.txt
ARLIB_B200_EVAL_CHUNK=32768 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_eval.csv python tools/eval_bench.py > $O/launches_eval.log 2>&1
