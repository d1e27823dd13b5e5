# 1 GPU: evaluation tuning (re-score at 4 CTAs/SM, user chunk size), then all GPU tests + bench of HEAD
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2n; mkdir -p $O
export PYTHONUNBUFFERED=1
for LIB in libagcf.so csrc/build/libagcf_ri4.so; do for CH in 8192 16384 32768; do
  ARLIB_B200_LIB=$PWD/arlib_b200/$LIB ARLIB_B200_EVAL_CHUNK=$CH timeout 300 python tools/eval_bench.py 2>&1 | head -1 | sed "s/^/$(basename $LIB) chunk=$CH /" >> $O/eval_tuning.txt
done; done
cat $O/eval_tuning.txt
timeout 1200 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -3 $O/all_tests.log
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['eval']['users_per_s'],d['eval']['e2e_users_per_s'],d['epoch_e2e'])"
