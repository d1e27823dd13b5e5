# 1 GPU: per-chunk evaluation workspaces that keep their mask bits -- all tests, eval timing at both shapes, bench
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ae; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -3 $O/all_tests.log
timeout 300 python tools/eval_bench.py 2>&1 | head -1 >> $O/eval.txt
timeout 300 python tools/eval_bench.py amazon-book 2>&1 | head -1 | sed "s/^/amazon /" >> $O/eval.txt
ARLIB_B200_EVAL_CHUNK=8192 timeout 300 python tools/eval_bench.py 2>&1 | head -1 | sed "s/^/chunk=8192 /" >> $O/eval.txt
cat $O/eval.txt
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['eval']['users_per_s'],d['eval']['roofline']['ms'],d['eval']['measure'][1],d['epoch_e2e']['train_epoch_s'],d['epoch_e2e']['test_s'])"
