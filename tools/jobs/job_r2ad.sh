# 1 GPU: evaluator keeps the mask bits between evaluations -- tests + timing
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ad; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_fullsize.py tests/test_gpu_golden_models.py tests/test_gpu_train.py tests/test_gpu_dropin_ref.py -x -q -m gpu > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log; tail -3 $O/tests.log
for rep in 1 2; do timeout 300 python tools/eval_bench.py 2>&1 | head -1 >> $O/eval.txt; done
timeout 300 python tools/eval_bench.py amazon-book 2>&1 | head -1 | sed "s/^/amazon /" >> $O/eval.txt
cat $O/eval.txt
timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['eval']['ms'],d['eval']['ms_mean'],d['eval']['users_per_s'],d['eval']['measure'],d['epoch_e2e']['train_epoch_s'],d['epoch_e2e']['test_s'])"
