# 1 GPU: adaptive user chunk -- evaluation tests + timings
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ah; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_fullsize.py tests/test_gpu_golden_models.py tests/test_gpu_train.py tests/test_gpu_dropin_ref.py -x -q -m gpu > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 300 python tools/eval_bench.py 2>&1 | head -1 >> $O/eval.txt
timeout 300 python tools/eval_bench.py amazon-book 2>&1 | head -1 | sed "s/^/amazon /" >> $O/eval.txt
cat $O/eval.txt
timeout 600 python bench.py --workload amazon-book --steps 100 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_amazon.json 2> $O/bench_amazon.err; python -c "
import json;d=json.loads(open('$O/bench_amazon.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['eval']['ms'],d['eval']['roofline']['ms'])"
