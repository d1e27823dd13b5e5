# 1 GPU: warp-per-user metric kernel: all tests + bench
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2w; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -3 $O/all_tests.log
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['eval']['ms_mean'],d['eval']['users_per_s'],d['eval']['e2e_users_per_s'],d['epoch_e2e']['train_epoch_s'],d['epoch_e2e']['test_s'])"
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-200 $O/bench_ref.json
