# 2 GPUs: last full pass on HEAD -- all GPU tests (incl. the 2-rank parity test), smoke, default bench, 2-GPU bench
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2last; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -3 $O/all_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]);print(d['steps'],d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['eval']['users_per_s'],d['epoch_e2e']['train_epoch_s'],d['epoch_e2e']['test_s'])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/bench_n2.json 2> $O/bench_n2.err; python -c "
import json;d=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1]);print('n2',d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'])"
