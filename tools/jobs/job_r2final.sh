# 1 GPU: last validation of HEAD + refreshed side benches / profiler evidence
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2final; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -3 $O/all_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]);print(d['steps'],d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['eval']['users_per_s'],d['epoch_e2e']['train_epoch_s'],d['epoch_e2e']['test_s'],d['cpu_baseline']['value'],d['gpu_eager_baseline']['value'])"
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-160 $O/bench_ref.json
timeout 900 python bench.py --workload c5b-small --steps 10 --warmup 3 > $O/c5b_small_n1.json 2> $O/c5b_small_n1.err; python -c "
import json;d=json.loads(open('$O/c5b_small_n1.json').read().strip().splitlines()[-1]);print('c5b-small',d['value'],d['ms_per_step'],d['roofline'])" 
timeout 600 python tools/whitebox_bench.py > $O/whitebox_ml1m.jsonl 2> $O/whitebox.err; cut -c1-300 $O/whitebox_ml1m.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2_ -c 3 -o $O/eval_stage2 python tools/eval_bench.py > $O/eval_stage2.log 2>&1
ncu -i $O/eval_stage2.ncu-rep --page raw --csv > $O/eval_stage2_raw.csv 2>/dev/null; rm -f $O/*.ncu-rep
ls -la $O
