# 1 GPU: final validation of HEAD + profiler evidence of the final kernels
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2zz; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -3 $O/all_tests.log
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['eval']['ms'],d['eval']['users_per_s'],d['epoch_e2e']['train_epoch_s'],d['epoch_e2e']['test_s'])"
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default bench rc=$?"; cut -c1-200 $O/bench_default.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_train.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-epoch-e2e > $O/launches_train.json 2> $O/launches_train.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 20 -c 2 -o $O/spmm_d128 python tools/spmm_variants.py amazon-book > $O/spmm_d128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 20 -c 2 -o $O/spmm_d64 python tools/spmm_variants.py > $O/spmm_d64.log 2>&1
for f in spmm_d128 spmm_d64; do ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null; done
rm -f $O/*.ncu-rep
ls -la $O
