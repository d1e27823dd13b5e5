# 1 GPU: validation after the SpMM launch changes (4-warp CTAs; shared-memory index broadcast for d >= 128): all tests, bench lines
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2s; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > $O/all_tests.log 2>&1; echo "rc=$?" >> $O/all_tests.log; tail -3 $O/all_tests.log
timeout 600 python bench.py --steps 500 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['avg_launch_ms'],d['roofline']['batch_sparse_launch_ms'],d['eval']['ms'],d['epoch_e2e']['train_epoch_s'])"
timeout 900 python bench.py --workload amazon-book --steps 200 --warmup 5 --no-epoch-e2e --no-cpu-baseline > $O/bench_amazon_n1.json 2> $O/bench_amazon_n1.err
python -c "
import json;d=json.loads(open('$O/bench_amazon_n1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline'],d['eval']['ms'])"
timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 >> $O/spmm_final.txt
timeout 300 python tools/spmm_variants.py amazon-book 2>&1 | tail -1 >> $O/spmm_final.txt
for D in 8 16 32; do SPMM_D=$D ARLIB_B200_SEGMENT=64 timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 >> $O/spmm_final.txt; done
SPMM_D=8 ARLIB_B200_SEGMENT=32 timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 >> $O/spmm_final.txt
cat $O/spmm_final.txt
timeout 600 python tools/contrast_bench.py yelp2018 100 XSimGCL,SimGCL > $O/contrast_yelp2018.jsonl 2> $O/contrast.err; cut -c1-300 $O/contrast_yelp2018.jsonl
timeout 600 python tools/ngcf_bench.py > $O/ngcf_gowalla.jsonl 2> $O/ngcf.err; cut -c1-300 $O/ngcf_gowalla.jsonl
