# 1 GPU: plan parameters re-swept with the 4-warp CTAs (work-list segment, static segment)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2aa; mkdir -p $O
export PYTHONUNBUFFERED=1
for WL in 32 64 128; do
  ARLIB_B200_WL_SEGMENT=$WL timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/b.json 2> $O/b.err; python -c "
import json;d=json.loads(open('$O/b.json').read().strip().splitlines()[-1]);print('wl_segment=$WL',d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['batch_sparse_launch_ms'])" >> $O/plan_sweep.txt
done
for SEG in 128 512; do
  ARLIB_B200_SEGMENT=$SEG timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/b.json 2> $O/b.err; python -c "
import json;d=json.loads(open('$O/b.json').read().strip().splitlines()[-1]);print('segment=$SEG',d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['batch_sparse_launch_ms'])" >> $O/plan_sweep.txt
done
cat $O/plan_sweep.txt
timeout 300 python -m pytest tests/test_gpu_topk.py -x -q -m gpu -k "exact_scores" > $O/tests_topk.log 2>&1; tail -2 $O/tests_topk.log
