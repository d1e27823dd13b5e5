# 1 GPU: Adam-coefficient kernel on a side branch (ARLIB_B200_COEF_STREAM) -- tests + A/B
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ab; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_fused_step.py tests/test_gpu_train.py tests/test_gpu_fullsize.py tests/test_gpu_golden_models.py tests/test_gpu_dropin_ref.py -x -q -m gpu > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log; tail -3 $O/tests.log
for rep in 1 2; do for CS in 0 1; do
  ARLIB_B200_COEF_STREAM=$CS timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-epoch-e2e > $O/b_$CS.json 2> $O/b_$CS.err; python -c "
import json;d=json.loads(open('$O/b_$CS.json').read().strip().splitlines()[-1]);print('coef_stream=$CS',d['value'],d['ms_per_step'],d['e2e']['value'],d['last_loss'],d['e2e']['last_loss'])" >> $O/coef_stream.txt
done; done
cat $O/coef_stream.txt
