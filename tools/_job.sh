python -m pytest tests/test_gpu_propagate.py tests/test_gpu_topk.py -x -q 2>&1 | tail -2
python tools/spmm_variants.py
ARLIB_B200_SEGMENT=128 python tools/spmm_variants.py
ARLIB_B200_SPLIT_ABOVE=512 python tools/spmm_variants.py
ARLIB_B200_SPLIT_ABOVE=192 python tools/spmm_variants.py
python tools/spmm_variants.py gowalla 0.8
python tools/spmm_variants.py amazon-book
python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s19_n1.json 2> gpurun_out/bench_s19_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s19_n1.json').read().strip().splitlines()[-1])
print('value %.3gM ms/step %.4f e2e %.3gM (%.4f ms) spmm %.1f us eval %.3gM users/s (%.2f ms)'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['e2e']['ms_per_step'],d['roofline']['avg_launch_ms']*1e3,d['eval']['users_per_s']/1e6,d['eval']['ms']))
PY
