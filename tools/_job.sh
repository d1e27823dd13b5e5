python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s12.log
tail -4 gpurun_out/pytest_s12.log
python tools/eval_bench.py 2>&1 | tail -6
AGCF_STAGE2_CTAS_PER_SM=4 python tools/eval_bench.py 2>&1 | head -1
AGCF_STAGE2_CTAS_PER_SM=3 python tools/eval_bench.py 2>&1 | head -1
