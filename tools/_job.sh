mkdir -p gpurun_out
python -m pytest tests/test_gpu_propagate.py -x -q > gpurun_out/pytest_s7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s7.log
grep -n "AgcfError\|passed\|failed" gpurun_out/pytest_s7.log | tail -5
python tools/spmm_variants.py
ARLIB_B200_ELL_HEAD=0 python tools/spmm_variants.py
for v in m5 m3; do echo $v; ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_$v.so python tools/spmm_variants.py; done
