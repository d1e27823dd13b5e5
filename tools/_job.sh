#!/bin/bash
mkdir -p gpurun_out
echo pb4; timeout 200 python tools/eval_bench.py 2>&1 | sed -n '1p'
echo pb8; ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_pb8.so timeout 200 python tools/eval_bench.py 2>&1 | sed -n '1p'
ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_pb8.so timeout 300 python -m pytest tests/test_gpu_topk.py -x -q 2>&1 | tail -1
