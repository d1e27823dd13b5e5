#!/bin/bash
echo default; timeout 300 python tools/spmm_variants.py 2>&1 | tail -1
for v in d64l8m4 d64l8m5 d64l8m3; do echo $v; ARLIB_B200_LIB=$PWD/arlib_b200/csrc/build/libagcf_$v.so timeout 300 python tools/spmm_variants.py 2>&1 | tail -1; done
