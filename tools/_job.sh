#!/bin/bash
for seg in 32 64 128 256; do echo "segment $seg"; ARLIB_B200_SEGMENT=$seg timeout 300 python tools/spmm_variants.py 2>&1 | tail -1; done
