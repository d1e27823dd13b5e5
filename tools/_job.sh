python tools/spmm_variants.py > gpurun_out/plain_stream3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 30 -c 1 -o gpurun_out/spmm_stream3 python tools/spmm_variants.py > gpurun_out/ncu_stream3.log 2>&1
