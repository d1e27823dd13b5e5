set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/gpus.txt; nproc >> gpurun_out/r2/gpus.txt
timeout 900 python -m pytest tests/test_gpu_golden_models.py tests/test_gpu_dropin_ref.py tests/test_gpu_bpr.py tests/test_gpu_propagate.py -x -q -m gpu -s > gpurun_out/r2/t1_new_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t1_new_tests.log
tail -30 gpurun_out/r2/t1_new_tests.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2/t1_all_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t1_all_tests.log
tail -5 gpurun_out/r2/t1_all_tests.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r2/b1_n1.json 2> gpurun_out/r2/b1_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r2/b1_ref.json 2> gpurun_out/r2/b1_ref.err; echo "ref rc=$?"
cat gpurun_out/r2/b1_n1.json | head -c 3000
# narrow-slice SpMM: lane-group kernel vs warp-cooperative kernel, one GPU
for D in 8 16 32; do
  for COOP in 0 1; do
    AGCF_SPMM_COOP=$COOP SPMM_D=$D ARLIB_B200_SEGMENT=64 timeout 300 python tools/spmm_variants.py 2>&1 | tail -1 | sed "s/^/coop=$COOP /" >> gpurun_out/r2/spmm_narrow.txt
  done
done
cat gpurun_out/r2/spmm_narrow.txt
