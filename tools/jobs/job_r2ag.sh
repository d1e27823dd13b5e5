# 1 GPU: one evaluation call for all users at the Amazon-book shape
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ag; mkdir -p $O
for CH in 32768 65536; do ARLIB_B200_EVAL_CHUNK=$CH timeout 300 python tools/eval_bench.py amazon-book 2>&1 | head -1 | sed "s/^/amazon chunk=$CH /" >> $O/eval.txt; done
for CH in 32768 65536; do ARLIB_B200_EVAL_CHUNK=$CH timeout 300 python tools/eval_bench.py yelp2018 2>&1 | head -1 | sed "s/^/yelp2018 chunk=$CH /" >> $O/eval.txt; done
cat $O/eval.txt
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
