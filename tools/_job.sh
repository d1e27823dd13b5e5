for mode in rows dshard; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29610 + RANDOM % 300)) tools/dist_breakdown.py $mode 2>&1 | grep -v "^\*\|OMP\|^$\|NCCL version"
done
