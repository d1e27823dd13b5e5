"""Worker for tests/test_gpu_dist.py (run under torchrun, one rank per GPU): the
row-partitioned and the column-sharded multi-GPU engines must reproduce the single-GPU engine."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from arlib_b200.dist import DistContext
    from arlib_b200.engine import DeviceTrainSet, LightGCNEngine
    from arlib_b200.evaluator import FullRankEvaluator
    from arlib_b200.graph import DeviceGraph
    from arlib_b200.util.synth import synth_edges

    U, I, E, d, L, B = 3000, 4000, 90000, 64, 3, 2048
    tu, ti, su, si = synth_edges(U, I, E, seed=11)
    N = U + I
    half = sp.csr_matrix((np.ones(E, dtype=np.float32), (tu, ti + U)), shape=(N, N), dtype=np.float32)
    g = DeviceGraph.from_dataloader_adj(half + half.T, dev)
    torch.manual_seed(0)
    table0 = (torch.rand(N, d) - 0.5) * 0.2
    ts = DeviceTrainSet.from_arrays(tu, ti, U, I, dev)

    single = LightGCNEngine(g, table0.clone().to(dev), U, L, 0.005, 1e-4, B, E)
    single.sample_epoch(ts, 7, 0)
    single.run_steps(0, 5, use_graph=False)
    Fs = single.forward_table().clone()

    comm = DistContext(dev)
    multi = LightGCNEngine(g, table0.clone().to(dev), U, L, 0.005, 1e-4, B, E, comm=comm)
    multi.sample_epoch(ts, 7, 0)
    assert torch.equal(multi.tu[:E], single.tu[:E]) and torch.equal(multi.tj[:E], single.tj[:E])
    multi.run_steps(0, 5, use_graph=False)
    torch.cuda.synchronize()
    dist.barrier()
    # every rank holds the full, identical table
    err = float((multi.E0 - single.E0).abs().max())
    loss_err = float((multi.out4[:5] - single.out4[:5]).abs().max())
    assert err < 1e-6, "rank %d: table differs from single-GPU by %g" % (rank, err)
    assert loss_err < 1e-6, loss_err
    Fm = multi.forward_table()
    torch.cuda.synchronize(); dist.barrier()
    assert float((Fm - Fs).abs().max()) < 1e-6
    # column-sharded engine ("dshard"): graph replicated, every table split by columns, one exchange per step
    comm2 = DistContext(dev)
    shard = LightGCNEngine(g, table0.clone().to(dev), U, L, 0.005, 1e-4, B, E, comm=comm2, mode="dshard")
    assert shard.d == d // world
    shard.sample_epoch(ts, 7, 0)
    shard.run_steps(0, 3, use_graph=False)
    shard.run_steps(3, 2, use_graph=True)                  # graph capture incl. the stamped exchange and its parity halves
    torch.cuda.synchronize(); dist.barrier()
    full = shard.full_table(shard.E0)
    derr = float((full - single.E0).abs().max())
    dloss = float((shard.out4[:5] - single.out4[:5]).abs().max())
    assert derr < 2e-6, "rank %d: d-sharded table differs from single-GPU by %g" % (rank, derr)
    assert dloss < 2e-6, dloss
    Fd = shard.full_table(shard.forward_table())
    torch.cuda.synchronize(); dist.barrier()
    assert float((Fd - Fs).abs().max()) < 2e-6
    # every rank computed identical loss rows (partials are summed in rank order everywhere)
    rows = [torch.empty_like(shard.out4[:5]) for _ in range(world)]
    dist.all_gather(rows, shard.out4[:5].contiguous())
    assert all(torch.equal(rows[0], r) for r in rows)
    # item-sharded evaluation == unsharded
    ev = FullRankEvaluator.from_arrays(U, I, tu, ti, su, si, dev)
    v1, i1 = ev.topk(Fs[:U], Fs[U:], 50)
    v2, i2 = ev.topk_sharded(Fs[:U], Fs[U:], 50, rank, world)
    assert torch.equal(i1, i2) and torch.equal(v1, v2)
    # user-sharded evaluation (the default multi-GPU layout of the bench) == unsharded, on every rank
    v3, i3 = ev.topk_user_sharded(Fs[:U], Fs[U:], 50, rank, world)
    assert torch.equal(i1, i3) and torch.equal(v1, v3)
    # many d-sharded steps back to back through a captured graph: the stamped exchange must stay in step without
    # any barrier (both halves of the double buffer are re-used hundreds of times)
    shard.run_steps(0, 40, use_graph=True)
    for _ in range(4):
        shard.run_steps(0, 40, use_graph=True)
    torch.cuda.synchronize(); dist.barrier()
    rows = [torch.empty_like(shard.out4[:40]) for _ in range(world)]
    dist.all_gather(rows, shard.out4[:40].contiguous())
    assert all(torch.equal(rows[0], r) for r in rows) and bool(torch.isfinite(rows[0]).all())
    dist.barrier()
    if rank == 0:
        print("DIST_CHECK_OK world=%d table_err=%.2e dshard_err=%.2e" % (world, err, derr))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
