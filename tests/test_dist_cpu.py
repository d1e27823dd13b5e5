"""World-size-2 gloo test (CPU) of the multi-GPU HOST logic: nnz-balanced row
ranges, per-rank segmented work plans, and the gather layout (every rank's rows
land at their global offsets).  The per-rank arithmetic here is the numpy oracle --
the CUDA kernels are covered by tests/test_gpu_dist.py on real GPUs."""
import os
import socket

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from arlib_b200.graph import DeviceGraph
    from arlib_b200.util.synth import synth_edges
    from oracle import port as oracle
    U, I, E, d = 200, 300, 5000, 8
    tu, ti, _, _ = synth_edges(U, I, E, seed=3)
    norm = oracle.normalize_graph_mat(oracle.bipartite_adjacency(tu, ti, U, I))
    g = DeviceGraph.from_scipy(norm, "cpu")
    bounds = g.row_ranges(world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    gp = g.partition(r0, r1)
    torch.manual_seed(0)
    X = torch.randn(U + I, d)
    # this rank's rows through its work plan (segments of rows; partial sums added per row)
    mine = torch.zeros(U + I, d)
    covered = torch.zeros(U + I, dtype=torch.int64)
    lens = gp.vrows[:, 1]
    assert bool((lens[:-1] >= lens[1:]).all()) and int(lens.max()) <= 64          # sorted, bounded work items
    for v in range(gp.n_vrows):
        a, n, r, seg = (int(x) for x in gp.vrows[v])
        assert r0 <= r < r1 and 0 <= (seg & 0xffff) < (seg >> 16)
        mine[r] += (gp.val[a:a + n, None] * X[gp.col[a:a + n].long()]).sum(0)
        covered[r] += n
    deg = (g.rowptr[1:] - g.rowptr[:-1]).long()
    assert torch.equal(covered[r0:r1], deg[r0:r1]) and int(covered.sum()) == int(deg[r0:r1].sum())
    # "all-gather": every rank contributes its row range at the global offsets
    parts = [torch.zeros(U + I, d) for _ in range(world)]
    dist.all_gather(parts, mine)
    full = sum(parts)
    ref = torch.sparse.mm(oracle.to_torch_coo(norm), X)
    ok = bool(torch.allclose(full, ref, atol=1e-5))
    nnz = [int(g.rowptr[bounds[p + 1]] - g.rowptr[bounds[p]]) for p in range(world)]
    balanced = max(nnz) <= 1.2 * (sum(nnz) / world) + 200
    if rank == 0:
        out.put((ok, balanced, bounds, nnz))
    dist.destroy_process_group()


def test_row_partition_and_gather_layout_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, balanced, bounds, nnz = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "gathered rows differ from the full product"
    assert balanced, (bounds, nnz)
    assert bounds[0] == 0 and bounds[-1] == 500


def test_balanced_row_ranges_edge_cases():
    from arlib_b200.graph import balanced_row_ranges
    indptr = np.array([0, 0, 0, 10, 10, 20, 1000, 1000])
    for w in (1, 2, 3, 8):
        b = balanced_row_ranges(indptr, w)
        assert len(b) == w + 1 and b[0] == 0 and b[-1] == 7 and all(x <= y for x, y in zip(b, b[1:]))


def _dshard_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import types
    from arlib_b200.engine import LightGCNEngine, dshard_columns
    N, d = 37, 32
    torch.manual_seed(0)
    table = torch.randn(N, d)
    c0, c1 = dshard_columns(d, world, rank)
    fake = types.SimpleNamespace(mode="dshard", N=N, d_full=d, comm=types.SimpleNamespace(world=world, group=dist.group.WORLD))
    full = LightGCNEngine.full_table(fake, table[:, c0:c1].contiguous())
    # the partial scores of the d-sharded loss: per-slice dots summed in rank order == the full dot
    u, i = torch.arange(10), torch.arange(10, 20)
    part = (table[u, c0:c1] * table[i, c0:c1]).sum(1)
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    ok_dot = torch.allclose(sum(parts), (table[u] * table[i]).sum(1), atol=1e-5)
    if rank == 0:
        out.put((bool(torch.equal(full, table)), bool(ok_dot), (c0, c1)))
    dist.destroy_process_group()


def test_dshard_column_layout_world2():
    """Column-sharded mode: slices [r*d/P, (r+1)*d/P), reassembly by all-gather + interleave, and
    additivity of the per-slice partial scores."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dshard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, ok_dot, cols = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and ok_dot and cols == (0, 16)


def test_dshard_columns_rejects_unsupported_widths():
    import pytest
    from arlib_b200.engine import dshard_columns
    assert dshard_columns(64, 8, 3) == (24, 32)
    with pytest.raises(ValueError):
        dshard_columns(64, 3, 0)
    with pytest.raises(ValueError):
        dshard_columns(64, 16, 0)          # 4-column slices are below the 16-byte lane granularity x 2
