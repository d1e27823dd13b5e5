"""GPU parity of the fused training-step variants: per-batch work lists (last forward
layer), the sparse column-masked kernel (first backward layer), Adam + re-zeroing of G
fused into the last backward SpMM, persistent CTAs with dynamic block scheduling.  Each
must reproduce the plain launch sequence (recommender/LightGCN.py:50-64): the optimizer /
scheduling variants bit for bit, the work lists up to the association of a row's partial sums."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand_graph(U, I, E, seed, hub=0):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, E)
    i = rng.integers(0, I, E)
    if hub:
        u[:hub] = 0
        i[:hub] = rng.permutation(I)[:hub]
        i[hub:2 * hub] = 1
        u[hub:2 * hub] = rng.permutation(U)[:hub]
    key = np.unique(u.astype(np.int64) * I + i)
    return key // I, key % I


def _batch_nodes(U, I, B, seed):
    from arlib_b200 import ops
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, B).astype(np.int32)
    u[:3] = 0                                              # the user hub is in the batch
    i = rng.integers(0, I, B).astype(np.int32)
    i[:3] = 1                                              # and the item hub
    j = rng.integers(0, I, B).astype(np.int32)
    d = lambda a: torch.from_numpy(a).to(DEV)
    occ = torch.empty(3 * B, dtype=torch.int32, device=DEV)
    seg_off = torch.empty(3 * B + 1, dtype=torch.int32, device=DEV)
    seg_node = torch.empty(3 * B, dtype=torch.int32, device=DEV)
    n_seg = torch.zeros(1, dtype=torch.int32, device=DEV)
    words = (U + I + 31) // 32
    mask = torch.zeros(words, dtype=torch.int32, device=DEV)
    ops.bpr_group_batches(d(u), d(i), d(j), B, B, U, occ, seg_off, seg_node, n_seg, U + I, mask)
    return seg_node, n_seg, mask


@pytest.mark.parametrize("d", [16, 64, 128])
@pytest.mark.parametrize("segment", [16, 64])
def test_batch_worklist_spmm_equals_row_masked_spmm(d, segment):
    from arlib_b200 import ops
    from arlib_b200.graph import DeviceGraph
    U, I, B = 700, 900, 256
    u, i = _rand_graph(U, I, 15000, 9, 500)
    adj = port.bipartite_adjacency(u, i, U, I)
    g = DeviceGraph.from_dataloader_adj(adj, DEV)
    n = U + I
    seg_node, n_seg, mask = _batch_nodes(U, I, B, 1)
    ns = int(n_seg)
    nodes = seg_node[:ns].cpu().numpy()
    deg = np.diff(g.rowptr.cpu().numpy())[nodes]
    nseg = np.where(deg > segment, (deg + segment - 1) // segment, 1)
    cap = int(nseg.sum()) + 37
    wv = torch.full((1, cap, 4), -1, dtype=torch.int32, device=DEV)
    wp = torch.full((1, cap), -1, dtype=torch.int32, device=DEV)
    wc = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.spmm_batch_worklists(seg_node, n_seg, 1, 3 * B, g.rowptr, 0, n, segment, segment, wv, wp, wc)
    assert int(wc) == int(nseg.sum())
    items = wv[0, :int(wc)].cpu().numpy()
    assert np.array_equal(np.unique(items[:, 2]), nodes)                    # every batch row, nothing else
    assert int(items[:, 1].sum()) == int(deg.sum()) and items[:, 1].max() <= max(segment, deg[nseg == 1].max())
    X = torch.randn(n, d, device=DEV)
    acc = torch.randn(n, d, device=DEV)
    ref = torch.full((n, d), 7.0, device=DEV)
    ops.spmm(g, X, acc_in=acc, acc_out=ref, acc_div=3.0, row_mask=mask)
    out = torch.full((n, d), 7.0, device=DEV)
    part = torch.empty((cap, d), device=DEV)
    tick = torch.zeros(cap, dtype=torch.int32, device=DEV)
    for _ in range(2):                                                       # tickets are left at zero
        out.fill_(7.0)
        ops.spmm(g, X, acc_in=acc, acc_out=out, acc_div=3.0, worklist=(wv[0], wp[0], wc, part, tick))
        torch.testing.assert_close(out, ref, rtol=2e-6, atol=2e-6)
    live = torch.zeros(n, dtype=torch.bool)
    live[torch.from_numpy(nodes).long()] = True
    assert bool((out.cpu()[~live] == 7.0).all())
    assert int(tick.abs().sum()) == 0
    # a rank that owns rows [r0, r1) only lists its own rows
    r0, r1 = 300, 1100
    ops.spmm_batch_worklists(seg_node, n_seg, 1, 3 * B, g.rowptr, r0, r1, segment, segment, wv, wp, wc)
    rows = wv[0, :int(wc), 2].cpu().numpy()
    assert np.array_equal(np.unique(rows), nodes[(nodes >= r0) & (nodes < r1)])


@pytest.mark.parametrize("d", [8, 16, 32, 64, 128, 256])
def test_sparse_colmask_kernel_is_bit_identical_to_the_dense_product(d):
    """col_mask runs spmm_colmask_kernel (ballot over live entries): the live entries are
    accumulated in entry order, exactly like the dense loop over a zero-padded X."""
    from arlib_b200 import ops
    from arlib_b200.graph import DeviceGraph
    U, I = 700, 900
    u, i = _rand_graph(U, I, 15000, 9, 500)
    adj = port.bipartite_adjacency(u, i, U, I)
    g = DeviceGraph.from_dataloader_adj(adj, DEV)
    n = U + I
    rng = np.random.default_rng(0)
    for frac in (0.0, 0.13, 1.0):
        live = rng.random(n) < frac
        if frac > 0:
            live[:2] = True
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        for k in np.flatnonzero(live):
            words[k >> 5] |= np.uint32(1) << np.uint32(k & 31)
        mask = torch.from_numpy(words.view(np.int32)).to(DEV)
        X = torch.randn(n, d)
        X[~torch.from_numpy(live)] = 0
        X = X.to(DEV)
        add = torch.randn(n, d, device=DEV)
        dense = torch.empty(n, d, device=DEV)
        ops.spmm(g, X, Y=dense, addend=add)                               # dense kernel over the zero-padded X
        sparse = torch.empty(n, d, device=DEV)
        ops.spmm(g, X, Y=sparse, addend=add, col_mask=mask)
        assert torch.equal(dense, sparse)


@pytest.mark.parametrize("L", [1, 2])
def test_adam_fused_into_the_spmm_epilogue_is_bit_identical(L):
    from arlib_b200 import ops
    from arlib_b200.graph import DeviceGraph
    U, I, d = 300, 400, 64
    u, i = _rand_graph(U, I, 6000, 3, 200)
    g = DeviceGraph.from_dataloader_adj(port.bipartite_adjacency(u, i, U, I), DEV)
    n = U + I
    gen = torch.Generator().manual_seed(1)
    H = torch.randn(n, d, generator=gen).to(DEV)
    G = torch.zeros(n, d)
    rows = torch.randperm(n, generator=gen)[:97]
    G[rows] = torch.randn(97, d, generator=gen)
    G = G.to(DEV)
    p0 = torch.randn(n, d, generator=gen).to(DEV)
    m0 = (torch.randn(n, d, generator=gen) * 0.01).to(DEV)
    v0 = (torch.rand(n, d, generator=gen) * 1e-4).to(DEV)
    step = torch.full((1,), 6, dtype=torch.int32, device=DEV)
    # separate kernels
    dE = torch.empty(n, d, device=DEV)
    ops.spmm(g, H, acc_in=G, acc_out=dE, acc_div=float(L + 1))
    p1, m1, v1 = p0.clone(), m0.clone(), v0.clone()
    ops.adam_step(p1, dE, m1, v1, 0.005, step_dev=step)
    # fused
    coefs = torch.zeros(2, device=DEV)
    ops.adam_coefs(step, coefs, 0.005)
    assert int(step) == 6
    p2, m2, v2, G2 = p0.clone(), m0.clone(), v0.clone(), G.clone()
    ops.spmm(g, H, acc_in=G2, acc_div=float(L + 1), adam=(p2, m2, v2, coefs, 0.9, 0.999, 1e-8), zero_acc_in=True)
    assert torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2)
    assert int((G2 != 0).sum()) == 0
    ops.adam_coefs(step, coefs, 0.005, increment=True)
    assert int(step) == 7
    b1, b2, lr = float(np.float32(0.9)), float(np.float32(0.999)), float(np.float32(0.005))   # the ABI takes floats
    want = [lr / (1 - b1 ** 8), (1 - b2 ** 8) ** 0.5]
    np.testing.assert_allclose(coefs.cpu().numpy(), want, rtol=1e-6)


def _run_epoch(golden, monkeypatch, flags, use_graph, L=2):
    from arlib_b200.engine import LightGCNEngine
    from arlib_b200.graph import DeviceGraph
    for k, v in flags.items():
        monkeypatch.setenv(k, v)
    U, I = golden["user_names"].shape[0], golden["item_names"].shape[0]
    adj = port.bipartite_adjacency(golden["train_u"].astype(np.int64), golden["train_i"].astype(np.int64), U, I)
    g = DeviceGraph.from_dataloader_adj(adj, DEV)
    table = torch.cat([torch.from_numpy(golden["init_user_emb"]), torch.from_numpy(golden["init_item_emb"])]).to(DEV)
    T = golden["batch_u"].shape[0]
    eng = LightGCNEngine(g, table, U, L, 0.005, 1e-4, 2048, T)
    eng.set_triples(golden["batch_u"], golden["batch_i"], golden["batch_j"])
    losses = eng.run_steps(0, use_graph=use_graph).clone()
    torch.cuda.synchronize()
    assert int(eng.step_dev) == len(golden["batch_len"])
    assert int((eng.G != 0).sum()) == 0                       # G is back to all-zero after every step
    return table.clone(), losses, eng.m.clone(), eng.v.clone()


PLAIN = {"ARLIB_B200_WORKLISTS": "0", "ARLIB_B200_FUSE_ADAM": "0"}


@pytest.mark.parametrize("L", [1, 2, 3])
@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_step_variants_reproduce_the_plain_step(golden, monkeypatch, use_graph, L):
    import arlib_b200.graph as graphmod
    monkeypatch.setattr(graphmod, "PERSISTENT_ENV", "0")
    ref = _run_epoch(golden, monkeypatch, PLAIN, use_graph, L)
    for persistent in ("0", "1"):
        monkeypatch.setattr(graphmod, "PERSISTENT_ENV", persistent)
        for flags in ({}, {"ARLIB_B200_FUSE_ADAM": "1"}):
            got = _run_epoch(golden, monkeypatch, dict(PLAIN, **flags), use_graph, L)
            for a, b in zip(ref, got):
                assert torch.equal(a, b), (persistent, flags)
    got = _run_epoch(golden, monkeypatch, {"ARLIB_B200_WORKLISTS": "1", "ARLIB_B200_FUSE_ADAM": "1"}, use_graph, L)
    torch.testing.assert_close(got[1], ref[1], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(got[0], ref[0], rtol=1e-5, atol=1e-7)
