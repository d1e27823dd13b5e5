"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle only finishes small cases in
seconds): linearity / symmetry / row sums / determinism / mask consistency of the propagation at the Gowalla (C2) and
Amazon-book (C5a) shapes, permutation + rejection properties of the device sampler over a whole epoch, and
cross-implementation / ordering / mask / optimality properties of the fused top-K over all test users."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _graph(name):
    from bench import make_data
    from arlib_b200.graph import DeviceGraph
    D = make_data(name, 0.5)
    U, I, E = D["U"], D["I"], D["E"]
    N = U + I
    half = sp.csr_matrix((np.ones(E, dtype=np.float32), (D["tu"], D["ti"] + U)), shape=(N, N), dtype=np.float32)
    return D, DeviceGraph.from_dataloader_adj(half + half.T, DEV)


@pytest.mark.parametrize("name", ["gowalla", "amazon-book"])
def test_propagation_properties_at_full_size(name):
    from arlib_b200 import ops
    D, g = _graph(name)
    N, d = g.n_rows, D["d"]
    assert g.nnz == 2 * D["E"]
    gen = torch.Generator(device=DEV).manual_seed(0)
    X = torch.randn(N, d, device=DEV, generator=gen)
    Z = torch.randn(N, d, device=DEV, generator=gen)
    AX, AZ, AL = torch.empty_like(X), torch.empty_like(X), torch.empty_like(X)
    ops.spmm(g, X, Y=AX)
    ops.spmm(g, Z, Y=AZ)
    # determinism: a second launch gives the same bits (no floating-point atomics anywhere)
    again = torch.empty_like(X)
    ops.spmm(g, X, Y=again)
    assert torch.equal(again, AX)
    # linearity
    ops.spmm(g, 0.75 * X - 1.5 * Z, Y=AL)
    lin = 0.75 * AX - 1.5 * AZ
    assert float((AL - lin).abs().max()) <= 2e-5 * float(lin.abs().max())
    # symmetry of the normalized adjacency: <A X, Z> == <X, A Z>
    lhs = float((AX.double() * Z.double()).sum())
    rhs = float((X.double() * AZ.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0)
    # A . 1 == row sums of the stored values (fp64 reference of a different code path)
    ones = torch.ones(N, d, device=DEV)
    A1 = torch.empty_like(ones)
    ops.spmm(g, ones, Y=A1)
    rows = torch.repeat_interleave(torch.arange(N, device=DEV), (g.rowptr[1:] - g.rowptr[:-1]).long())
    rs = torch.zeros(N, dtype=torch.float64, device=DEV).index_add_(0, rows, g.val.double())
    assert float((A1[:, 0].double() - rs).abs().max()) < 1e-5
    assert torch.equal(A1[:, 0], A1[:, d - 1])
    # batch-sparse launches are consistent with the full product: rows of a row-masked launch are the same bits,
    # a column-masked launch on a table that is zero outside the mask equals the full product
    rng = np.random.default_rng(0)
    nodes = np.unique(rng.integers(0, N, 6000))
    words = np.zeros((N + 31) // 32, dtype=np.uint32)
    np.bitwise_or.at(words, nodes >> 5, (np.uint32(1) << (nodes & 31).astype(np.uint32)))
    mask = torch.from_numpy(words.view(np.int32)).to(DEV)
    idx = torch.from_numpy(nodes).to(DEV)
    out = torch.full_like(X, 3.0)
    ops.spmm(g, X, Y=out, row_mask=mask)
    assert torch.equal(out[idx], AX[idx])
    untouched = torch.ones(N, dtype=torch.bool, device=DEV)
    untouched[idx] = False
    assert bool((out[untouched] == 3.0).all())
    Xs = torch.zeros_like(X)
    Xs[idx] = X[idx]
    full, colm = torch.empty_like(X), torch.empty_like(X)
    ops.spmm(g, Xs, Y=full)
    ops.spmm(g, Xs, Y=colm, col_mask=mask)
    assert torch.equal(full, colm)


def test_device_sampler_epoch_properties_at_gowalla_size():
    from arlib_b200.engine import DeviceTrainSet, LightGCNEngine
    D, g = _graph("gowalla")
    U, I, E, d, B = D["U"], D["I"], D["E"], D["d"], D["B"]
    ts = DeviceTrainSet.from_arrays(D["tu"], D["ti"], U, I, DEV)
    table = torch.zeros(U + I, d, device=DEV)
    eng = LightGCNEngine(g, table, U, D["L"], 0.005, 1e-4, B, E)
    eng.sample_epoch(ts, 2018, 0)
    u0, i0, j0 = eng.tu[:E].clone(), eng.ti[:E].clone(), eng.tj[:E].clone()
    # every training edge exactly once per epoch
    key = torch.sort(u0.long() * I + i0.long()).values
    ref = torch.sort(torch.from_numpy(D["tu"].astype(np.int64) * I + D["ti"].astype(np.int64)).to(DEV)).values
    assert torch.equal(key, ref)
    # negatives are never training items of their user
    train_keys = ref
    neg = u0.long() * I + j0.long()
    pos = torch.searchsorted(train_keys, neg).clamp(max=E - 1)
    assert not bool((train_keys[pos] == neg).any())
    assert int(j0.min()) >= 0 and int(j0.max()) < I
    # uniform over items (chi-square on 64 buckets; rejection only removes ~0.1 % of the mass per user)
    hist = torch.bincount((j0.long() * 64) // I, minlength=64).double()
    chi2 = float(((hist - E / 64) ** 2 / (E / 64)).sum())
    assert chi2 < 140, chi2                                   # 63 dof: P(chi2 > 140) ~ 1e-7
    # same (seed, epoch) -> same triples; another epoch -> another permutation
    eng.sample_epoch(ts, 2018, 0)
    assert torch.equal(eng.tu[:E], u0) and torch.equal(eng.tj[:E], j0)
    eng.sample_epoch(ts, 2018, 1)
    assert not torch.equal(eng.tu[:E], u0)
    # batches are a segmentation of the epoch: per-batch grouping covers 3 * nb occurrences
    nb_last = E - (eng.n_batches - 1) * B
    assert int(eng.seg_off[(eng.n_batches - 1) * (3 * B + 1) + int(eng.n_seg[eng.n_batches - 1])]) == 3 * nb_last


def test_topk_properties_all_test_users_at_gowalla_size():
    from arlib_b200 import ops
    from arlib_b200.evaluator import FullRankEvaluator
    D, _ = _graph("gowalla")
    U, I, d, K = D["U"], D["I"], D["d"], 50
    gen = torch.Generator(device=DEV).manual_seed(1)
    ue = torch.randn(U, d, device=DEV, generator=gen) * 0.1
    ie = torch.randn(I, d, device=DEV, generator=gen) * 0.1
    ev = FullRankEvaluator.from_arrays(U, I, D["tu"], D["ti"], D["su"], D["si"], DEV)
    v1, i1 = ev.topk(ue, ie, K, impl=1)
    v0, i0 = ev.topk(ue, ie, K, impl=0)
    # tensor-core stage 1 and CUDA-core stage 1 give the same bits; a second run too
    assert torch.equal(i1, i0) and torch.equal(v1, v0)
    v1b, i1b = ev.topk(ue, ie, K, impl=1)
    assert torch.equal(i1, i1b) and torch.equal(v1, v1b)
    n = i1.shape[0]
    assert n == ev.user_rows.numel() and n > 20000
    # sorted by score descending, ids valid and distinct per user
    assert bool((v1[:, :-1] >= v1[:, 1:]).all())
    assert int(i1.min()) >= 0 and int(i1.max()) < I
    srt = torch.sort(i1.long(), dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    # a training item of the user never appears
    train_keys = torch.sort(torch.from_numpy(D["tu"].astype(np.int64) * I + D["ti"].astype(np.int64)).to(DEV)).values
    keys = (ev.user_rows.long()[:, None] * I + i1.long()).reshape(-1)
    pos = torch.searchsorted(train_keys, keys).clamp(max=train_keys.numel() - 1)
    assert not bool((train_keys[pos] == keys).any())
    # optimality on a sample: values are the exact fp32 scores and nothing outside the list beats its K-th value
    sample = torch.arange(0, n, n // 128, device=DEV)[:128]
    exact = ops.score_rows(ue, ev.user_rows[sample].contiguous(), ie)
    got = torch.gather(exact, 1, i1[sample].long())
    assert torch.equal(got, v1[sample])
    masked = exact.clone()
    mrp, mit = ev.mask_rowptr.cpu().numpy(), ev.mask_items.cpu().numpy()
    for r, u in enumerate(ev.user_rows[sample].cpu().tolist()):
        masked[r, torch.from_numpy(mit[mrp[u]:mrp[u + 1]].astype(np.int64)).to(DEV)] = -1e9
    masked.scatter_(1, i1[sample].long(), float("-inf"))
    assert bool((masked.max(dim=1).values <= v1[sample][:, -1]).all())


def test_training_steps_at_gowalla_size_graph_replay_equals_eager():
    from arlib_b200.engine import DeviceTrainSet, LightGCNEngine
    from bench import xavier_tables
    D, g = _graph("gowalla")
    U, I, E, d, B, L = D["U"], D["I"], D["E"], D["d"], D["B"], D["L"]
    ts = DeviceTrainSet.from_arrays(D["tu"], D["ti"], U, I, DEV)
    ue, ie = xavier_tables(U, I, d)
    runs = []
    for use_graph in (False, True):
        table = torch.cat([ue, ie]).to(DEV)
        eng = LightGCNEngine(g, table, U, L, 0.005, 1e-4, B, E)
        eng.sample_epoch(ts, 7, 0)
        losses = eng.run_steps(0, 150, use_graph=use_graph)[:, :2].clone()
        runs.append((table.clone(), losses))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    bpr = runs[0][1][:, 1].cpu().numpy()                      # the ranking term (the regulariser grows from a xavier init)
    assert np.isfinite(runs[0][1].cpu().numpy()).all() and bpr[-5:].mean() < bpr[:5].mean() - 1e-3 and bpr[0] < 0.70
