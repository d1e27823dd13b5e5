"""GPU parity: adjacency normalization (bit-exact), SpMM + fused epilogues, SDDMM,
encoders forward/backward -- against oracle/port.py on the same inputs."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _rand_graph(U, I, E, seed, hub=0):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, E)
    i = rng.integers(0, I, E)
    if hub:                                   # one very long row (> long-row threshold) per side
        u[:hub] = 0
        i[:hub] = rng.permutation(I)[:hub] if hub <= I else rng.integers(0, I, hub)
        i[hub:2 * hub] = 1
        u[hub:2 * hub] = rng.permutation(U)[:hub] if hub <= U else rng.integers(0, U, hub)
    key = np.unique(u.astype(np.int64) * I + i)
    return key // I, key % I


def test_norm_adj_bit_exact_golden(golden):
    from arlib_b200.graph import DeviceGraph
    U, I = golden["user_names"].shape[0], golden["item_names"].shape[0]
    adj = port.bipartite_adjacency(golden["train_u"].astype(np.int64), golden["train_i"].astype(np.int64), U, I)
    g = DeviceGraph.from_dataloader_adj(adj, _dev())
    assert np.array_equal(g.rowptr.cpu().numpy(), golden["adj_indptr"])
    assert np.array_equal(g.col.cpu().numpy(), golden["adj_indices"])
    assert np.array_equal(g.val.cpu().numpy().view(np.uint32), golden["adj_data"].view(np.uint32))
    g2 = DeviceGraph.from_ui_adj(adj, _dev())
    coo = g2.to_coo_tensor()
    assert np.array_equal(coo.indices()[0].cpu().numpy(), golden["uiadj_row"])
    assert np.array_equal(coo.indices()[1].cpu().numpy(), golden["uiadj_col"])
    assert np.array_equal(g2.val.cpu().numpy().view(np.uint32), golden["uiadj_data"].view(np.uint32))


def test_norm_adj_fractional_weights_bit_exact():
    """PGA writes rand() / 1e-7 weights (attack/White/PGA.py:73,139)."""
    from arlib_b200.graph import DeviceGraph
    rng = np.random.default_rng(5)
    U, I = 50, 70
    u, i = _rand_graph(U, I, 600, 1)
    w = rng.random(u.shape[0]).astype(np.float32)
    w[::7] = 1e-7
    n = U + I
    half = sp.csr_matrix((w, (u, i + U)), shape=(n, n), dtype=np.float32)
    adj = half + half.T
    ref = port.to_torch_coo(port.init_uiadj_norm(adj)).coalesce()
    g = DeviceGraph.from_ui_adj(adj, _dev())
    assert np.array_equal(g.val.cpu().numpy().view(np.uint32), ref.values().numpy().view(np.uint32))


@pytest.mark.parametrize("d", [8, 16, 32, 64, 128, 256])
@pytest.mark.parametrize("hub", [0, 700])
def test_spmm_matches_oracle(d, hub):
    from arlib_b200 import ops
    from arlib_b200.graph import DeviceGraph
    U, I = 900, 1100
    u, i = _rand_graph(U, I, 20000, 2, hub)
    adj = port.bipartite_adjacency(u, i, U, I)
    norm = port.normalize_graph_mat(adj)
    g = DeviceGraph.from_dataloader_adj(adj, _dev())
    if hub:
        assert g.max_segments >= 10 and g.n_partial > 20       # hub rows are cut into many segments
    X = torch.randn(U + I, d)
    ref = torch.sparse.mm(port.to_torch_coo(norm), X)
    Xd = X.to(_dev())
    Y = torch.empty_like(Xd)
    ops.spmm(g, Xd, Y=Y)
    torch.testing.assert_close(Y.cpu(), ref, rtol=1e-5, atol=1e-6)
    # fused epilogue: addend, running sum, mean division
    add = torch.randn(U + I, d)
    acc = torch.randn(U + I, d)
    accd = acc.to(_dev())
    Y2 = torch.empty_like(Xd)
    ops.spmm(g, Xd, Y=Y2, addend=add.to(_dev()), acc_in=accd, acc_out=accd, acc_div=3.0)
    torch.testing.assert_close(Y2.cpu(), ref + add, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(accd.cpu(), (acc + ref + add) / 3.0, rtol=1e-5, atol=1e-6)
    # acc only (no Y), identity row order == degree order result
    out = torch.empty_like(Xd)
    ops.spmm(g, Xd, acc_out=out)
    assert torch.equal(out, Y)


def test_spmm_noise_epilogue_matches_simgcl_lines():
    from arlib_b200 import ops
    from arlib_b200.graph import DeviceGraph
    U, I, d = 300, 500, 64
    u, i = _rand_graph(U, I, 6000, 3, 400)
    adj = port.bipartite_adjacency(u, i, U, I)
    g = DeviceGraph.from_dataloader_adj(adj, _dev())
    X = torch.randn(U + I, d)
    noise = torch.rand(U + I, d)
    ego = torch.sparse.mm(port.to_torch_coo(port.normalize_graph_mat(adj)), X)
    ref = ego + torch.sign(ego) * torch.nn.functional.normalize(noise, dim=-1) * 0.1
    Y = torch.empty(U + I, d, device=_dev())
    ops.spmm(g, X.to(_dev()), Y=Y, noise=noise.to(_dev()), eps=0.1)
    torch.testing.assert_close(Y.cpu(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("d", [32, 64, 128])
def test_sddmm_matches_pattern_masked_products(d):
    from arlib_b200 import ops
    from arlib_b200.graph import DeviceGraph
    U, I = 200, 300
    u, i = _rand_graph(U, I, 4000, 4, 280)
    adj = port.bipartite_adjacency(u, i, U, I)
    g = DeviceGraph.from_dataloader_adj(adj, _dev())
    H, E = torch.randn(U + I, d), torch.randn(U + I, d)
    coo = port.normalize_graph_mat(adj).tocoo()
    csr = port.normalize_graph_mat(adj).tocsr(); csr.sort_indices(); coo = csr.tocoo()
    ref = (H[coo.row] * E[coo.col]).sum(1)
    gval = torch.zeros(g.nnz, device=_dev())
    ops.sddmm(g, H.to(_dev()), E.to(_dev()), gval)
    torch.testing.assert_close(gval.cpu(), ref, rtol=1e-4, atol=1e-5)
    ops.sddmm(g, H.to(_dev()), E.to(_dev()), gval, accumulate=True)
    torch.testing.assert_close(gval.cpu(), 2 * ref, rtol=1e-4, atol=1e-5)


class _Data:
    def __init__(self, U, I, u, i):
        self.user_num, self.item_num = U, I
        self.ui_adj = port.bipartite_adjacency(u, i, U, I)
        self.norm_adj = port.normalize_graph_mat(self.ui_adj)


def _loss(fu, fi):
    w_u = torch.linspace(-1, 1, fu.shape[1], device=fu.device)
    return (fu * w_u).sum() * 0.01 + (fi ** 2).sum() * 0.5 + (fu[:7] @ fi[:9].T).sum()


@pytest.mark.parametrize("L", [1, 2, 3])
def test_lightgcn_encoder_forward_backward_and_adjacency_grad(L):
    from arlib_b200.encoder import LGCN_Encoder
    U, I, d = 120, 150, 64
    u, i = _rand_graph(U, I, 2500, 6)
    data = _Data(U, I, u, i)
    torch.manual_seed(0)
    enc = LGCN_Encoder(data, d, L)
    ue = enc.embedding_dict['user_emb'].detach().cpu().clone().requires_grad_(True)
    ie = enc.embedding_dict['item_emb'].detach().cpu().clone().requires_grad_(True)
    adj = port.to_torch_coo(data.norm_adj).coalesce().requires_grad_(True)
    ru, ri = port.lightgcn_forward(adj, ue, ie, L)
    _loss(ru, ri).backward()
    enc.sparse_norm_adj.requires_grad = True
    fu, fi = enc()
    torch.testing.assert_close(fu.detach().cpu(), ru.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(fi.detach().cpu(), ri.detach(), rtol=1e-5, atol=1e-6)
    _loss(fu, fi).backward()
    torch.testing.assert_close(enc.embedding_dict['user_emb'].grad.cpu(), ue.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(enc.embedding_dict['item_emb'].grad.cpu(), ie.grad, rtol=1e-4, atol=1e-5)
    ga = enc.sparse_norm_adj.grad.coalesce()
    gr = adj.grad.coalesce()
    assert torch.equal(ga.indices().cpu(), gr.indices())
    torch.testing.assert_close(ga.values().cpu(), gr.values(), rtol=1e-4, atol=1e-5)


def test_simgcl_xsimgcl_ngcf_encoders_match_oracle():
    from arlib_b200.encoder import NGCF_Encoder, SimGCL_Encoder, XSimGCL_Encoder
    U, I, d = 100, 140, 64
    u, i = _rand_graph(U, I, 2000, 7)
    data = _Data(U, I, u, i)
    adj = port.to_torch_coo(data.norm_adj)
    noises = [torch.rand(U + I, d) for _ in range(2)]
    src = lambda k, like: noises[k].to(like.device)
    for cls, fwd in ((SimGCL_Encoder, "sim"), (XSimGCL_Encoder, "xsim")):
        torch.manual_seed(1)
        enc = cls(data, d, 0.1, 2) if fwd == "sim" else cls(data, d, 0.1, 2, 1)
        enc.noise_source = src
        ue = enc.embedding_dict['user_emb'].detach().cpu().clone().requires_grad_(True)
        ie = enc.embedding_dict['item_emb'].detach().cpu().clone().requires_grad_(True)
        if fwd == "sim":
            ref = port.simgcl_forward(adj, ue, ie, 2, 0.1, noises)
        else:
            ref = port.xsimgcl_forward(adj, ue, ie, 2, 0.1, 1, noises)
        got = enc(True)
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            torch.testing.assert_close(a.detach().cpu(), b.detach(), rtol=1e-5, atol=1e-6)
        lr = sum(_loss(ref[k], ref[k + 1]) for k in range(0, len(ref), 2))
        lg = sum(_loss(got[k], got[k + 1]) for k in range(0, len(got), 2))
        lr.backward(); lg.backward()
        torch.testing.assert_close(enc.embedding_dict['user_emb'].grad.cpu(), ue.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(enc.embedding_dict['item_emb'].grad.cpu(), ie.grad, rtol=1e-4, atol=1e-5)
        # unperturbed pass
        clean = enc(False)
        rclean = port.simgcl_forward(adj, ue.detach(), ie.detach(), 2, 0.1, None)
        torch.testing.assert_close(clean[0].detach().cpu(), rclean[0], rtol=1e-5, atol=1e-6)
    torch.manual_seed(2)
    enc = NGCF_Encoder(data, d, 2)
    ue = enc.embedding_dict['user_emb'].detach().cpu().clone().requires_grad_(True)
    ie = enc.embedding_dict['item_emb'].detach().cpu().clone().requires_grad_(True)
    w1 = [enc.W['w1_%d' % k].detach().cpu().clone().requires_grad_(True) for k in range(2)]
    w2 = [enc.W['w2_%d' % k].detach().cpu().clone().requires_grad_(True) for k in range(2)]
    ref = port.ngcf_forward(adj, ue, ie, w1, w2)
    got = enc()
    torch.testing.assert_close(got[0].detach().cpu(), ref[0].detach(), rtol=1e-4, atol=1e-5)
    _loss(*ref).backward(); _loss(*got).backward()
    torch.testing.assert_close(enc.embedding_dict['item_emb'].grad.cpu(), ie.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(enc.W['w2_1'].grad.cpu(), w2[1].grad, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("d", [8, 16, 64, 128])
def test_spmm_row_and_column_masks(d):
    """Batch-sparse layers: row_mask = only those rows are computed/written; col_mask =
    rows of X outside it are zero and are skipped (same result as the dense product)."""
    from arlib_b200 import ops
    from arlib_b200.graph import DeviceGraph
    U, I = 700, 900
    u, i = _rand_graph(U, I, 15000, 9, 500)
    adj = port.bipartite_adjacency(u, i, U, I)
    g = DeviceGraph.from_dataloader_adj(adj, _dev())
    n = U + I
    rng = np.random.default_rng(0)
    live = rng.random(n) < 0.15
    live[:2] = True                                       # the two hub rows (long-row path)
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    for k in np.flatnonzero(live):
        words[k >> 5] |= np.uint32(1) << np.uint32(k & 31)
    mask = torch.from_numpy(words.view(np.int32)).to(_dev())
    X = torch.randn(n, d)
    Xz = X.clone(); Xz[~torch.from_numpy(live)] = 0
    ref = torch.sparse.mm(port.to_torch_coo(port.normalize_graph_mat(adj)), Xz)
    Y = torch.empty(n, d, device=_dev())
    ops.spmm(g, Xz.to(_dev()), Y=Y, col_mask=mask)
    torch.testing.assert_close(Y.cpu(), ref, rtol=1e-5, atol=1e-6)
    ref_full = torch.sparse.mm(port.to_torch_coo(port.normalize_graph_mat(adj)), X)
    out = torch.full((n, d), 7.0, device=_dev())
    ops.spmm(g, X.to(_dev()), acc_out=out, row_mask=mask)
    out = out.cpu()
    torch.testing.assert_close(out[torch.from_numpy(live)], ref_full[torch.from_numpy(live)], rtol=1e-5, atol=1e-6)
    assert bool((out[~torch.from_numpy(live)] == 7.0).all())


def test_ngcf_adjacency_gradient_matches_sparse_mm_autograd():
    """with NGCF as the victim, attack/White/PGA.py:98,117 sets
    ``model.sparse_norm_adj.requires_grad = True`` and calls ``torch.autograd.grad(Loss, model.sparse_norm_adj)``; the
    reference gets that through torch.sparse.mm (recommender/NGCF.py:197-212, two products per layer)."""
    from arlib_b200.encoder import NGCF_Encoder
    U, I, d = 100, 140, 64
    u, i = _rand_graph(U, I, 2000, 7)
    data = _Data(U, I, u, i)
    torch.manual_seed(2)
    enc = NGCF_Encoder(data, d, 2)
    ue = enc.embedding_dict['user_emb'].detach().cpu().clone().requires_grad_(True)
    ie = enc.embedding_dict['item_emb'].detach().cpu().clone().requires_grad_(True)
    w1 = [enc.W['w1_%d' % k].detach().cpu().clone() for k in range(2)]
    w2 = [enc.W['w2_%d' % k].detach().cpu().clone() for k in range(2)]
    adj = port.to_torch_coo(data.norm_adj).coalesce().requires_grad_(True)
    gr = torch.autograd.grad(_loss(*port.ngcf_forward(adj, ue, ie, w1, w2)), adj)[0].coalesce()
    enc.sparse_norm_adj.requires_grad = True
    ga = torch.autograd.grad(_loss(*enc()), enc.sparse_norm_adj)[0].coalesce()
    assert torch.equal(ga.indices().cpu(), gr.indices())
    torch.testing.assert_close(ga.values().cpu(), gr.values(), rtol=1e-3, atol=1e-4)
    # .backward() fills .grad like LightGCN's (recommender/NGCF.py:41-43,59-60: Matgrad += sparse_norm_adj.grad)
    _loss(*enc()).backward()
    torch.testing.assert_close(enc.sparse_norm_adj.grad.coalesce().values().cpu(), gr.values(), rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("segment", [64, 256])
@pytest.mark.parametrize("d", [8, 16])
def test_narrow_rows_cooperative_kernel(d, segment):
    """Column slices of the d-sharded tables (d = 8 / 16) run the warp-cooperative kernel (csrc/propagate.cu
    spmm_block_coop): items longer than one 64-entry index block, hub rows cut into many segments, the fused epilogue,
    the column-masked variant (bitmap in shared memory), per-batch work lists and run-to-run determinism."""
    from arlib_b200 import ops
    from arlib_b200.graph import DeviceGraph
    U, I = 1300, 1700
    u, i = _rand_graph(U, I, 40000, 12, 900)
    adj = port.bipartite_adjacency(u, i, U, I)
    norm = port.to_torch_coo(port.normalize_graph_mat(adj))
    g = DeviceGraph.from_dataloader_adj(adj, _dev()).replan(segment, segment)
    assert g.segment == segment and g.max_segments >= 3
    n = U + I
    X = torch.randn(n, d)
    ref = torch.sparse.mm(norm, X)
    Xd = X.to(_dev())
    Y = torch.empty_like(Xd)
    ops.spmm(g, Xd, Y=Y)
    torch.testing.assert_close(Y.cpu(), ref, rtol=1e-5, atol=2e-6)
    Y2 = torch.empty_like(Xd)
    ops.spmm(g, Xd, Y=Y2)
    assert torch.equal(Y, Y2)                                   # fixed summation order
    add, acc = torch.randn(n, d), torch.randn(n, d)
    accd = acc.to(_dev())
    ops.spmm(g, Xd, Y=Y2, addend=add.to(_dev()), acc_in=accd, acc_out=accd, acc_div=4.0)
    torch.testing.assert_close(Y2.cpu(), ref + add, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(accd.cpu(), (acc + ref + add) / 4.0, rtol=1e-5, atol=2e-6)
    # column mask: X is zero outside the live rows
    rng = np.random.default_rng(d)
    live = rng.random(n) < 0.13
    live[:3] = True
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    for k in np.flatnonzero(live):
        words[k >> 5] |= np.uint32(1) << np.uint32(k & 31)
    mask = torch.from_numpy(words.view(np.int32)).to(_dev())
    Xz = X.clone(); Xz[~torch.from_numpy(live)] = 0
    refz = torch.sparse.mm(norm, Xz)
    ops.spmm(g, Xz.to(_dev()), Y=Y2, addend=add.to(_dev()), col_mask=mask)
    torch.testing.assert_close(Y2.cpu(), refz + add, rtol=1e-5, atol=2e-6)
    # row mask: only the live rows are written
    out = torch.full((n, d), 7.0, device=_dev())
    ops.spmm(g, Xd, acc_out=out, row_mask=mask)
    out = out.cpu()
    torch.testing.assert_close(out[torch.from_numpy(live)], ref[torch.from_numpy(live)], rtol=1e-5, atol=2e-6)
    assert bool((out[~torch.from_numpy(live)] == 7.0).all())


def test_init_uiadj_same_pattern_reuses_the_plan_and_is_bit_equal():
    """SURVEY.md 8f-3 / attack/White/PGA.py:93-97: _init_uiAdj is re-entered once per 128-item batch with the same
    pattern (dense fractional fake rows) and new weights: the device indices and the SpMM plan are re-used, the values
    are bit-equal to a full rebuild; another pattern rebuilds."""
    from arlib_b200.encoder import LGCN_Encoder
    from arlib_b200.graph import DeviceGraph
    U, I, d = 120, 150, 64
    u, i = _rand_graph(U, I, 2500, 6)
    data = _Data(U, I, u, i)
    enc = LGCN_Encoder(data, d, 2)
    rng = np.random.default_rng(1)
    n = U + I

    have = set(zip(u.tolist(), i.tolist()))
    new_u, new_i = next((a, b) for a in range(U) for b in range(I) if (a, b) not in have)

    def adjacency(weights_seed, extra=False):
        w = np.random.default_rng(weights_seed).random(u.shape[0]).astype(np.float32)
        w[::9] = 1e-7
        uu, ii = (np.concatenate([u, [new_u]]), np.concatenate([i, [new_i]])) if extra else (u, i)
        ww = np.concatenate([w, [0.5]]).astype(np.float32) if extra else w
        half = sp.csr_matrix((ww, (uu, ii + U)), shape=(n, n), dtype=np.float32)
        return half + half.T

    a1, a2 = adjacency(1), adjacency(2)
    enc._init_uiAdj(a1)
    g1 = enc._graph
    enc._init_uiAdj(a2)
    g2 = enc._graph
    assert g2 is not g1 and g2.vrows.data_ptr() == g1.vrows.data_ptr() and g2.col.data_ptr() == g1.col.data_ptr()
    full = DeviceGraph.from_ui_adj(a2, _dev())
    assert torch.equal(g2.val, full.val) and not torch.equal(g2.val, g1.val)
    ref = port.to_torch_coo(port.init_uiadj_norm(a2)).coalesce()
    assert np.array_equal(g2.val.cpu().numpy().view(np.uint32), ref.values().numpy().view(np.uint32))
    X = torch.randn(n, d, device=_dev())
    Y1, Y2 = torch.empty_like(X), torch.empty_like(X)
    from arlib_b200 import ops
    ops.spmm(g2, X, Y=Y1); ops.spmm(full, X, Y=Y2)
    assert torch.equal(Y1, Y2)
    enc.sparse_norm_adj.requires_grad = True               # the COO view follows the new values
    assert torch.equal(enc.sparse_norm_adj.coalesce().values().detach(), g2.val)
    a3 = adjacency(3, extra=True)                          # one more edge: another pattern -> rebuilt
    enc._init_uiAdj(a3)
    assert enc._graph.nnz == g2.nnz + 2 and enc._graph.vrows.data_ptr() != g2.vrows.data_ptr()
    full3 = DeviceGraph.from_ui_adj(a3, _dev())
    assert torch.equal(enc._graph.val, full3.val) and torch.equal(enc._graph.col, full3.col)
