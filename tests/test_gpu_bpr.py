"""GPU parity: Philox sampler (statistical), batch grouping, fused BPR forward /
backward, Adam -- against oracle/port.py (torch autograd on CPU)."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_sampler_every_edge_once_negatives_valid_and_uniform():
    from arlib_b200 import ops
    from arlib_b200.engine import DeviceTrainSet
    from arlib_b200.util.synth import synth_edges
    U, I, E = 500, 64, 9000
    tu, ti, _, _ = synth_edges(U, I, E, seed=1)
    ts = DeviceTrainSet.from_arrays(tu, ti, U, I, DEV)
    out = [torch.empty(E, dtype=torch.int32, device=DEV) for _ in range(3)]
    hist = np.zeros(I)
    seen_orders = []
    for epoch in range(4):
        ops.bpr_sample_epoch(ts.e_user, ts.e_item, ts.rej_rowptr, ts.rej_items, I, 1234, epoch, *out)
        u, i, j = (o.cpu().numpy().astype(np.int64) for o in out)
        # every edge exactly once per epoch (a permutation of the edge list)
        assert np.array_equal(np.sort(u * I + i), np.sort(tu * I + ti))
        seen_orders.append(u * I + i)
        # negatives are never train items of that user
        train = set((tu * I + ti).tolist())
        assert not any(k in train for k in (u * I + j).tolist())
        assert j.min() >= 0 and j.max() < I
        hist += np.bincount(j, minlength=I)
    assert not np.array_equal(seen_orders[0], seen_orders[1])          # epochs differ
    # determinism: same (seed, epoch) -> same triples
    ops.bpr_sample_epoch(ts.e_user, ts.e_item, ts.rej_rowptr, ts.rej_items, I, 1234, 3, *out)
    assert np.array_equal(out[0].cpu().numpy().astype(np.int64) * I + out[1].cpu().numpy(), seen_orders[3])
    # uniformity over the complement: chi-square of negatives for users with few train items is loose;
    # check the order is shuffled (rank correlation with input order ~ 0) and the j histogram is not degenerate
    pos = np.argsort(np.argsort(seen_orders[0]))
    assert abs(np.corrcoef(pos, np.arange(E))[0, 1]) < 0.1
    expected = hist.sum() / I
    assert hist.min() > 0.5 * expected * (1 - (np.bincount(ti, minlength=I) / U).max())


def test_sampler_uniform_over_complement_chi2():
    from arlib_b200 import ops
    from arlib_b200.engine import DeviceTrainSet
    U, I = 4, 50
    tu = np.repeat(np.arange(U), 5000)
    ti = np.tile(np.array([3, 7, 11, 13, 17]), U * 1000)           # duplicates: same 5 items per user
    ts = DeviceTrainSet.from_arrays(tu, ti, U, I, DEV)
    E = tu.shape[0]
    out = [torch.empty(E, dtype=torch.int32, device=DEV) for _ in range(3)]
    ops.bpr_sample_epoch(ts.e_user, ts.e_item, ts.rej_rowptr, ts.rej_items, I, 99, 0, *out)
    j = out[2].cpu().numpy()
    cnt = np.bincount(j, minlength=I).astype(np.float64)
    assert cnt[[3, 7, 11, 13, 17]].sum() == 0
    allowed = np.delete(cnt, [3, 7, 11, 13, 17])
    exp = E / allowed.shape[0]
    chi2 = ((allowed - exp) ** 2 / exp).sum()
    assert chi2 < 90.0            # 44 dof: p(chi2 > 90) ~ 5e-5


def _group_ref(u, i, j, U):
    nb = len(u)
    nodes = np.concatenate([u, U + i, U + j])
    order = np.lexsort((np.arange(3 * nb), nodes))
    return nodes[order], order


@pytest.mark.parametrize("nb,B", [(2048, 2048), (1203, 2048), (5, 8), (5000, 5461)])
def test_group_batches_is_a_sorted_segmentation(nb, B):
    from arlib_b200 import ops
    rng = np.random.default_rng(nb)
    U, I = 300, 500
    T = B + nb                                            # one full batch + one of nb triples
    u = rng.integers(0, U, T).astype(np.int32)
    i = rng.integers(0, I, T).astype(np.int32)
    j = rng.integers(0, I, T).astype(np.int32)
    d = lambda a: torch.from_numpy(a).to(DEV)
    nbat = 2
    occ = torch.full((nbat * 3 * B,), -1, dtype=torch.int32, device=DEV)
    seg_off = torch.full((nbat * (3 * B + 1),), -1, dtype=torch.int32, device=DEV)
    seg_node = torch.full((nbat * 3 * B,), -1, dtype=torch.int32, device=DEV)
    n_seg = torch.zeros(nbat, dtype=torch.int32, device=DEV)
    words = (U + I + 31) // 32
    mask = torch.full((nbat * words,), -1, dtype=torch.int32, device=DEV)
    ops.bpr_group_batches(d(u), d(i), d(j), T, B, U, occ, seg_off, seg_node, n_seg, U + I, mask)
    for b, (lo, n) in enumerate(((0, B), (B, nb))):
        nodes, order = _group_ref(u[lo:lo + n], i[lo:lo + n], j[lo:lo + n], U)
        got_occ = occ[b * 3 * B: b * 3 * B + 3 * n].cpu().numpy()
        assert np.array_equal(got_occ, order)
        ns = int(n_seg[b])
        uniq, first = np.unique(nodes, return_index=True)
        assert ns == uniq.shape[0]
        assert np.array_equal(seg_node[b * 3 * B: b * 3 * B + ns].cpu().numpy(), uniq)
        off = seg_off[b * (3 * B + 1): b * (3 * B + 1) + ns + 1].cpu().numpy()
        assert np.array_equal(off[:-1], first) and off[-1] == 3 * n
        bits = mask[b * words:(b + 1) * words].cpu().numpy().view(np.uint32)
        got_nodes = np.flatnonzero(np.unpackbits(bits.view(np.uint8), bitorder="little"))
        assert np.array_equal(got_nodes, uniq)


@pytest.mark.parametrize("d", [32, 64, 128, 256])
def test_bpr_forward_backward_match_autograd(d):
    from arlib_b200 import ops
    rng = np.random.default_rng(d)
    U, I, nb, B = 200, 300, 777, 1024
    F = (torch.randn(U + I, d) * 0.3)
    u = rng.integers(0, U, nb).astype(np.int32)
    i = rng.integers(0, 40, nb).astype(np.int32)              # heavy repetition of positives
    j = rng.integers(0, I, nb).astype(np.int32)
    Fr = F.clone().requires_grad_(True)
    ue, pe, ne = Fr[torch.from_numpy(u).long()], Fr[U + torch.from_numpy(i).long()], Fr[U + torch.from_numpy(j).long()]
    loss = port.bpr_loss(ue, pe, ne) + port.l2_reg_loss(1e-2, ue, pe)
    loss.backward()
    dv = lambda a: torch.from_numpy(a).to(DEV)
    Fd = F.to(DEV)
    out4 = torch.zeros(4, device=DEV)
    coef = torch.empty(B, device=DEV)
    ws = torch.zeros(ops.bpr_ws_bytes(B), dtype=torch.uint8, device=DEV)
    du, di, dj = dv(u), dv(i), dv(j)
    for rep in range(2):                                      # second call checks the ticket reset
        ops.bpr_forward(Fd, du, di, dj, nb, U, 1e-2, out4, coef, ws)
    assert abs(out4[0].item() - loss.item()) <= 2e-6 * abs(loss.item())
    assert abs(out4[2].item() - torch.norm(ue).item()) <= 1e-5 * torch.norm(ue).item()
    occ = torch.empty(3 * B, dtype=torch.int32, device=DEV)
    seg_off = torch.empty(3 * B + 1, dtype=torch.int32, device=DEV)
    seg_node = torch.empty(3 * B, dtype=torch.int32, device=DEV)
    n_seg = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.bpr_group_batches(du, di, dj, nb, B, U, occ, seg_off, seg_node, n_seg)
    G = torch.zeros_like(Fd)
    ops.bpr_backward(Fd, du, di, dj, nb, U, 1e-2, 1.0, out4, coef, occ, seg_off, seg_node, n_seg, G)
    torch.testing.assert_close(G.cpu(), Fr.grad, rtol=1e-4, atol=1e-7)
    ops.zero_rows(seg_node, n_seg, 3 * nb, G)
    assert float(G.abs().max()) == 0.0


def test_adam_matches_torch_optim():
    from arlib_b200 import ops
    torch.manual_seed(0)
    n = 64 * 37 + 3
    p0 = torch.randn(n)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=0.005)
    p, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step_dev = torch.zeros(1, dtype=torch.int32, device=DEV)
    for t in range(1, 6):
        g = torch.randn(n) * (10.0 ** -t)
        p_ref.grad = g.clone()
        opt.step()
        if t % 2:
            ops.adam_step(p, g.to(DEV), m, v, 0.005, step=t)
            ops.increment(step_dev)
        else:
            ops.adam_step(p, g.to(DEV), m, v, 0.005, step_dev=step_dev)
            ops.increment(step_dev)
        torch.testing.assert_close(p.cpu(), p_ref.detach(), rtol=2e-6, atol=1e-7)
    assert int(step_dev) == 5


@pytest.mark.parametrize("d,nb", [(64, 2048), (32, 77), (128, 1), (64, 5461)])
def test_bpr_l2_fused_autograd_op_matches_the_reference_expressions(d, nb):
    """util/loss.py:5-9,25-29 behind ONE autograd function (the loss line of the reference-shaped loops,
    recommender/LightGCN.py:51-54): value, both table gradients, and the upstream gradient scale."""
    from arlib_b200.util.loss import bpr_l2_fused
    U, I, reg = 300, 500, 1e-3
    gen = torch.Generator().manual_seed(nb)
    table = ((torch.rand(U + I, d, generator=gen) - 0.5) * 0.6).to("cuda:0")
    u = torch.randint(0, U, (nb,), generator=gen)
    i = torch.randint(0, I, (nb,), generator=gen)
    j = torch.randint(0, I, (nb,), generator=gen)
    ref_t = table.double().cpu().requires_grad_(True)
    ue, pe, ne = ref_t[:U][u], ref_t[U:][i], ref_t[U:][j]
    ref = port.bpr_loss(ue, pe, ne) + port.l2_reg_loss(reg, ue, pe)
    (ref * 1.7).backward()
    t = table.clone().requires_grad_(True)
    got, parts = bpr_l2_fused(t[:U], t[U:], u.tolist(), i.to("cuda:0"), j.tolist(), reg, return_parts=True)
    (got * 1.7).backward()
    assert abs(float(got) - float(ref)) <= 2e-6 * abs(float(ref))
    assert abs(float(parts[1]) - float(port.bpr_loss(ue, pe, ne))) <= 2e-6 * abs(float(ref))
    err = float((t.grad.double().cpu() - ref_t.grad).abs().max() / ref_t.grad.abs().max())
    assert err < 2e-6, err
