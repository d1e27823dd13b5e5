"""GPU parity of the fused SimGCL / XSimGCL training step (arlib_b200.engine.ContrastiveEngine) against the
oracle's restatement of recommender/SimGCL.py:46-64, 198-219 and recommender/XSimGCL.py:39-44, 56-75, 205-223
(oracle.port, torch CPU autograd + torch.optim.Adam) on the same graph, parameters, triples and NOISE (injected
tables on both sides).  Bars: losses 1e-5 relative, parameters after the steps 1e-4 relative."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
U, I, E, D, L, B = 300, 400, 6000, 64, 2, 512
EPS, CL_RATE, LR, REG = 0.1, 0.2, 0.005, 1e-4


def _setup(seed=3):
    from arlib_b200.graph import DeviceGraph
    from arlib_b200.util.synth import synth_edges
    tu, ti, _, _ = synth_edges(U, I, E, seed=seed)
    adj = port.bipartite_adjacency(tu, ti, U, I)
    norm = port.normalize_graph_mat(adj)
    g = DeviceGraph.from_dataloader_adj(adj, DEV)
    gen = torch.Generator().manual_seed(seed)
    ue = (torch.rand(U, D, generator=gen) - 0.5) * 0.2
    ie = (torch.rand(I, D, generator=gen) - 0.5) * 0.2
    rng = np.random.default_rng(seed)
    n_tr = 3 * B - 100                                           # two full batches and a short one
    bu = tu[:n_tr].astype(np.int32)
    bi = ti[:n_tr].astype(np.int32)
    bj = rng.integers(0, I, n_tr).astype(np.int32)
    noise = lambda: torch.rand(U + I, D, generator=gen)
    return g, port.to_torch_coo(norm), ue, ie, (bu, bi, bj), noise


def _oracle(kind, adj, ue, ie, triples, noises, tau, dtype=torch.float32):
    adj = adj.to(dtype)
    noises = {k: v.to(dtype) for k, v in noises.items()}
    ue, ie = ue.clone().to(dtype).requires_grad_(True), ie.clone().to(dtype).requires_grad_(True)
    opt = torch.optim.Adam([ue, ie], lr=LR)
    bu, bi, bj = triples
    out = []
    for lo in range(0, len(bu), B):
        u, i, j = (torch.from_numpy(x[lo:lo + B].astype(np.int64)) for x in (bu, bi, bj))
        uu, ii = torch.unique(u), torch.unique(i)
        if kind == "xsimgcl":
            ru, ri, cu, ci = port.xsimgcl_forward(adj, ue, ie, L, EPS, 1, [noises[(0, 1)], noises[(0, 2)]])
            cl = port.infonce(ru[uu], cu[uu], tau) + port.infonce(ri[ii], ci[ii], tau)
        else:
            ru, ri = port.simgcl_forward(adj, ue, ie, L, EPS, None)
            au, ai = port.simgcl_forward(adj, ue, ie, L, EPS, [noises[(1, 1)], noises[(1, 2)]])
            bu_, bi_ = port.simgcl_forward(adj, ue, ie, L, EPS, [noises[(2, 1)], noises[(2, 2)]])
            cl = port.infonce(au[uu], bu_[uu], tau) + port.infonce(ai[ii], bi_[ii], tau)
        rec = port.bpr_loss(ru[u], ri[i], ri[j])
        loss = rec + port.l2_reg_loss(REG, ru[u], ri[i]) + CL_RATE * cl
        opt.zero_grad()
        loss.backward()
        opt.step()
        out.append((float(rec), float(CL_RATE * cl)))
    return ue.detach(), ie.detach(), out


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("kind,tau", [("xsimgcl", 0.1), ("simgcl", 0.2)])
def test_contrastive_engine_matches_the_reference_loop(kind, tau, use_graph):
    from arlib_b200.engine import ContrastiveEngine
    g, adj, ue, ie, triples, noise = _setup()
    keys = [(0, 1), (0, 2)] if kind == "xsimgcl" else [(1, 1), (1, 2), (2, 1), (2, 2)]
    noises = {k: noise() for k in keys}
    ref_u, ref_i, ref_losses = _oracle(kind, adj, ue, ie, triples, noises, tau)
    # how well conditioned is "parameters after 3 Adam steps"?  The same loop in float64 tells: Adam divides by
    # sqrt(v) + 1e-8, elements whose gradient is ~1e-8 amplify rounding differences of ANY fp32 implementation
    # (the reference's own included) far above 1e-7.
    r64_u, r64_i, _ = _oracle(kind, adj, ue, ie, triples, noises, tau, torch.float64)
    want64 = torch.cat([r64_u, r64_i])
    ref_err = float((torch.cat([ref_u, ref_i]).double() - want64).abs().max() / want64.abs().max())
    table = torch.cat([ue, ie]).to(DEV)
    eng = ContrastiveEngine(g, table, U, kind, L, EPS, CL_RATE, tau, LR, REG, B, len(triples[0]),
                            noise_tables={k: v.to(DEV) for k, v in noises.items()})
    eng.set_triples(*triples)
    eng.run_steps(0, use_graph=use_graph)
    rec, cl = eng.losses()
    np.testing.assert_allclose(rec.cpu().numpy(), [x[0] for x in ref_losses], rtol=2e-5)
    np.testing.assert_allclose(cl.cpu().numpy(), [x[1] for x in ref_losses], rtol=2e-5)
    err = float((table.cpu().double() - want64).abs().max() / want64.abs().max())
    print("%s: engine vs fp64 reference %.3e, fp32 reference vs fp64 reference %.3e" % (kind, err, ref_err))
    assert err < 1e-4 + 2 * ref_err, (err, ref_err)
    assert int((eng.G != 0).sum()) == 0 and (kind != "xsimgcl" or int((eng.Gcl != 0).sum()) == 0)
    assert int(eng.step_dev) == 3
    # the unperturbed forward the evaluation reads
    fu, fi = port.simgcl_forward(adj, table[:U].cpu(), table[U:].cpu(), L, EPS, None)
    F = eng.forward_table(out=torch.empty_like(table)).cpu()
    assert float((F - torch.cat([fu, fi])).abs().max() / fu.abs().max()) < 1e-4


@pytest.mark.parametrize("cl_rate", [0.2, 0.0, 5.0])
@pytest.mark.parametrize("kind,tau,n_layers,layer_cl", [("xsimgcl", 0.1, 2, 1), ("simgcl", 0.2, 2, 0), ("xsimgcl", 0.1, 3, 1),
                                                        ("xsimgcl", 0.1, 3, 2), ("simgcl", 0.2, 3, 0), ("simgcl", 0.2, 1, 0)])
def test_first_batch_gradient_is_the_reference_gradient(kind, tau, n_layers, layer_cl, cl_rate):
    """The well-conditioned form of the parity check: dLoss/dE0 of one batch, recovered from Adam's first moment
    after one step (m = 0.1 g), against torch autograd through the reference expressions: 1e-5 of the largest entry.
    Also at depths the reference hard-codes away (n_layers = 1, 3; layer_cl = 2)."""
    from arlib_b200.engine import ContrastiveEngine
    nl = n_layers
    g, adj, ue, ie, triples, noise = _setup()
    triples = tuple(x[:B] for x in triples)
    passes = (0,) if kind == "xsimgcl" else (1, 2)
    noises = {(p, k): noise() for p in passes for k in range(1, nl + 1)}
    ue_, ie_ = ue.clone().requires_grad_(True), ie.clone().requires_grad_(True)
    u, i, j = (torch.from_numpy(x.astype(np.int64)) for x in triples)
    uu, ii = torch.unique(u), torch.unique(i)
    if kind == "xsimgcl":
        ru, ri, cu, ci = port.xsimgcl_forward(adj, ue_, ie_, nl, EPS, layer_cl, [noises[(0, k)] for k in range(1, nl + 1)])
        cl = port.infonce(ru[uu], cu[uu], tau) + port.infonce(ri[ii], ci[ii], tau)
    else:
        ru, ri = port.simgcl_forward(adj, ue_, ie_, nl, EPS, None)
        au, ai = port.simgcl_forward(adj, ue_, ie_, nl, EPS, [noises[(1, k)] for k in range(1, nl + 1)])
        bu_, bi_ = port.simgcl_forward(adj, ue_, ie_, nl, EPS, [noises[(2, k)] for k in range(1, nl + 1)])
        cl = port.infonce(au[uu], bu_[uu], tau) + port.infonce(ai[ii], bi_[ii], tau)
    (port.bpr_loss(ru[u], ri[i], ri[j]) + port.l2_reg_loss(REG, ru[u], ri[i]) + cl_rate * cl).backward()
    gref = torch.cat([ue_.grad, ie_.grad])
    table = torch.cat([ue, ie]).to(DEV)
    eng = ContrastiveEngine(g, table, U, kind, nl, EPS, cl_rate, tau, LR, REG, B, B, layer_cl=max(layer_cl, 1),
                            noise_tables={k: v.to(DEV) for k, v in noises.items()})
    eng.set_triples(*triples)
    eng.run_steps(0, 1, use_graph=False)
    got = (eng.m / 0.1).cpu()
    assert float((got - gref).abs().max() / gref.abs().max()) < 1e-5


def test_cl_id_lists_are_the_unique_users_and_positive_items():
    from arlib_b200 import ops
    rng = np.random.default_rng(1)
    nu, ni, Bt, T = 50, 70, 64, 64 + 37
    u = rng.integers(0, nu, T).astype(np.int32)
    i = rng.integers(0, ni, T).astype(np.int32)
    j = rng.integers(0, ni, T).astype(np.int32)
    d = lambda a: torch.from_numpy(a).to(DEV)
    occ = torch.empty(2 * 3 * Bt, dtype=torch.int32, device=DEV)
    seg_off = torch.empty(2 * (3 * Bt + 1), dtype=torch.int32, device=DEV)
    seg_node = torch.empty(2 * 3 * Bt, dtype=torch.int32, device=DEV)
    n_seg = torch.zeros(2, dtype=torch.int32, device=DEV)
    ops.bpr_group_batches(d(u), d(i), d(j), T, Bt, nu, occ, seg_off, seg_node, n_seg)
    cu = torch.full((2, Bt), -1, dtype=torch.int32, device=DEV)
    ci = torch.full((2, Bt), -1, dtype=torch.int32, device=DEV)
    n_cl = torch.zeros((2, 2), dtype=torch.int32, device=DEV)
    ops.bpr_cl_ids(occ, seg_off, seg_node, n_seg, T, Bt, nu, cu, ci, n_cl)
    for b, (lo, hi) in enumerate(((0, Bt), (Bt, T))):
        wu, wi = np.unique(u[lo:hi]), np.unique(i[lo:hi]) + nu
        assert n_cl[b].tolist() == [len(wu), len(wi)]
        assert np.array_equal(cu[b, :len(wu)].cpu().numpy(), wu)
        assert np.array_equal(ci[b, :len(wi)].cpu().numpy(), wi)


def test_philox_perturbation_has_the_reference_shape():
    """E' = E + sign(E) * normalize(noise) * eps with noise ~ U[0,1) drawn in the epilogue: every row moves by exactly
    eps in L2, every element moves away from zero, the draw depends on (seed, stream, step) only."""
    from arlib_b200 import ops
    g, _, ue, ie, _, _ = _setup()
    X = torch.cat([ue, ie]).to(DEV)
    clean = torch.empty_like(X)
    ops.spmm(g, X, Y=clean)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    outs = {}
    for key in ((7, 1, 0), (7, 1, 0), (7, 2, 0), (8, 1, 0), (7, 1, 5)):
        step.fill_(key[2])
        y = torch.empty_like(X)
        ops.spmm(g, X, Y=y, eps=EPS, philox=(key[0], key[1], step))
        outs.setdefault(key, []).append(y)
    a = outs[(7, 1, 0)]
    assert torch.equal(a[0], a[1])
    for other in ((7, 2, 0), (8, 1, 0), (7, 1, 5)):
        assert not torch.equal(a[0], outs[other][0])
    delta = a[0] - clean
    np.testing.assert_allclose(delta.norm(dim=1).cpu().numpy(), EPS, rtol=1e-4)
    nz = clean != 0
    assert bool((torch.sign(delta[nz]) == torch.sign(clean[nz])).all())
    # the implied noise direction is uniform-like: the normalized |delta| / eps has mean ~ E[u]/sqrt(d E[u^2]) = 0.866/8
    ratio = float((delta.abs() / EPS).mean())
    assert abs(ratio - 0.5 / np.sqrt(D / 3.0)) < 2e-3
