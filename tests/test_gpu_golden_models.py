"""GPU parity against golden vectors frozen from the LIVE reference (oracle/make_golden_models.py): one ml-100k epoch
of NGCF / SimGCL / XSimGCL with the reference's own initial parameters, triples and perturbation noise, and known-answer
vectors of InfoNCE.  Reference: recommender/NGCF.py:31-79,197-212; SimGCL.py:36-85,198-219; XSimGCL.py:39-95,205-223;
util/loss.py:42-49.  Bars: losses 2e-5 relative, embeddings after the epoch 1e-4 relative, metrics 1e-3."""
import os
import random
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NOISE_SEED = 20180


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def _meta(g):
    return dict(str(s).split("=", 1) for s in g["meta"])


def _args(**kw):
    base = dict(topK="50", emb_size=64, n_layers=2, batch_size=2048, lRate=0.005, reg=1e-4, maxEpoch=1, seed=2018,
                sampler="host", model_name="LightGCN")
    base.update(kw)
    return types.SimpleNamespace(**base)


def _noise_stream(n, d, g):
    """the reference's torch.rand_like stream of the frozen run: CPU generator seeded with NOISE_SEED"""
    gen = torch.Generator().manual_seed(NOISE_SEED)
    sums = g["noise_sum"].tolist()
    k = [0]

    def draw():
        t = torch.rand(n, d, generator=gen)
        assert float(t.double().sum()) == sums[k[0]], "noise stream not regenerated bit for bit"
        k[0] += 1
        return t
    return draw, k


def _fp64_yardstick(kind, g, golden):
    """``parameters after an epoch of Adam`` is ill-conditioned for the contrastive models: Adam divides by sqrt(v) + 1e-8
    and rows whose gradient is ~1e-8 amplify the rounding of ANY fp32 implementation, the reference's own included.  The
    same loop in float64 (oracle.port, pinned to the reference in fp32) tells how far the frozen fp32 reference run is
    from exact arithmetic; an implementation is on the bar if it is not further away than that (plus 1e-4).
    -> (fp64 user table, fp64 item table, error of the reference's fp32 run against them)"""
    from oracle import port
    hy = _meta(g)
    U, I = golden["user_names"].shape[0], golden["item_names"].shape[0]
    adj = port.normalize_graph_mat(port.bipartite_adjacency(golden["train_u"].astype(np.int64), golden["train_i"].astype(np.int64), U, I))
    draw, _ = _noise_stream(U + I, 64, g)
    iu, ii = torch.from_numpy(golden["init_user_emb"]), torch.from_numpy(golden["init_item_emb"])
    L, eps, cl_rate = int(hy["n_layers"]), float(hy["eps"]), float(hy["cl_rate"])
    if kind == "simgcl":
        tr = port.SimGCLTrainer(adj, iu, ii, L, eps, cl_rate, float(hy["lr"]), float(hy["reg"]), draw, dtype=torch.float64)
    else:
        tr = port.XSimGCLTrainer(adj, iu, ii, L, eps, cl_rate, int(hy["layer_cl"]), float(hy["lr"]), float(hy["reg"]), draw,
                                 tau=float(hy["temp"]), dtype=torch.float64)
    off = np.concatenate([[0], np.cumsum(golden["batch_len"])])
    for b in range(len(off) - 1):
        sl = slice(off[b], off[b + 1])
        tr.step(golden["batch_u"][sl].tolist(), golden["batch_i"][sl].tolist(), golden["batch_j"][sl].tolist())
    wu, wi = tr.user_emb.detach(), tr.item_emb.detach()
    ref_err = max(_rel(g["param_user_emb"], wu), _rel(g["param_item_emb"], wi))
    return wu, wi, ref_err


def _check_metrics(rec, g, min_same=0.97):
    rec_list, measure = rec.test()
    assert list(rec_list.keys()) == [str(u) for u in g["topk_users"]]
    for a, b in zip(measure[1:], [str(x) for x in g["measure"]][1:]):
        assert a.split(":")[0] == b.split(":")[0]
        assert abs(float(a.split(":")[1]) - float(b.split(":")[1])) < 1e-3, (a, b)
    same = sum(set(int(p[0]) for p in rec_list[str(u)]) == set(g["topk_items"][k].tolist())
               for k, u in enumerate(g["topk_users"]))
    assert same >= min_same * len(g["topk_users"]), same       # trained tables differ by ~1e-6: boundary swaps only


def _topk_sets_from_golden_tables(g, golden, golden_rows):
    """final embeddings of the frozen run -> our evaluator returns the reference's top-50 SETS, except where the
    reference's own scores (torch.matmul on the CPU: another summation order than the device's k-ascending fmaf chain)
    put the 50th / 51st item within a few ulp of each other: such users are counted as boundary near-ties
    and every swapped item must sit on that boundary; the metric strings agree to 1e-4 (bar: 1e-3)"""
    from arlib_b200 import ops
    from arlib_b200.evaluator import FullRankEvaluator
    from arlib_b200.util.DataLoader import DataLoader
    train, test = golden_rows
    data = DataLoader.from_rows([list(r) for r in train], (), test)
    ev = FullRankEvaluator(data, DEV)
    fu, fi = torch.from_numpy(g["final_user_emb"]).to(DEV), torch.from_numpy(g["final_item_emb"]).to(DEV)
    rec_list, measure = ev.test(fu, fi, [50], 50)
    for a, b in zip(measure[1:], [str(x) for x in g["measure"]][1:]):
        assert a.split(":")[0] == b.split(":")[0] and abs(float(a.split(":")[1]) - float(b.split(":")[1])) < 1e-4, (a, b)
    near_ties = 0
    for k, u in enumerate(g["topk_users"]):
        got, want = set(int(p[0]) for p in rec_list[str(u)]), set(g["topk_items"][k].tolist())
        if got == want:
            continue
        near_ties += 1
        uid = data.user[str(u)]
        s_ = ops.score_rows(fu, torch.tensor([uid], dtype=torch.int32, device=DEV), fi)[0].cpu().numpy()
        kth = float(np.sort(g["topk_scores"][k])[0])                      # the reference's 50th score
        tol = 8 * np.finfo(np.float32).eps * float(fu[uid].norm() * fi.norm(dim=1).max())
        for name in got ^ want:
            assert abs(float(s_[data.item[str(name)]]) - kth) <= tol, "user %s: item %s is not a boundary near-tie" % (u, name)
    assert near_ties <= 0.01 * len(g["topk_users"]), near_ties
    return near_ties


@pytest.mark.parametrize("name", ["ngcf", "simgcl", "xsimgcl"])
def test_reference_final_tables_give_the_reference_topk_sets(name, golden, golden_rows):
    g = np.load(os.path.join(GOLD, "ml100k_%s.npz" % name), allow_pickle=False)
    _topk_sets_from_golden_tables(g, golden, golden_rows)


def test_ngcf_class_reproduces_the_reference_epoch(golden, golden_rows):
    from arlib_b200.recommender.NGCF import NGCF
    from arlib_b200.util.DataLoader import DataLoader
    g = np.load(os.path.join(GOLD, "ml100k_ngcf.npz"), allow_pickle=False)
    train, test = golden_rows
    random.seed(2018); np.random.seed(2018); torch.manual_seed(2018)
    data = DataLoader.from_rows([list(r) for r in train], (), test)
    rec = NGCF(_args(model_name="NGCF"), data)
    m = rec.model
    # same host-generator draws as the reference: tables first, then w1_k, w2_k per layer
    assert torch.equal(m.embedding_dict["user_emb"].detach().cpu(), torch.from_numpy(golden["init_user_emb"]))
    for k in range(2):
        assert torch.equal(m.W["w1_%d" % k].detach().cpu(), torch.from_numpy(g["init_w1_%d" % k]))
        assert torch.equal(m.W["w2_%d" % k].detach().cpu(), torch.from_numpy(g["init_w2_%d" % k]))
    # forward of the reference's TRAINED parameters (deterministic on both sides)
    with torch.no_grad():
        saved = {n: p.detach().clone() for n, p in m.named_parameters()}
        m.embedding_dict["user_emb"].copy_(torch.from_numpy(g["param_user_emb"]))
        m.embedding_dict["item_emb"].copy_(torch.from_numpy(g["param_item_emb"]))
        for k in range(2):
            m.W["w1_%d" % k].copy_(torch.from_numpy(g["param_w1_%d" % k]))
            m.W["w2_%d" % k].copy_(torch.from_numpy(g["param_w2_%d" % k]))
        fu, fi = m()
        assert _rel(fu, g["final_user_emb"]) < 1e-5 and _rel(fi, g["final_item_emb"]) < 1e-5
        for n, p in m.named_parameters():
            p.copy_(saved[n])
    rec.train()                                            # the fused NGCFEngine (the recommender owns the optimizer)
    np.testing.assert_allclose(rec.last_train_losses[:, 0].cpu().numpy(), g["batch_loss"], rtol=2e-5)
    eu = _rel(m.embedding_dict["user_emb"].detach(), g["param_user_emb"])
    ei = _rel(m.embedding_dict["item_emb"].detach(), g["param_item_emb"])
    ew = max(_rel(m.W["w%d_%d" % (a, k)].detach(), g["param_w%d_%d" % (a, k)]) for a in (1, 2) for k in range(2))
    print("NGCF epoch (fused engine) vs reference: user %.2e item %.2e W %.2e" % (eu, ei, ew))
    assert eu < 1e-4 and ei < 1e-4 and ew < 1e-4
    _check_metrics(rec, g)


def test_ngcf_class_reference_loop_reproduces_the_reference_epoch(golden, golden_rows):
    """the same epoch on the reference-shaped loop (torch autograd over the agcf SpMM + the fused BPR op + torch Adam):
    what a caller's optimizer / gradient export gets"""
    from arlib_b200.recommender.NGCF import NGCF
    from arlib_b200.util.DataLoader import DataLoader
    g = np.load(os.path.join(GOLD, "ml100k_ngcf.npz"), allow_pickle=False)
    train, test = golden_rows
    random.seed(2018); np.random.seed(2018); torch.manual_seed(2018)
    data = DataLoader.from_rows([list(r) for r in train], (), test)
    rec = NGCF(_args(model_name="NGCF", fused=False), data)
    m = rec.model
    losses = []
    real_backward = torch.Tensor.backward

    def rec_backward(self, *a, **k):
        losses.append(self.detach())
        return real_backward(self, *a, **k)
    torch.Tensor.backward = rec_backward
    try:
        rec.train()
    finally:
        torch.Tensor.backward = real_backward
    np.testing.assert_allclose(torch.stack(losses).cpu().numpy(), g["batch_loss"], rtol=2e-5)
    eu = _rel(m.embedding_dict["user_emb"].detach(), g["param_user_emb"])
    ei = _rel(m.embedding_dict["item_emb"].detach(), g["param_item_emb"])
    ew = max(_rel(m.W["w%d_%d" % (a, k)].detach(), g["param_w%d_%d" % (a, k)]) for a in (1, 2) for k in range(2))
    print("NGCF epoch (reference-shaped loop) vs reference: user %.2e item %.2e W %.2e" % (eu, ei, ew))
    assert eu < 1e-4 and ei < 1e-4 and ew < 1e-4
    _check_metrics(rec, g)


@pytest.mark.parametrize("name", ["SimGCL", "XSimGCL"])
def test_contrastive_class_reference_loop_reproduces_the_reference_epoch(name, golden, golden_rows):
    """the drop-in class on its reference-shaped loop (autograd over the agcf kernels), host sampler, the
    reference's noise stream injected through the encoder's noise hook"""
    import importlib
    from arlib_b200.util.DataLoader import DataLoader
    g = np.load(os.path.join(GOLD, "ml100k_%s.npz" % name.lower()), allow_pickle=False)
    train, test = golden_rows
    random.seed(2018); np.random.seed(2018); torch.manual_seed(2018)
    data = DataLoader.from_rows([list(r) for r in train], (), test)
    cls = getattr(importlib.import_module("arlib_b200.recommender." + name), name)
    rec = cls(_args(model_name=name, fused=False), data)
    assert torch.equal(rec.model.embedding_dict["item_emb"].detach().cpu(), torch.from_numpy(golden["init_item_emb"]))
    hy = _meta(g)
    assert (rec.n_layers, rec.cl_rate, rec.eps) == (int(hy["n_layers"]), float(hy["cl_rate"]), float(hy["eps"]))
    draw, count = _noise_stream(data.user_num + data.item_num, 64, g)
    rec.model.noise_source = lambda k, like: draw().to(like.device)
    rec.train()
    assert count[0] == len(g["noise_sum"])
    eu = _rel(rec.model.embedding_dict["user_emb"].detach(), g["param_user_emb"])
    ei = _rel(rec.model.embedding_dict["item_emb"].detach(), g["param_item_emb"])
    wu, wi, ref_err = _fp64_yardstick(name.lower(), g, golden)
    e64 = max(_rel(rec.model.embedding_dict["user_emb"].detach(), wu), _rel(rec.model.embedding_dict["item_emb"].detach(), wi))
    print("%s reference-shaped loop vs reference: user %.2e item %.2e | vs the fp64 loop %.2e (the reference's own fp32 run: %.2e)"
          % (name, eu, ei, e64, ref_err))
    assert (eu < 1e-4 and ei < 1e-4) or e64 < 1e-4 + 2 * ref_err
    assert _rel(rec.user_emb, g["final_user_emb"]) < 5e-4 and _rel(rec.item_emb, g["final_item_emb"]) < 5e-4
    _check_metrics(rec, g)


@pytest.mark.parametrize("kind", ["simgcl", "xsimgcl"])
def test_contrastive_engine_reproduces_the_reference_epoch(kind, golden):
    """the FUSED engine (shared first layer, one backward propagation, Adam in the last SpMM) stepped through the
    reference's 22 batches with the reference's noise tables"""
    from oracle import port
    from arlib_b200.engine import ContrastiveEngine
    from arlib_b200.graph import DeviceGraph
    g = np.load(os.path.join(GOLD, "ml100k_%s.npz" % kind), allow_pickle=False)
    hy = _meta(g)
    L, eps, cl_rate = int(hy["n_layers"]), float(hy["eps"]), float(hy["cl_rate"])
    tau = float(hy["temp"]) if kind == "xsimgcl" else 0.2
    U, I = golden["user_names"].shape[0], golden["item_names"].shape[0]
    adj = port.bipartite_adjacency(golden["train_u"].astype(np.int64), golden["train_i"].astype(np.int64), U, I)
    graph = DeviceGraph.from_dataloader_adj(adj, DEV)
    table = torch.cat([torch.from_numpy(golden["init_user_emb"]), torch.from_numpy(golden["init_item_emb"])]).to(DEV)
    T = golden["batch_u"].shape[0]
    draw, count = _noise_stream(U + I, 64, g)
    passes = (0,) if kind == "xsimgcl" else (1, 2)
    tabs = {(p, k): torch.empty((U + I, 64), device=DEV) for p in passes for k in range(1, L + 1)}
    eng = ContrastiveEngine(graph, table, U, kind, L, eps, cl_rate, tau, float(hy["lr"]), float(hy["reg"]), 2048, T,
                            layer_cl=int(hy.get("layer_cl", 1)), noise_tables=tabs)
    eng.set_triples(golden["batch_u"], golden["batch_i"], golden["batch_j"])
    for b in range(eng.n_batches):
        for p in passes:                          # reference order: pass 1 layers 1..L, then pass 2 layers 1..L
            for k in range(1, L + 1):
                tabs[(p, k)].copy_(draw())
        eng.run_steps(b, 1, use_graph=False)
    assert count[0] == len(g["noise_sum"]) and int(eng.step_dev) == len(golden["batch_len"])
    rec, cl = eng.losses()
    np.testing.assert_allclose(rec.cpu().numpy(), g["rec_loss"], rtol=2e-5)
    np.testing.assert_allclose(cl.cpu().numpy(), cl_rate * g["nce_loss"].sum(1), rtol=5e-5)
    eu, ei = _rel(table[:U], g["param_user_emb"]), _rel(table[U:], g["param_item_emb"])
    wu, wi, ref_err = _fp64_yardstick(kind, g, golden)
    e64 = max(_rel(table[:U], wu), _rel(table[U:], wi))
    print("%s fused engine vs reference: user %.2e item %.2e | vs the fp64 loop %.2e (the reference's own fp32 run: %.2e)"
          % (kind, eu, ei, e64, ref_err))
    # Where the bar sits for the FUSED SimGCL step: it adds the gradients of the three views into one table and runs ONE
    # backward propagation (autograd runs three and adds at E0) -- the same mathematics in another summation order, i.e.
    # ~1e-11 absolute differences in the gradient, which Adam's 1 / (sqrt(v) + 1e-8) turns into ~5e-6 per step on the
    # few elements whose gradient is below 1e-8 (they move in the regime update = lr * g / 1e-8).  So: the bulk of the
    # table within 1e-4, and no element further than 1e-3.
    diff = (table.cpu().double() - torch.cat([wu, wi])).abs() / float(torch.cat([wu, wi]).abs().max())
    frac_off = float((diff > 1e-4).double().mean())
    print("   elements further than 1e-4 from the fp64 loop: %.2e of the table, worst %.2e" % (frac_off, float(diff.max())))
    assert (eu < 1e-4 and ei < 1e-4) or e64 < 1e-4 + 2 * ref_err or (frac_off < 1e-3 and float(diff.max()) < 1e-3)
    F = eng.forward_table(out=torch.empty_like(table))
    assert _rel(F[:U], g["final_user_emb"]) < 5e-4 and _rel(F[U:], g["final_item_emb"]) < 5e-4


def test_infonce_known_answers_on_device():
    from arlib_b200.util.loss import InfoNCE
    g = np.load(os.path.join(GOLD, "infonce_kat.npz"), allow_pickle=False)
    for n, d, tau, seed in g["cases"]:
        n, d, seed = int(n), int(d), int(seed)
        gen = torch.Generator().manual_seed(seed)
        v1 = (torch.rand(n, d, generator=gen) - 0.5).to(DEV).requires_grad_(True)
        v2 = (torch.rand(n, d, generator=gen) - 0.3).to(DEV).requires_grad_(True)
        loss = InfoNCE(v1, v2, float(tau))
        loss.backward()
        key = "n%d_d%d_s%d" % (n, d, seed)
        want = float(g[key + "_loss"][0])
        assert abs(float(loss) - want) <= 1e-5 * max(abs(want), 1e-3), (key, float(loss), want)
        if key + "_g1" in g:
            for got, ref in ((v1.grad, g[key + "_g1"]), (v2.grad, g[key + "_g2"])):
                ref = torch.from_numpy(ref)
                assert float((got.cpu() - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-9, key
        else:
            for got, ref in ((v1.grad, g[key + "_g1rows"]), (v2.grad, g[key + "_g2rows"])):
                np.testing.assert_allclose(got.double().sum(1).cpu().numpy(), ref, rtol=0,
                                           atol=2e-5 * float(np.abs(ref).max()) + 1e-9)
