"""GPU integration parity: one epoch of LightGCN on the golden ml-100k fixture with
the reference's own triples, through the fused engine and through the drop-in
recommender class (embeddings within 1e-4 relative, metrics within 1e-3)."""
import random
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("use_graph", [False, True])
def test_engine_epoch_matches_reference_golden(golden, use_graph):
    from oracle import port
    from arlib_b200.engine import LightGCNEngine
    from arlib_b200.graph import DeviceGraph
    U, I = golden["user_names"].shape[0], golden["item_names"].shape[0]
    adj = port.bipartite_adjacency(golden["train_u"].astype(np.int64), golden["train_i"].astype(np.int64), U, I)
    g = DeviceGraph.from_dataloader_adj(adj, DEV)
    table = torch.cat([torch.from_numpy(golden["init_user_emb"]), torch.from_numpy(golden["init_item_emb"])]).to(DEV)
    T = golden["batch_u"].shape[0]
    eng = LightGCNEngine(g, table, U, 2, 0.005, 1e-4, 2048, T)
    eng.set_triples(golden["batch_u"], golden["batch_i"], golden["batch_j"])
    losses = eng.run_steps(0, use_graph=use_graph)[:, 0].cpu().numpy()
    np.testing.assert_allclose(losses, golden["batch_loss"], rtol=2e-5)
    assert _rel(table[:U].cpu(), torch.from_numpy(golden["param_user_emb"])) < 1e-4
    assert _rel(table[U:].cpu(), torch.from_numpy(golden["param_item_emb"])) < 1e-4
    F = eng.forward_table().cpu()
    assert _rel(F[:U], torch.from_numpy(golden["final_user_emb"])) < 1e-4
    assert int(eng.step_dev) == len(golden["batch_len"])


def test_engine_step_external_host_triples_matches_golden(golden):
    """The end-to-end entry (triples in HOST memory, graph replay per staging slot) consumes the
    reference's batches identically: same losses, same parameters after the epoch."""
    from oracle import port
    from arlib_b200.engine import LightGCNEngine
    from arlib_b200.graph import DeviceGraph
    U, I = golden["user_names"].shape[0], golden["item_names"].shape[0]
    adj = port.bipartite_adjacency(golden["train_u"].astype(np.int64), golden["train_i"].astype(np.int64), U, I)
    g = DeviceGraph.from_dataloader_adj(adj, DEV)
    table = torch.cat([torch.from_numpy(golden["init_user_emb"]), torch.from_numpy(golden["init_item_emb"])]).to(DEV)
    B = 2048
    eng = LightGCNEngine(g, table, U, 2, 0.005, 1e-4, B, B)
    bu, bi, bj = golden["batch_u"], golden["batch_i"], golden["batch_j"]
    lens = golden["batch_len"]
    losses, t0 = [], 0
    for nb in lens:
        nb = int(nb)
        host = torch.zeros((3, B), dtype=torch.int32)
        host[0, :nb] = torch.from_numpy(bu[t0:t0 + nb].astype(np.int32))
        host[1, :nb] = torch.from_numpy(bi[t0:t0 + nb].astype(np.int32))
        host[2, :nb] = torch.from_numpy(bj[t0:t0 + nb].astype(np.int32))
        row = eng.step_external(host, nb)
        torch.cuda.synchronize()
        losses.append(float(row[0]))
        t0 += nb
    np.testing.assert_allclose(np.array(losses), golden["batch_loss"], rtol=2e-5)
    assert _rel(table[:U].cpu(), torch.from_numpy(golden["param_user_emb"])) < 1e-4
    assert _rel(table[U:].cpu(), torch.from_numpy(golden["param_item_emb"])) < 1e-4
    assert int(eng.step_dev) == len(lens)


def _args(**kw):
    base = dict(topK="50", emb_size=64, n_layers=2, batch_size=2048, lRate=0.005, reg=1e-4, maxEpoch=1, seed=2018,
                sampler="host", model_name="LightGCN")
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_dropin_lightgcn_class_reproduces_reference_run(golden, golden_rows, capsys):
    """Same seed, host sampler -> the reference's triples; metrics within 1e-3; API shape."""
    from arlib_b200.recommender.LightGCN import LightGCN
    from arlib_b200.util.DataLoader import DataLoader
    train, test = golden_rows
    random.seed(2018); np.random.seed(2018); torch.manual_seed(2018)
    data = DataLoader.from_rows([list(r) for r in train], (), test)
    rec = LightGCN(_args(), data)
    # xavier init on the host generator: identical draws to the reference
    assert torch.equal(rec.model.embedding_dict["user_emb"].detach().cpu(), torch.from_numpy(golden["init_user_emb"]))
    assert torch.equal(rec.model.embedding_dict["item_emb"].detach().cpu(), torch.from_numpy(golden["init_item_emb"]))
    with pytest.raises(TypeError):
        rec.train(requires_grad=False)                   # ARLib.py:123-126 relies on this TypeError
    rec.train()
    np.testing.assert_allclose(rec.last_train_losses[:, 0].cpu().numpy(), golden["batch_loss"], rtol=2e-5)
    assert _rel(rec.model.embedding_dict["user_emb"].detach().cpu(), torch.from_numpy(golden["param_user_emb"])) < 1e-4
    rec_list, measure = rec.test()
    assert list(rec_list.keys()) == [str(u) for u in golden["topk_users"]]
    assert all(len(v) == 50 and isinstance(v[0][0], str) for v in rec_list.values())
    want = [str(x) for x in golden["measure"]]
    for a, b in zip(measure[1:], want[1:]):
        assert abs(float(a.split(":")[1]) - float(b.split(":")[1])) < 1e-3
    same = sum(set(int(p[0]) for p in rec_list[str(u)]) == set(golden["topk_items"][k].tolist())
               for k, u in enumerate(golden["topk_users"]))
    assert same >= 0.97 * len(golden["topk_users"])       # embeddings differ by ~1e-6: boundary swaps only
    s = rec.predict(str(golden["topk_users"][0]))
    assert isinstance(s, np.ndarray) and s.dtype == np.float32 and s.shape == (data.item_num,)
    assert rec.bestPerformance[0] == 1 and set(rec.bestPerformance[1]) == {"Hit Ratio", "Precision", "Recall", "NDCG"}


def test_dropin_general_path_external_optimizer_and_grad_export(golden_rows):
    """Caller-supplied optimizer (attack/White/CLeaR.py:59,72) and requires_embgrad /
    requires_adjgrad (recommender/LightGCN.py:36-43,74-80) go through autograd on the
    agcf kernels and match the fused path's first epoch."""
    import copy
    from arlib_b200.recommender.LightGCN import LightGCN
    from arlib_b200.util.DataLoader import DataLoader
    train, test = golden_rows
    train = [list(r) for r in train[:6000]]
    data = DataLoader.from_rows(train, (), test)
    torch.manual_seed(1)
    a = LightGCN(_args(), data)
    b = copy.deepcopy(a)
    random.seed(5); a.train()
    b.args = _args(fused=False)                           # the reference-shaped loop on autograd
    opt = torch.optim.Adam(b.model.parameters(), lr=0.005)
    random.seed(5); b.train(optimizer=opt)
    assert _rel(b.model.embedding_dict["user_emb"].detach(), a.model.embedding_dict["user_emb"].detach()) < 1e-4
    out = b.train(requires_embgrad=True, Epoch=1)
    assert len(out) == 4 and out[2].shape == (data.user_num, 64) and out[3].shape == (data.item_num, 64)
    assert float(out[2].abs().sum()) > 0 and float(out[3].abs().sum()) > 0
    mat = b.train(requires_adjgrad=True, Epoch=1)
    assert mat.shape == (data.user_num, data.item_num) and float(mat.abs().sum()) > 0
    c = copy.deepcopy(b)                                  # deepcopy / re-__init__ (ARLib.py:139,241)
    c.__init__(c.args, c.data)
    c.model._init_uiAdj(data.ui_adj)
    fu, fi = c.model()
    assert fu.shape == (data.user_num, 64) and fi.requires_grad


@pytest.mark.parametrize("sampler", ["host", "device"])
@pytest.mark.parametrize("name", ["NGCF", "SimGCL", "XSimGCL"])
def test_other_graph_recommenders_train_and_eval(name, sampler, golden_rows):
    """The three other drop-in classes run their reference loop on the agcf kernels:
    loss decreases, metrics improve over random, API shape holds."""
    import importlib
    from arlib_b200.util.DataLoader import DataLoader
    train, test = golden_rows
    data = DataLoader.from_rows([list(r) for r in train[:20000]], (), test)
    random.seed(3); torch.manual_seed(3)
    cls = getattr(importlib.import_module("arlib_b200.recommender." + name), name)
    rec = cls(_args(model_name=name, maxEpoch=2, sampler=sampler), data)
    rec.train(evalNum=1)
    rec_list, measure = rec.test()
    assert len(rec_list) == len(data.test_set) and measure[0] == "Top 50\n"
    recall = float(measure[3].split(":")[1])
    assert recall > 0.05, measure                      # random top-50 of ~1200 items would give ~0.04 at best
    out = rec.train(requires_embgrad=True, Epoch=1)
    assert len(out) == 4


def test_callers_plain_adam_runs_on_the_fused_engine_with_state_handover(golden_rows, monkeypatch):
    """attack/White/CLeaR.py:145-146 -- ``recommender.train(Epoch=innerEpoch, optimizer=optimizer)`` with the attacker's
    ``torch.optim.Adam(model.parameters())``: the fused engine stands in for optimizer.step(), starts from the optimizer's
    state and writes it back, so repeated calls continue the same Adam trajectory as the autograd loop with torch's own
    optimizer.  A stale optimizer (built on other tensors -- a quirk callers rely on) keeps the reference-shaped loop."""
    import copy
    from arlib_b200.recommender.LightGCN import LightGCN
    from arlib_b200.util.DataLoader import DataLoader
    train, test = golden_rows
    data = DataLoader.from_rows([list(r) for r in train[:6000]], (), test)
    torch.manual_seed(2)
    ref = LightGCN(_args(fused=False), data)
    fus = copy.deepcopy(ref)
    fus.args = _args()
    opt_ref = torch.optim.Adam(ref.model.parameters(), lr=0.003, betas=(0.8, 0.99), eps=1e-7)
    opt_fus = torch.optim.Adam(fus.model.parameters(), lr=0.003, betas=(0.8, 0.99), eps=1e-7)
    calls = []
    orig = LightGCN._train_fused
    monkeypatch.setattr(LightGCN, "_train_fused", lambda self, *a, **k: (calls.append(self), orig(self, *a, **k))[1])
    for rnd in range(2):                                   # two calls: the second starts from the handed-back state
        random.seed(7 + rnd); ref.train(Epoch=1, optimizer=opt_ref)
        random.seed(7 + rnd); fus.train(Epoch=1, optimizer=opt_fus)
    assert calls == [fus, fus]
    n_steps = 2 * ((len(data.training_data) + 2047) // 2048)
    for name in ("user_emb", "item_emb"):
        pr, pf = ref.model.embedding_dict[name], fus.model.embedding_dict[name]
        assert _rel(pf.detach(), pr.detach()) < 1e-4
        sr, sf = opt_ref.state[pr], opt_fus.state[pf]
        assert float(sf["step"]) == float(sr["step"]) == n_steps
        assert _rel(sf["exp_avg"], sr["exp_avg"]) < 1e-3 and _rel(sf["exp_avg_sq"], sr["exp_avg_sq"]) < 1e-3
    # the caller can keep using its optimizer afterwards
    fu, fi = fus.model()
    opt_fus.zero_grad(); (fu.sum() + fi.sum()).backward(); opt_fus.step()
    assert float(opt_fus.state[fus.model.embedding_dict["user_emb"]]["step"]) == n_steps + 1
    # stale optimizer: built on tensors that are not the model's parameters
    stale = torch.optim.Adam([torch.nn.Parameter(torch.zeros(3, 64, device=DEV)), torch.nn.Parameter(torch.zeros(3, 64, device=DEV))])
    n_before = len(calls)
    random.seed(1); fus.train(Epoch=1, optimizer=stale)
    assert len(calls) == n_before
    assert LightGCN._fusable_adam(torch.optim.SGD(fus.model.parameters(), lr=0.1), fus.model) is None
    assert LightGCN._fusable_adam(torch.optim.Adam(fus.model.parameters(), lr=0.1, amsgrad=True), fus.model) is None
