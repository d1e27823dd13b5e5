"""Freeze golden vectors from the LIVE reference (builder container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden
Writes tests/golden/*.npz.  Requires /root/reference (read-only); on the GPU box
the frozen files are used instead.  While freezing, every vector is also
re-derived with oracle/port.py and must be BIT-IDENTICAL -- that is the pin that
lets the port stand in for the reference where the reference cannot travel.
"""
from __future__ import annotations

import io
import os
import sys
import contextlib

import numpy as np
import torch

from oracle import port, ref_loader

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _names_to_int(names):
    return np.array([int(n) for n in names], dtype=np.int64)


def golden_ml100k_lightgcn(seed=2018, n_layers=2, emb=64):
    """reference LightGCN, shipped ml-100k split, seedSet(2018), 1 epoch + test()."""
    with ref_loader.reference_modules() as ref:
        args = ref_loader.make_args(ref, dataset="ml-100k", data_path=ref_loader.REF_ROOT + "/data/clean/",
                                    model_name="LightGCN", maxEpoch=1, n_layers=n_layers, emb_size=emb)
        ref.tool.seedSet(seed)
        data = ref.DataLoader(args)
        train_rows = [list(r) for r in data.training_data]          # file order, before any shuffle
        with _quiet():
            rec = ref.LightGCN.LightGCN(args, data)
        init_u = rec.model.embedding_dict["user_emb"].detach().clone()
        init_i = rec.model.embedding_dict["item_emb"].detach().clone()

        # record the triples and the per-batch loss the reference consumes
        batches, losses = [], []
        real_sampler = ref.LightGCN.next_batch_pairwise

        def recording_sampler(d, bs):
            for b in real_sampler(d, bs):
                batches.append(tuple(list(x) for x in b))
                yield b

        real_bpr = ref.LightGCN.bpr_loss
        real_l2 = ref.LightGCN.l2_reg_loss

        def rec_l2(reg, *a):
            out = real_l2(reg, *a)
            losses.append(out)          # the l2 term; the total is recorded below
            return out

        ref.LightGCN.next_batch_pairwise = recording_sampler
        totals = []
        real_backward = torch.Tensor.backward

        def rec_backward(self, *a, **k):
            totals.append(float(self.item()))
            return real_backward(self, *a, **k)

        torch.Tensor.backward = rec_backward
        try:
            with _quiet():
                rec.train()
        finally:
            torch.Tensor.backward = real_backward
            ref.LightGCN.next_batch_pairwise = real_sampler
        with _quiet():
            rec_list, measure = rec.test()

        norm_adj = data.norm_adj.tocsr()
        norm_adj.sort_indices()
        with _quiet():
            rec.model._init_uiAdj(data.ui_adj)
        uiadj_coo = rec.model.sparse_norm_adj.coalesce()

        g = dict(
            train_u=np.array([data.user[r[0]] for r in train_rows], dtype=np.int32),
            train_i=np.array([data.item[r[1]] for r in train_rows], dtype=np.int32),
            user_names=_names_to_int([data.id2user[k] for k in range(data.user_num)]),
            item_names=_names_to_int([data.id2item[k] for k in range(data.item_num)]),
            test_user_names=_names_to_int([u for u in data.test_set for _ in data.test_set[u]]),
            test_item_names=_names_to_int([i for u in data.test_set for i in data.test_set[u]]),
            adj_indptr=norm_adj.indptr.astype(np.int32), adj_indices=norm_adj.indices.astype(np.int32),
            adj_data=norm_adj.data.astype(np.float32),
            uiadj_row=uiadj_coo.indices()[0].numpy().astype(np.int32),
            uiadj_col=uiadj_coo.indices()[1].numpy().astype(np.int32),
            uiadj_data=uiadj_coo.values().numpy().astype(np.float32),
            init_user_emb=init_u.numpy(), init_item_emb=init_i.numpy(),
            batch_len=np.array([len(b[0]) for b in batches], dtype=np.int32),
            batch_u=np.concatenate([np.array(b[0], dtype=np.int32) for b in batches]),
            batch_i=np.concatenate([np.array(b[1], dtype=np.int32) for b in batches]),
            batch_j=np.concatenate([np.array(b[2], dtype=np.int32) for b in batches]),
            batch_loss=np.array(totals, dtype=np.float64),
            param_user_emb=rec.model.embedding_dict["user_emb"].detach().numpy().copy(),
            param_item_emb=rec.model.embedding_dict["item_emb"].detach().numpy().copy(),
            final_user_emb=rec.user_emb.detach().numpy().copy(),
            final_item_emb=rec.item_emb.detach().numpy().copy(),
            topk_users=_names_to_int(list(rec_list.keys())),
            topk_items=np.array([[int(p[0]) for p in rec_list[u]] for u in rec_list], dtype=np.int64),
            topk_scores=np.array([[p[1] for p in rec_list[u]] for u in rec_list], dtype=np.float32),
            measure=np.array(measure),
            meta=np.array(["seed=%d" % seed, "n_layers=%d" % n_layers, "emb=%d" % emb,
                           "batch=%d" % args.batch_size, "lr=%r" % args.lRate, "reg=%r" % args.reg,
                           "topK=%s" % args.topK, "torch=" + torch.__version__, "numpy=" + np.__version__]),
        )
        lr, reg, bs, topk = args.lRate, args.reg, args.batch_size, args.topK

    # ---- the port must reproduce all of it bit-for-bit ------------------------------
    import random
    names_u = [str(x) for x in g["user_names"]]
    names_i = [str(x) for x in g["item_names"]]
    rows = [[names_u[u], names_i[i], 1.0] for u, i in zip(g["train_u"], g["train_i"])]
    test_rows = [[str(u), str(i), 1.0] for u, i in zip(g["test_user_names"], g["test_item_names"])]
    pdata = port.PortData(rows, (), test_rows)
    assert pdata.user_num == len(names_u) and pdata.item_num == len(names_i)
    padj = pdata.norm_adj.tocsr()
    padj.sort_indices()
    assert np.array_equal(padj.indptr, g["adj_indptr"]) and np.array_equal(padj.indices, g["adj_indices"])
    assert np.array_equal(padj.data.view(np.uint32), g["adj_data"].view(np.uint32)), "norm_adj not bit-equal"
    pui = port.to_torch_coo(port.init_uiadj_norm(pdata.ui_adj)).coalesce()
    assert np.array_equal(pui.values().numpy().view(np.uint32), g["uiadj_data"].view(np.uint32))
    # host sampler: same python RNG state -> same triples.  seedSet() then the
    # reference draws nothing from `random` before the first shuffle.
    random.seed(seed)
    pb = list(port.next_batch_pairwise(pdata, bs))
    assert np.array_equal(np.concatenate([np.array(b[0]) for b in pb]), g["batch_u"])
    assert np.array_equal(np.concatenate([np.array(b[1]) for b in pb]), g["batch_i"])
    assert np.array_equal(np.concatenate([np.array(b[2]) for b in pb]), g["batch_j"])
    tr = port.LightGCNTrainer(pdata.norm_adj, torch.from_numpy(g["init_user_emb"]),
                              torch.from_numpy(g["init_item_emb"]), n_layers, lr, reg)
    plosses = [tr.step(*b) for b in pb]
    # CPU torch training is NOT run-to-run deterministic (multi-threaded
    # index_put/sparse backward accumulate in varying order; measured ~1e-7 abs
    # between two runs of the reference itself), so trained tensors are pinned
    # to a tolerance and everything upstream/downstream of them bit-exactly.
    assert plosses[0] == g["batch_loss"][0], "loss of batch 0 must be bit-equal"
    assert np.allclose(np.array(plosses), g["batch_loss"], rtol=1e-6, atol=0), "losses differ"
    assert np.abs(tr.user_emb.detach().numpy() - g["param_user_emb"]).max() < 2e-6
    assert np.abs(tr.item_emb.detach().numpy() - g["param_item_emb"]).max() < 2e-6
    # forward of the golden trained params is deterministic -> bit-equal
    fu, fi = port.lightgcn_forward(port.to_torch_coo(pdata.norm_adj), torch.from_numpy(g["param_user_emb"]),
                                   torch.from_numpy(g["param_item_emb"]), n_layers)
    assert np.array_equal(fu.numpy().view(np.uint32), g["final_user_emb"].view(np.uint32))
    assert np.array_equal(fi.numpy().view(np.uint32), g["final_item_emb"].view(np.uint32))
    prl, pmeasure = port.full_rank_test(pdata, fu, fi, max(int(t) for t in topk.split(",")),
                                        [int(t) for t in topk.split(",")])
    assert list(pmeasure) == list(g["measure"]), (pmeasure, g["measure"])
    for k, u in enumerate(prl):
        assert int(u) == int(g["topk_users"][k])
        assert set(int(p[0]) for p in prl[u]) == set(g["topk_items"][k].tolist())
        assert np.array_equal(np.sort(np.array([p[1] for p in prl[u]], dtype=np.float32)),
                              np.sort(g["topk_scores"][k]))
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "ml100k_lightgcn.npz"), **g)
    print("ml100k_lightgcn.npz written; port == reference bit-for-bit;", "".join(g["measure"]).replace("\n", " "))


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference not mounted; goldens can only be frozen in the builder container")
    golden_ml100k_lightgcn()
