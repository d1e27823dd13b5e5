"""Freeze golden vectors for NGCF / SimGCL / XSimGCL / InfoNCE from the LIVE reference (builder container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden_models
Writes tests/golden/ml100k_{ngcf,simgcl,xsimgcl}.npz and tests/golden/infonce_kat.npz.

Per model: the UNMODIFIED reference class (recommender/NGCF.py:31-79,197-212; SimGCL.py:36-85,198-219;
XSimGCL.py:39-95,205-223) is run on the shipped ml-100k split with seedSet(2018) for one epoch + test(); recorded are
the initial parameters (incl. NGCF's W matrices), the triples the Python sampler produced, the per-batch losses
(total; rec / cl parts), the parameters after the epoch, the final (unperturbed) embeddings, all top-50 lists and the
metric strings.  The perturbation noise of SimGCL / XSimGCL is pinned by RE-SEEDING torch's CPU generator right before
train() (``torch.manual_seed(NOISE_SEED)``: the loop's only consumer of that generator is ``torch.rand_like`` --
SimGCL.py:204, XSimGCL.py:214); the stream is ``torch.rand(N, d)`` repeated, reproducible in the same image, and every
draw's float64 sum is frozen so a test can prove it regenerated the same tensors.

While freezing, every vector is re-derived with oracle/port.py (NGCFTrainer / SimGCLTrainer / XSimGCLTrainer,
ngcf_forward / simgcl_forward / xsimgcl_forward, infonce) and must agree: first-batch losses and every forward /
top-K / metric bit for bit, trained tensors to the run-to-run noise of CPU torch.
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys

import numpy as np
import torch

from oracle import port, ref_loader

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 2018
NOISE_SEED = 20180          # torch CPU generator state at the start of train()


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _names_to_int(names):
    return np.array([int(n) for n in names], dtype=np.int64)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _run_reference(model_name):
    """-> dict of golden arrays for one reference model (1 epoch + test on ml-100k)."""
    with ref_loader.reference_modules() as ref:
        args = ref_loader.make_args(ref, dataset="ml-100k", data_path=ref_loader.REF_ROOT + "/data/clean/",
                                    model_name=model_name, maxEpoch=1, n_layers=2, emb_size=64)
        ref.tool.seedSet(SEED)
        data = ref.DataLoader(args)
        mod = getattr(ref, model_name)
        with _quiet():
            rec = getattr(mod, model_name)(args, data)
        model = rec.model
        g = {"init_user_emb": model.embedding_dict["user_emb"].detach().clone().numpy(),
             "init_item_emb": model.embedding_dict["item_emb"].detach().clone().numpy()}
        if model_name == "NGCF":
            for k in range(args.n_layers):
                g["init_w1_%d" % k] = model.W["w1_%d" % k].detach().clone().numpy()
                g["init_w2_%d" % k] = model.W["w2_%d" % k].detach().clone().numpy()
            g["param_order"] = np.array([n for n, _ in model.named_parameters()])

        batches, totals, rec_losses, nce_losses, noise_sums = [], [], [], [], []
        real_sampler = mod.next_batch_pairwise

        def recording_sampler(d, bs):
            for b in real_sampler(d, bs):
                batches.append(tuple(list(x) for x in b))
                yield b

        real_bpr = mod.bpr_loss

        def rec_bpr(*a):
            out = real_bpr(*a)
            rec_losses.append(float(out.item()))
            return out

        real_backward = torch.Tensor.backward

        def rec_backward(self, *a, **k):
            totals.append(float(self.item()))
            return real_backward(self, *a, **k)

        real_rand_like = torch.rand_like

        def rec_rand_like(t, *a, **k):
            out = real_rand_like(t, *a, **k)
            noise_sums.append(float(out.double().sum()))
            return out

        patched = [(mod, "next_batch_pairwise", recording_sampler), (mod, "bpr_loss", rec_bpr),
                   (torch.Tensor, "backward", rec_backward), (torch, "rand_like", rec_rand_like)]
        if hasattr(mod, "InfoNCE"):
            real_nce = mod.InfoNCE

            def rec_nce(*a):
                out = real_nce(*a)
                nce_losses.append(float(out.item()))
                return out
            patched.append((mod, "InfoNCE", rec_nce))
        saved = [(o, n, getattr(o, n)) for o, n, _ in patched]
        for o, n, f in patched:
            setattr(o, n, f)
        torch.manual_seed(NOISE_SEED)
        try:
            with _quiet():
                rec.train()
        finally:
            for o, n, f in saved:
                setattr(o, n, f)
        with _quiet():
            rec_list, measure = rec.test()
        g.update(
            batch_len=np.array([len(b[0]) for b in batches], dtype=np.int32),
            batch_u=np.concatenate([np.array(b[0], dtype=np.int32) for b in batches]),
            batch_i=np.concatenate([np.array(b[1], dtype=np.int32) for b in batches]),
            batch_j=np.concatenate([np.array(b[2], dtype=np.int32) for b in batches]),
            batch_loss=np.array(totals, dtype=np.float64),
            rec_loss=np.array(rec_losses, dtype=np.float64),
            nce_loss=np.array(nce_losses, dtype=np.float64).reshape(len(batches), -1),     # [user side, item side] per batch
            noise_sum=np.array(noise_sums, dtype=np.float64),
            param_user_emb=model.embedding_dict["user_emb"].detach().numpy().copy(),
            param_item_emb=model.embedding_dict["item_emb"].detach().numpy().copy(),
            final_user_emb=rec.user_emb.detach().numpy().copy(),
            final_item_emb=rec.item_emb.detach().numpy().copy(),
            topk_users=_names_to_int(list(rec_list.keys())),
            topk_items=np.array([[int(p[0]) for p in rec_list[u]] for u in rec_list], dtype=np.int64),
            topk_scores=np.array([[p[1] for p in rec_list[u]] for u in rec_list], dtype=np.float32),
            measure=np.array(measure),
        )
        if model_name == "NGCF":
            for k in range(args.n_layers):
                g["param_w1_%d" % k] = model.W["w1_%d" % k].detach().numpy().copy()
                g["param_w2_%d" % k] = model.W["w2_%d" % k].detach().numpy().copy()
        hyper = {"lr": args.lRate, "reg": args.reg, "batch": args.batch_size, "topK": args.topK, "n_layers": args.n_layers}
        for k in ("n_layers", "cl_rate", "eps", "layer_cl", "temp"):
            if hasattr(rec, k):
                hyper[k] = getattr(rec, k)          # hard-coded in the reference class (SimGCL.py:31-33, XSimGCL.py:32-36)
        g["meta"] = np.array(["seed=%d" % SEED, "noise_seed=%d" % NOISE_SEED, "torch=" + torch.__version__,
                              "numpy=" + np.__version__] + ["%s=%r" % kv for kv in sorted(hyper.items())])
    return g, hyper


def _check_port(model_name, g, hyper, base):
    """the port must reproduce the frozen run"""
    names_u = [str(x) for x in base["user_names"]]
    names_i = [str(x) for x in base["item_names"]]
    rows = [[names_u[u], names_i[i], 1.0] for u, i in zip(base["train_u"], base["train_i"])]
    test_rows = [[str(u), str(i), 1.0] for u, i in zip(base["test_user_names"], base["test_item_names"])]
    pdata = port.PortData(rows, (), test_rows)
    N, d = pdata.user_num + pdata.item_num, 64
    random.seed(SEED)
    pb = list(port.next_batch_pairwise(pdata, hyper["batch"]))
    for key, col in (("batch_u", 0), ("batch_i", 1), ("batch_j", 2)):
        assert np.array_equal(np.concatenate([np.array(b[col]) for b in pb]), g[key]), key
    iu, ii = torch.from_numpy(g["init_user_emb"]), torch.from_numpy(g["init_item_emb"])
    torch.manual_seed(NOISE_SEED)
    drawn = []

    def noise():
        t = torch.rand(N, d)
        drawn.append(float(t.double().sum()))
        return t

    L = hyper["n_layers"]
    if model_name == "NGCF":
        w1 = [torch.from_numpy(g["init_w1_%d" % k]) for k in range(L)]
        w2 = [torch.from_numpy(g["init_w2_%d" % k]) for k in range(L)]
        tr = port.NGCFTrainer(pdata.norm_adj, iu, ii, w1, w2, hyper["lr"], hyper["reg"])
        tot = [tr.step(*b) for b in pb]
        assert tot[0] == g["batch_loss"][0], "first NGCF loss must be bit-equal"
        assert np.allclose(tot, g["batch_loss"], rtol=2e-6, atol=0)
        for k in range(L):
            assert np.abs(tr.w1[k].detach().numpy() - g["param_w1_%d" % k]).max() < 5e-6
            assert np.abs(tr.w2[k].detach().numpy() - g["param_w2_%d" % k]).max() < 5e-6
        fwd = lambda: port.ngcf_forward(port.to_torch_coo(pdata.norm_adj), torch.from_numpy(g["param_user_emb"]),
                                        torch.from_numpy(g["param_item_emb"]),
                                        [torch.from_numpy(g["param_w1_%d" % k]) for k in range(L)],
                                        [torch.from_numpy(g["param_w2_%d" % k]) for k in range(L)])
    else:
        if model_name == "SimGCL":
            tr = port.SimGCLTrainer(pdata.norm_adj, iu, ii, L, hyper["eps"], hyper["cl_rate"], hyper["lr"], hyper["reg"], noise)
        else:
            tr = port.XSimGCLTrainer(pdata.norm_adj, iu, ii, L, hyper["eps"], hyper["cl_rate"], hyper["layer_cl"],
                                     hyper["lr"], hyper["reg"], noise, tau=hyper["temp"])
        parts = [tr.step(*b) for b in pb]
        assert drawn == g["noise_sum"].tolist(), "the noise stream was not regenerated bit for bit"
        assert parts[0][0] == g["rec_loss"][0], "first rec loss must be bit-equal"
        cl0 = hyper["cl_rate"] * (torch.tensor(g["nce_loss"][0][0], dtype=torch.float32) +
                                  torch.tensor(g["nce_loss"][0][1], dtype=torch.float32))
        assert parts[0][1] == float(cl0), "first cl loss must be bit-equal"
        assert np.allclose([p[0] for p in parts], g["rec_loss"], rtol=2e-6, atol=0)
        assert np.allclose([p[1] for p in parts], hyper["cl_rate"] * g["nce_loss"].sum(1), rtol=2e-5, atol=0)
        fwd = lambda: port.simgcl_forward(port.to_torch_coo(pdata.norm_adj), torch.from_numpy(g["param_user_emb"]),
                                          torch.from_numpy(g["param_item_emb"]), L, hyper["eps"], None)
    err_u = np.abs(tr.user_emb.detach().numpy() - g["param_user_emb"]).max()
    err_i = np.abs(tr.item_emb.detach().numpy() - g["param_item_emb"]).max()
    # CPU torch training is not run-to-run deterministic (threaded index_put / sparse backward), and Adam's
    # m / (sqrt(v) + 1e-8) amplifies that on rows with ~1e-8 gradients: tolerance, not bits
    assert err_u < 5e-5 and err_i < 5e-5, (err_u, err_i)
    fu, fi = fwd()
    assert np.array_equal(_bits(fu.detach().numpy()), _bits(g["final_user_emb"]))
    assert np.array_equal(_bits(fi.detach().numpy()), _bits(g["final_item_emb"]))
    topk = [int(t) for t in str(hyper["topK"]).split(",")]
    prl, pmeasure = port.full_rank_test(pdata, fu.detach(), fi.detach(), max(topk), topk)
    assert list(pmeasure) == list(g["measure"]), (pmeasure, g["measure"])
    for k, u in enumerate(prl):
        assert int(u) == int(g["topk_users"][k])
        assert set(int(p[0]) for p in prl[u]) == set(g["topk_items"][k].tolist())
    return float(max(err_u, err_i))


def golden_models():
    base = np.load(os.path.join(GOLD, "ml100k_lightgcn.npz"), allow_pickle=False)
    for name in ("NGCF", "SimGCL", "XSimGCL"):
        g, hyper = _run_reference(name)
        # same seed, same shapes, user table drawn first: the initial embeddings and the sampler stream are those of
        # the LightGCN golden run -- checked here, then not stored twice
        assert np.array_equal(_bits(g["init_user_emb"]), _bits(base["init_user_emb"]))
        assert np.array_equal(_bits(g["init_item_emb"]), _bits(base["init_item_emb"]))
        for k in ("batch_len", "batch_u", "batch_i", "batch_j"):
            assert np.array_equal(g[k], base[k]), k
        err = _check_port(name, g, hyper, base)
        for k in ("init_user_emb", "init_item_emb", "batch_len", "batch_u", "batch_i", "batch_j"):
            del g[k]
        out = os.path.join(GOLD, "ml100k_%s.npz" % name.lower())
        np.savez_compressed(out, **g)
        print("%s written (%.1f MB); port == reference (trained tensors within %.1e);" % (
            os.path.basename(out), os.path.getsize(out) / 1e6, err), "".join(g["measure"]).replace("\n", " "))


def golden_infonce():
    """known-answer vectors of util/loss.py:42-49 (value and both input gradients) on seeded inputs."""
    g = {}
    cases = [(1, 64, 0.2, 11), (7, 64, 0.1, 12), (300, 64, 0.2, 13), (300, 64, 0.1, 14), (1500, 64, 0.2, 15), (257, 128, 0.1, 16),
             (64, 32, 0.2, 17)]
    with ref_loader.reference_modules() as ref:
        for n, d, tau, seed in cases:
            gen = torch.Generator().manual_seed(seed)
            v1 = (torch.rand(n, d, generator=gen) - 0.5).requires_grad_(True)
            v2 = (torch.rand(n, d, generator=gen) - 0.3).requires_grad_(True)
            loss = ref.loss.InfoNCE(v1, v2, tau)
            loss.backward()
            p1, p2 = v1.detach().clone().requires_grad_(True), v2.detach().clone().requires_grad_(True)
            ploss = port.infonce(p1, p2, tau)
            ploss.backward()
            assert float(ploss) == float(loss), "port InfoNCE must be bit-equal to the reference's"
            assert torch.equal(p1.grad, v1.grad) and torch.equal(p2.grad, v2.grad)
            key = "n%d_d%d_s%d" % (n, d, seed)
            g[key + "_loss"] = np.array([float(loss)], dtype=np.float64)
            if n <= 300:
                g[key + "_g1"] = v1.grad.numpy().copy()
                g[key + "_g2"] = v2.grad.numpy().copy()
            else:                     # large case: gradient checksums (row sums) keep the fixture small
                g[key + "_g1rows"] = v1.grad.double().sum(1).numpy()
                g[key + "_g2rows"] = v2.grad.double().sum(1).numpy()
    g["cases"] = np.array(cases, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "infonce_kat.npz"), **g)
    print("infonce_kat.npz written; port == reference bit-for-bit on %d cases" % len(cases))


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference not mounted; goldens can only be frozen in the builder container")
    golden_infonce()
    golden_models()
