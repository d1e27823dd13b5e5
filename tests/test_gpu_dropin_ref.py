"""The drop-in claim, proved with the reference's OWN callers: the unmodified
``ARLib`` driver (ARLib.py:94-257) and the attack classes CLeaR (attack/White/CLeaR.py:56-159), GTA with its
``proxyLG(LightGCN)`` subclass (attack/Black/GTA.py:57-236) and PGA (attack/White/PGA.py:54-140) run on a B200 against
``arlib_b200.recommender.*``.  The reference tree is the staged byte-for-byte copy ``oracle/_ref`` (oracle/make_ref.py);
only ``recommender.{LightGCN,NGCF,SimGCL,XSimGCL}`` are re-pointed at this package (the import switch of INTEGRATION.md
section 2, done in ``sys.modules``).  The same flows are also run on the reference's own recommender classes on the same
GPU (their eager PyTorch + cuSPARSE path) for the metric comparison."""
import contextlib
import copy
import io
import os
import random
import shutil

import numpy as np
import pytest
import torch

from oracle import ref_loader

CUDA = torch.cuda.is_available()      # (False only when the harness itself is exercised on the reference's CPU path)

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="reference copy oracle/_ref not staged")]


@contextlib.contextmanager
def _workdir(tmp_path):
    """the reference reads ./data/clean/<ds>/ and writes ./log, ./data/poison, ./modelsaved relative to cwd"""
    dst = tmp_path / "data" / "clean" / "ml-100k"
    dst.mkdir(parents=True, exist_ok=True)
    for f in ("train.txt", "val.txt", "test.txt"):
        shutil.copyfile(os.path.join(ref_loader.REF_ROOT, "data", "clean", "ml-100k", f), dst / f)
    old = os.getcwd()
    os.chdir(tmp_path)
    try:
        yield
    finally:
        os.chdir(old)


@contextlib.contextmanager
def _env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update(kw)
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _values(measure):
    return {m.split(":")[0]: float(m.split(":")[1]) for m in measure if ":" in m}


def _run_driver(dropin, model_name, tmp_path, max_epoch=2, save=False):
    """main.py:16-51 with NoneAttack, times = 1 -- the reference's own statements, in order"""
    out = {}
    with _workdir(tmp_path), ref_loader.reference_modules(cuda=CUDA, dropin=dropin, callers=True) as ref:
        rargs = ref_loader.make_args(ref, dataset="ml-100k", model_name=model_name, maxEpoch=max_epoch, save=save)
        aargs = ref_loader.make_attack_args(ref, attackCategory="Black", attackModelName="NoneAttack", times=1)
        ref.tool.seedSet(rargs.seed)
        data = ref.DataLoader(rargs)
        rec_cls = getattr(getattr(ref, model_name), model_name)
        recommend_model = rec_cls(rargs, data)
        attack_model = ref.attack("Black", "NoneAttack")(aargs, data)
        arlib = ref.ARLib(recommend_model, attack_model, rargs, aargs)
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            arlib.RecommendTrain()
            arlib.RecommendTest()
            out["clean"] = list(arlib.rawRecommendresult)
            if save:
                path = "%s%s/%s_%d_%d_%s" % (rargs.save_dir, model_name, model_name, rargs.emb_size, rargs.n_layers, "ml-100k")
                assert os.path.isfile(path), "ARLib.py:127-131 torch.save(recommendModel) did not happen"
                loaded = torch.load(path, weights_only=False)                    # ARLib.py:112
                out["loaded"] = loaded.test()[1]
            arlib.PoisonDataAttack()
            for step in range(arlib.times):
                arlib.RecommendTrain(attack=step)
                arlib.RecommendTest(attack=step)
            arlib.ResultAnalysis()
        out["poison"] = list(arlib.attackRecommendresult)
        out["hitRate"], out["ndcg"] = arlib.avgHitRateAttack, arlib.avgNDCGAttack
        out["cls"] = type(arlib.recommendModel).__module__
        out["stdout"] = sink.getvalue()
    return out


@pytest.mark.parametrize("model_name", ["LightGCN", "XSimGCL"])
def test_arlib_driver_with_noneattack_runs_unmodified_on_the_dropin(model_name, tmp_path):
    with _env(ARLIB_B200_SAMPLER="host"):            # the reference's Python RNG stream: same triples on both sides
        ours = _run_driver(True, model_name, tmp_path / "ours", save=True)
    theirs = _run_driver(False, model_name, tmp_path / "ref")
    assert ours["cls"].startswith("arlib_b200.recommender.") and theirs["cls"] == "recommender." + model_name
    for key in ("clean", "poison"):
        got, want = ours[key], theirs[key]
        assert [m.split(":")[0] for m in got] == [m.split(":")[0] for m in want]          # 'Top 50\n', 'Hit Ratio:..\n', ...
        assert got[0] == "Top 50\n" and all(m.endswith("\n") for m in got)
        gv, wv = _values(got), _values(want)
        tol = 1e-3 if model_name == "LightGCN" else 2e-2     # XSimGCL: torch.rand_like noise differs between the two runs
        for k in wv:
            assert abs(gv[k] - wv[k]) < tol, (key, k, gv[k], wv[k])
    # torch.save / torch.load of the whole recommender object (ARLib.py:104-131)
    assert _values(ours["loaded"]) == _values(ours["clean"])
    assert "Recommender Test Result in Poisoning Environment on Average" in ours["stdout"]
    assert len(ours["hitRate"]) == 1 and np.isfinite(ours["hitRate"][0]) and np.isfinite(ours["ndcg"][0])


def _trained(ref, tmp_path, model_name="LightGCN", epochs=2):
    rargs = ref_loader.make_args(ref, dataset="ml-100k", model_name=model_name, maxEpoch=epochs)
    ref.tool.seedSet(rargs.seed)
    data = ref.DataLoader(rargs)
    rec = getattr(getattr(ref, model_name), model_name)(rargs, data)
    rec.train()
    return rargs, data, rec


def test_clear_bilevel_attack_runs_unmodified_on_the_dropin(tmp_path):
    """attack/White/CLeaR.py:56-159: fakeUserInject (re-__init__ on a live object, slice-assignment into the
    parameters), deepcopy per epoch, _init_uiAdj of the grown adjacency, differentiable model() + the attacker's Adam,
    recommender.train(Epoch, optimizer, evalNum) and AttackMetric through predict()."""
    with _workdir(tmp_path), ref_loader.reference_modules(cuda=CUDA, dropin=True, callers=True) as ref, \
            contextlib.redirect_stdout(io.StringIO()):
        rargs, data, rec = _trained(ref, tmp_path)
        aargs = ref_loader.make_attack_args(ref, attackCategory="White", attackModelName="CLeaR", Epoch=2, innerEpoch=1,
                                            outerEpoch=2, maliciousUserSize=0.01)
        attack = ref.attack("White", "CLeaR")(aargs, data)
        n_real, n_items = attack.userNum, attack.itemNum
        before = copy.deepcopy(rec).model.embedding_dict["item_emb"].detach().clone()
        poisoned = attack.posionDataAttack(copy.deepcopy(rec))
    assert type(rec).__module__ == "arlib_b200.recommender.LightGCN"
    assert poisoned.shape == (n_real + attack.fakeUserNum, n_items) and attack.fakeUserNum == int(n_real * 0.01)
    fake = poisoned[n_real:].toarray()
    assert (fake[:, attack.targetItem] == 1).all()                      # CLeaR.py:133-134
    assert ((fake > 0).sum(1) >= len(attack.targetItem)).all()
    assert torch.equal(before, rec.model.embedding_dict["item_emb"].detach())   # the caller's copy was attacked, not ours


def test_gta_proxy_subclass_runs_unmodified_on_the_dropin(tmp_path, monkeypatch):
    """attack/Black/GTA.py:57-236: ``class proxyLG(LightGCN)`` overrides train() with its own loop over the
    reference's sampler, model(), util.loss and evaluate()/save() of the base class."""
    with _workdir(tmp_path), ref_loader.reference_modules(cuda=CUDA, dropin=True, callers=True) as ref, \
            contextlib.redirect_stdout(io.StringIO()):
        rargs, data, rec = _trained(ref, tmp_path, epochs=1)
        GTA = ref.attack("Black", "GTA")
        import attack.Black.GTA as gta_mod
        assert issubclass(gta_mod.proxyLG, type(rec))
        # GTA.py:151 hard-codes train(Epoch=30) of the proxy inside fakeUserInject; 3 epochs exercise the same code
        real_train = gta_mod.proxyLG.train
        monkeypatch.setattr(gta_mod.proxyLG, "train",
                            lambda self, *a, **k: real_train(self, *a, **{**k, "Epoch": min(k.get("Epoch", 0) or 3, 3)}))
        aargs = ref_loader.make_attack_args(ref, attackCategory="Black", attackModelName="GTA", Epoch=1, innerEpoch=1,
                                            maliciousUserSize=0.01)
        attack = GTA(aargs, data)
        n_real = attack.userNum
        poisoned = attack.posionDataAttack(copy.deepcopy(rec))
    assert poisoned.shape[0] == n_real + attack.fakeUserNum
    assert np.isfinite(poisoned[n_real:].toarray()).all()


def test_pga_adjacency_gradient_attack_runs_on_the_dropin(tmp_path, monkeypatch):
    """attack/White/PGA.py:54-140: ``sparse_norm_adj.requires_grad = True`` + ``torch.autograd.grad(Loss,
    model.sparse_norm_adj)`` per 128-item batch, the D^-1/2 products on the returned sparse gradient and the fake-row
    update.  Harness shim: scipy >= 1.11 rejects a torch tensor as a fancy index (PGA.py:73), so
    ``torch.topk``'s index result is handed to scipy as a numpy array."""
    import scipy.sparse._index as sp_index
    real_validate = sp_index._validate_indices

    def validate(key, *a, **k):
        if isinstance(key, tuple):
            key = tuple(np.array(x.cpu().numpy()) if torch.is_tensor(x) else x for x in key)
        return real_validate(key, *a, **k)
    monkeypatch.setattr(sp_index, "_validate_indices", validate)
    with _workdir(tmp_path), ref_loader.reference_modules(cuda=CUDA, dropin=True, callers=True) as ref, \
            contextlib.redirect_stdout(io.StringIO()):
        rargs, data, rec = _trained(ref, tmp_path, epochs=1)
        aargs = ref_loader.make_attack_args(ref, attackCategory="White", attackModelName="PGA", Epoch=1, innerEpoch=1,
                                            outerEpoch=1, maliciousUserSize=0.01)
        attack = ref.attack("White", "PGA")(aargs, data)
        n_real = attack.userNum
        poisoned = attack.posionDataAttack(copy.deepcopy(rec))
    fake = poisoned[n_real:].toarray()
    assert poisoned.shape[0] == n_real + attack.fakeUserNum and np.isfinite(fake).all()
    assert (fake[:, attack.targetItem] == 1).all()                      # PGA.py:143-145
