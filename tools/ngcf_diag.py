import copy, random, sys, types, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from oracle import port
import arlib_b200.util.sampler as sampler_mod
from arlib_b200.recommender.NGCF import NGCF
from arlib_b200.util.DataLoader import DataLoader
g = np.load("tests/golden/ml100k_lightgcn.npz")
names_u = [str(x) for x in g["user_names"]]; names_i = [str(x) for x in g["item_names"]]
train = [[names_u[u], names_i[i], 1.0] for u, i in zip(g["train_u"], g["train_i"])]
test = [[str(u), str(i), 1.0] for u, i in zip(g["test_user_names"], g["test_item_names"])]
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())
real = sampler_mod.next_batch_pairwise
for nrows in (20000, 44212):
  for deep in (False, True):
    for epochs in (1, 2):
        data = DataLoader.from_rows([list(r) for r in train[:nrows]], (), test)
        args = types.SimpleNamespace(topK="50", emb_size=64, n_layers=2, batch_size=2048, lRate=0.005, reg=1e-4, maxEpoch=epochs,
                                     seed=2018, sampler="host", model_name="NGCF", fused=False)
        random.seed(3); torch.manual_seed(3)
        rec = NGCF(args, data)
        if deep:
            rec = copy.deepcopy(rec)
        init = {n: p.detach().cpu().clone() for n, p in rec.model.named_parameters()}
        norm_adj = rec.data.norm_adj.copy()
        batches = []
        def recording(d, bs):
            for bt in real(d, bs):
                batches.append(tuple(list(x) for x in bt)); yield bt
        sampler_mod.next_batch_pairwise = recording
        import io, contextlib
        random.seed(5)
        with contextlib.redirect_stdout(io.StringIO()):
            rec.train(evalNum=1)
        sampler_mod.next_batch_pairwise = real
        tr = port.NGCFTrainer(norm_adj, init["embedding_dict.user_emb"], init["embedding_dict.item_emb"],
                              [init["W.w1_%d" % k] for k in range(2)], [init["W.w2_%d" % k] for k in range(2)], 0.005, 1e-4)
        errs = []
        for k, bt in enumerate(batches):
            tr.step(*bt)
        print("rows %d deepcopy %s epochs %d steps %d: user %.2e item %.2e w1_0 %.2e" % (
            nrows, deep, epochs, len(batches), rel(rec.model.embedding_dict["user_emb"].detach(), tr.user_emb.detach()),
            rel(rec.model.embedding_dict["item_emb"].detach(), tr.item_emb.detach()), rel(rec.model.W["w1_0"].detach(), tr.w1[0].detach())), flush=True)
