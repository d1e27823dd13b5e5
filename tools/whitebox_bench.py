"""BASELINE.json configs[3]: the white-box poisoning inner loop (PGA / CLeaR skeleton, attack/White/PGA.py:86-140,
CLeaR.py:56-159) on a synthetic ml-1M-shaped graph with 1 % fake users, through the drop-in LightGCN class:

  1. model._init_uiAdj(ui_adj)           -- re-normalise + re-upload the (U+F+I)^2 adjacency      (PGA.py:93-97)
  2. train(Epoch=1, optimizer=...)       -- retrain with the attacker's optimizer                  (CLeaR.py:145-146)
  3. adjacency gradient                  -- torch.autograd.grad(loss, model.sparse_norm_adj)       (PGA.py:98-117)
  4. gradient to the embedding rows      -- CW-style loss on model() outputs, backward             (CLeaR.py:89-129)
  5. masked_score_topk                   -- dense scores + mask + topk of the attack               (CLeaR.py:75-81)
  6. AttackMetric(...).hitRate()         -- target-item hit ratio over all users                   (CLeaR.py:148-149)

Prints one JSON line with the wall time of each stage (CUDA-synchronised)."""
import contextlib, io, json, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch

from arlib_b200.recommender.LightGCN import LightGCN
from arlib_b200.util.DataLoader import DataLoader
from arlib_b200.util.algorithm import masked_score_topk
from arlib_b200.util.loss import bpr_loss
from arlib_b200.util.metrics import AttackMetric
from arlib_b200.util.synth import SHAPES, synth_edges

name = sys.argv[1] if len(sys.argv) > 1 else "ml-1m"
U, I, E = SHAPES[name]
F = max(1, U // 100)
tu, ti, su, si = synth_edges(U, I, E, 0.5, 0.5, 0)
rng = np.random.default_rng(1)
# fake users appended after the real ones, each with 1 % of the items (a poisoned train.txt re-loaded)
fu = np.repeat(np.arange(U, U + F), max(1, I // 100))
fi = np.concatenate([rng.choice(I, max(1, I // 100), replace=False) for _ in range(F)])
data = DataLoader.from_arrays(np.concatenate([tu, fu]), np.concatenate([ti, fi]), su, si, name=name)
args = types.SimpleNamespace(topK="50", emb_size=64, n_layers=2, batch_size=2048, lRate=0.005, reg=1e-4, maxEpoch=1,
                             seed=2018, sampler="device", model_name="LightGCN")
quiet = lambda: contextlib.redirect_stdout(io.StringIO())


def timed(fn, reps=3):
    fn()                                                     # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


with quiet():
    rec = LightGCN(args, data)
model = rec.model.cuda()
N = data.user_num + data.item_num
R = data.interaction_mat.tocsr()
ui_adj = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
out = {"workload": name, "users": data.user_num, "fake_users": F, "items": data.item_num, "edges": int(R.nnz)}

out["init_uiAdj_s"], _ = timed(lambda: model._init_uiAdj(ui_adj))
opt = torch.optim.Adam(model.parameters(), lr=args.lRate)
with quiet():
    t, _ = timed(lambda: rec.train(Epoch=1, optimizer=opt, evalNum=5), reps=1)
out["retrain_epoch_s"] = t
out["retrain_steps"] = (len(data.training_data) + 2047) // 2048
targets = torch.tensor(rng.choice(I, 5, replace=False), device="cuda")
users = torch.arange(data.user_num - F, device="cuda")


def cw_loss():
    pu, pi = model()
    s = pu[users] @ pi[targets].T
    return -torch.log(torch.sigmoid(s) + 1e-8).mean()


def adj_grad():
    model.sparse_norm_adj.requires_grad = True
    g = torch.autograd.grad(cw_loss(), model.sparse_norm_adj)[0]
    model.sparse_norm_adj.requires_grad = False
    return g


out["adjacency_grad_s"], g = timed(adj_grad)
out["adjacency_grad_nnz"] = int(g._nnz())


def emb_grad():
    opt.zero_grad()
    cw_loss().backward()
    return model.embedding_dict["user_emb"].grad


out["embedding_grad_s"], _ = timed(emb_grad)
with torch.no_grad():
    pu, pi = model()
out["masked_score_topk_s"], (vals, idx) = timed(lambda: masked_score_topk(pu, pi, 50, R))
rec.user_emb, rec.item_emb = pu, pi
names = [data.id2item[int(t)] for t in targets.cpu()]
out["attack_metric_hitrate_s"], hr = timed(lambda: AttackMetric(rec, names, [50]).hitRate())
out["hit_rate"] = [float(x) for x in hr]
print(json.dumps(out), flush=True)
