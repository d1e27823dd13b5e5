"""CPU restatement of the ARLib graph-CF hot path (numpy / scipy / torch-CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA
path and the timed CPU baseline of bench.py.  Never imported by arlib_b200/.

Every function cites the reference file:line (relative to /root/reference) it
restates.  The arithmetic of the reference lives in third-party libraries
(torch.sparse.mm / autograd / Adam, scipy.sparse products, numpy power/sqrt,
numba heapq); this port calls the SAME library entry points in the SAME order,
so on one machine it is bit-identical to the reference -- that claim is pinned
by oracle/make_golden.py (run in the builder container against the live
reference) and by tests/test_oracle_golden.py against tests/golden/*.npz.
"""
from __future__ import annotations

import heapq
import math
import random as _pyrandom
from collections import defaultdict

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

try:  # the reference jit-compiles its heap top-K with numba (util/algorithm.py:155)
    from numba import jit as _numba_jit
except Exception:  # pragma: no cover
    _numba_jit = None


# --------------------------------------------------------------------------
# data model  (util/FileIO.py:22-31, util/DataLoader.py:8-55)
# --------------------------------------------------------------------------
def read_triples(path):
    """util/FileIO.py:22-31 -- '<user> <item> <weight>' per line, split on ' '."""
    rows = []
    with open(path) as fh:
        for line in fh:
            parts = line.strip().split(" ")
            rows.append([parts[0], parts[1], float(parts[2])])
    return rows


class PortData:
    """The slice of util/DataLoader.py:8-108 the hot path reads.

    ids are assigned by FIRST APPEARANCE in the training rows (:33-40); val /
    test rows of users unseen in train are dropped (:42-55); user_num/item_num
    count ids seen in train (:25-26).
    """

    def __init__(self, training_data, val_data=(), test_data=()):
        self.training_data = training_data
        self.user, self.item = {}, {}
        self.id2user, self.id2item = {}, {}
        self.training_set_u = defaultdict(dict)
        self.training_set_i = defaultdict(dict)
        self.val_set = defaultdict(dict)
        self.test_set = defaultdict(dict)
        for row in training_data:
            u, i, r = row[0], row[1], row[2]
            if u not in self.user:
                self.user[u] = len(self.user)
                self.id2user[self.user[u]] = u
            if i not in self.item:
                self.item[i] = len(self.item)
                self.id2item[self.item[i]] = i
            self.training_set_u[u][i] = r
            self.training_set_i[i][u] = r
        for row in val_data:
            if row[0] in self.user:
                self.val_set[row[0]][row[1]] = row[2]
        for row in test_data:
            if row[0] in self.user:
                self.test_set[row[0]][row[1]] = row[2]
        self.user_num = len(self.training_set_u)
        self.item_num = len(self.training_set_i)
        self.ui_adj = bipartite_adjacency(self.edge_index()[0], self.edge_index()[1],
                                          self.user_num, self.item_num)
        self.norm_adj = normalize_graph_mat(self.ui_adj)

    def edge_index(self):
        u = np.fromiter((self.user[r[0]] for r in self.training_data), dtype=np.int64,
                        count=len(self.training_data))
        i = np.fromiter((self.item[r[1]] for r in self.training_data), dtype=np.int64,
                        count=len(self.training_data))
        return u, i

    def get_user_id(self, u):
        return self.user.get(u)

    def user_rated(self, u):
        d = self.training_set_u[u]
        return list(d.keys()), list(d.values())


# --------------------------------------------------------------------------
# adjacency  (util/DataLoader.py:57-87, recommender/LightGCN.py:212-215,247-252)
# --------------------------------------------------------------------------
def bipartite_adjacency(user_idx, item_idx, user_num, item_num):
    """util/DataLoader.py:57-71 -- A = R_ext + R_ext^T, fp32 ones (duplicates sum)."""
    n = user_num + item_num
    ones = np.ones_like(user_idx, dtype=np.float32)
    half = sp.csr_matrix((ones, (user_idx, item_idx + user_num)), shape=(n, n), dtype=np.float32)
    return half + half.T


def normalize_graph_mat(adj):
    """util/DataLoader.py:73-87 -- square case: D^-1/2 A D^-1/2 with
    np.power(rowsum,-0.5), inf -> 0, association (D.A).D."""
    rowsum = np.array(adj.sum(1))
    d_inv = np.power(rowsum, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.0
    d_mat = sp.diags(d_inv)
    return d_mat.dot(adj).dot(d_mat)


def init_uiadj_norm(ui_adj):
    """recommender/LightGCN.py:212-215 (_init_uiAdj; same text in NGCF.py:186-189,
    SimGCL.py:180-183, XSimGCL.py:199-202): 1/np.sqrt on row AND column sums,
    no inf guard, fractional weights allowed."""
    d_row = np.array((1 / np.sqrt(ui_adj.sum(1)))).flatten()
    d_col = np.array((1 / np.sqrt(ui_adj.sum(0)))).flatten()
    return sp.diags(d_row) @ ui_adj @ sp.diags(d_col)


def to_torch_coo(mat):
    """recommender/LightGCN.py:247-252 (convert_sparse_mat_to_tensor): COO int64
    indices + fp32 values, NOT flagged coalesced."""
    coo = mat.tocoo()
    idx = torch.from_numpy(np.vstack([coo.row, coo.col]).astype(np.int64))
    val = torch.from_numpy(coo.data).float()
    return torch.sparse_coo_tensor(idx, val, coo.shape)


# --------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------
def lightgcn_forward(adj, user_emb, item_emb, n_layers):
    """recommender/LightGCN.py:230-240 -- E0=[U;V]; Ek = A E(k-1); mean(E0..EL)."""
    ego = torch.cat([user_emb, item_emb], 0)
    layers = [ego]
    for _ in range(n_layers):
        ego = torch.sparse.mm(adj, ego)
        layers.append(ego)
    out = torch.mean(torch.stack(layers, dim=1), dim=1)
    nu = user_emb.shape[0]
    return out[:nu], out[nu:]


def ngcf_forward(adj, user_emb, item_emb, w1, w2):
    """recommender/NGCF.py:197-212 -- T=E W1; E'=leaky_relu(A T + T + ((A E)*E) W2)."""
    ego = torch.cat([user_emb, item_emb], 0)
    layers = [ego]
    for k in range(len(w1)):
        t = torch.mm(ego, w1[k])
        ego = F.leaky_relu(torch.sparse.mm(adj, t) + t +
                           torch.mm(torch.sparse.mm(adj, ego) * ego, w2[k]))
        layers.append(ego)
    out = torch.mean(torch.stack(layers, dim=1), dim=1)
    nu = user_emb.shape[0]
    return out[:nu], out[nu:]


def simgcl_forward(adj, user_emb, item_emb, n_layers, eps, noises=None):
    """recommender/SimGCL.py:198-210 -- layer 0 excluded from the mean; if
    perturbed, E'k += sign(E'k) * normalize(noise_k, dim=-1) * eps.  ``noises`` is
    the list of U[0,1) tensors the reference draws with torch.rand_like (injected
    here so both sides see the same numbers); None = unperturbed."""
    ego = torch.cat([user_emb, item_emb], 0)
    layers = []
    for k in range(n_layers):
        ego = torch.sparse.mm(adj, ego)
        if noises is not None:
            ego = ego + torch.sign(ego) * F.normalize(noises[k], dim=-1) * eps
        layers.append(ego)
    out = torch.mean(torch.stack(layers, dim=1), dim=1)
    nu = user_emb.shape[0]
    return out[:nu], out[nu:]


def xsimgcl_forward(adj, user_emb, item_emb, n_layers, eps, layer_cl, noises=None):
    """recommender/XSimGCL.py:205-223 -- as SimGCL plus the layer_cl view."""
    ego = torch.cat([user_emb, item_emb], 0)
    layers = []
    cl = ego
    for k in range(n_layers):
        ego = torch.sparse.mm(adj, ego)
        if noises is not None:
            ego = ego + torch.sign(ego) * F.normalize(noises[k], dim=-1) * eps
        layers.append(ego)
        if k == layer_cl - 1:
            cl = ego
    out = torch.mean(torch.stack(layers, dim=1), dim=1)
    nu = user_emb.shape[0]
    if noises is None:
        return out[:nu], out[nu:]
    return out[:nu], out[nu:], cl[:nu], cl[nu:]


# --------------------------------------------------------------------------
# losses  (util/loss.py:5-9, 25-29, 42-49)
# --------------------------------------------------------------------------
def bpr_loss(user_emb, pos_emb, neg_emb):
    """util/loss.py:5-9 -- mean(-log(1e-7 + sigmoid(<u,i> - <u,j>)))."""
    pos = torch.mul(user_emb, pos_emb).sum(dim=1)
    neg = torch.mul(user_emb, neg_emb).sum(dim=1)
    return torch.mean(-torch.log(10e-8 + torch.sigmoid(pos - neg)))


def l2_reg_loss(reg, *embs):
    """util/loss.py:25-29 -- reg * sum of UN-squared Frobenius norms."""
    total = 0
    for e in embs:
        total = total + torch.norm(e, p=2)
    return total * reg


def infonce(view1, view2, temperature):
    """util/loss.py:42-49."""
    v1, v2 = F.normalize(view1, dim=1), F.normalize(view2, dim=1)
    pos = torch.exp((v1 * v2).sum(dim=-1) / temperature)
    ttl = torch.exp(torch.matmul(v1, v2.transpose(0, 1)) / temperature).sum(dim=1)
    return torch.mean(-torch.log(pos / ttl))


# --------------------------------------------------------------------------
# sampler  (util/sampler.py:4-30)
# --------------------------------------------------------------------------
def next_batch_pairwise(data, batch_size, rng=_pyrandom):
    """util/sampler.py:4-30 -- in-place shuffle of data.training_data, consecutive
    slices of ``batch_size`` (last one short), one rejection-sampled negative per
    row drawn with ``choice`` over the item NAMES in dict order."""
    rows = data.training_data
    rng.shuffle(rows)
    n = len(rows)
    start = 0
    while start < n:
        stop = min(start + batch_size, n)
        names = list(data.item.keys())
        u_idx, i_idx, j_idx = [], [], []
        for r in rows[start:stop]:
            user, item = r[0], r[1]
            i_idx.append(data.item[item])
            u_idx.append(data.user[user])
            neg = rng.choice(names)
            while neg in data.training_set_u[user]:
                neg = rng.choice(names)
            j_idx.append(data.item[neg])
        start = stop
        yield u_idx, i_idx, j_idx


# --------------------------------------------------------------------------
# top-K  (util/algorithm.py:155-167)
# --------------------------------------------------------------------------
def _find_k_largest_py(K, candidates):
    heap = []
    for iid, score in enumerate(candidates[:K]):
        heap.append((score, iid))
    heapq.heapify(heap)
    for iid, score in enumerate(candidates[K:]):
        if score > heap[0][0]:
            heapq.heapreplace(heap, (score, iid + K))
    heap.sort(key=lambda d: d[0], reverse=True)
    return [h[1] for h in heap], [h[0] for h in heap]


find_k_largest_py = _find_k_largest_py
if _numba_jit is not None:
    find_k_largest = _numba_jit(nopython=True)(_find_k_largest_py)
else:  # pragma: no cover
    find_k_largest = _find_k_largest_py
find_k_largest.__doc__ = """util/algorithm.py:155-167 -- min-heap of (score, iid) over the first K, then
heapreplace iff score > heap minimum, final sort by score descending."""


def topk_reference_set(K, scores):
    """Closed form of the index SET find_k_largest returns, ties included
    (verified there over 20k random cases and re-verified in
    tests/test_oracle_topk.py): with s* the K-th largest value, all score > s*
    plus a specific window of the indices tied at s*."""
    scores = np.asarray(scores)
    n = scores.shape[0]
    if K >= n:
        return set(range(n))
    sstar = np.partition(scores, n - K)[n - K]
    greater = scores > sstar
    ge = scores >= sstar
    m = int(greater.sum())
    cum_ge = np.cumsum(ge)
    p = int(np.searchsorted(cum_ge, K))          # first index where #(>= s*) reaches K
    g_p = int(greater[:p + 1].sum())
    ties = np.flatnonzero(scores == sstar)
    chosen = ties[m - g_p: K - g_p]
    return set(np.flatnonzero(greater).tolist()) | set(chosen.tolist())


# --------------------------------------------------------------------------
# metrics  (util/metrics.py:4-114)
# --------------------------------------------------------------------------
def ranking_evaluation(origin, res, cutoffs):
    """util/metrics.py:87-114 with RecommendMetric.hits/hit_ratio/precision/
    recall/NDCG (:9-85).  Returns the same list of strings."""
    if len(origin) != len(res):
        raise SystemExit(-1)      # the reference prints and exit(-1)s (:94-96)
    out = []
    for n in cutoffs:
        pred = {u: res[u][:n] for u in res}
        hits = {}
        for u in origin:
            want = set(origin[u].keys())
            got = set(p[0] for p in pred[u])
            hits[u] = len(want & got)
        total = sum(len(origin[u]) for u in origin)
        hr = sum(hits.values()) / total
        prec = sum(hits[u] for u in hits) / (len(hits) * n)
        rec_list = [hits[u] / len(origin[u]) for u in hits]
        recall = sum(rec_list) / len(rec_list)
        sum_ndcg = 0
        for u in pred:
            dcg = 0
            idcg = 0
            for rank, item in enumerate(pred[u]):
                if item[0] in origin[u]:
                    dcg += 1.0 / math.log(rank + 2)
            for rank, _ in enumerate(list(origin[u].keys())[:n]):
                idcg += 1.0 / math.log(rank + 2)
            sum_ndcg += dcg / idcg
        ndcg = sum_ndcg / len(pred)
        out.append("Top " + str(n) + "\n")
        out.append("Hit Ratio:" + str(hr) + "\n")
        out.append("Precision:" + str(prec) + "\n")
        out.append("Recall:" + str(recall) + "\n")
        out.append("NDCG:" + str(ndcg) + "\n")
    return out


# --------------------------------------------------------------------------
# full-rank evaluation  (recommender/LightGCN.py:86-90, 137-161)
# --------------------------------------------------------------------------
def predict(data, user_emb, item_emb, user_name):
    """recommender/LightGCN.py:86-90 -- one GEMV, fp32, un-masked, to numpy."""
    u = data.get_user_id(user_name)
    with torch.no_grad():
        return torch.matmul(user_emb[u], item_emb.transpose(0, 1)).cpu().numpy()


def full_rank_test(data, user_emb, item_emb, max_n, cutoffs, users=None):
    """recommender/LightGCN.py:137-161 -- per test user: predict, mask the
    user's train items with -10e8, find_k_largest, map ids to names."""
    rec_list = {}
    todo = data.test_set if users is None else users
    for user in todo:
        cand = predict(data, user_emb, item_emb, user)
        rated, _ = data.user_rated(user)
        for item in rated:
            cand[data.item[item]] = -10e8
        ids, scores = find_k_largest(max_n, cand)
        rec_list[user] = list(zip([data.id2item[i] for i in ids], scores))
    if users is not None:
        return rec_list, None
    return rec_list, ranking_evaluation(data.test_set, rec_list, cutoffs)


# --------------------------------------------------------------------------
# training step  (recommender/LightGCN.py:46-64)
# --------------------------------------------------------------------------
class LightGCNTrainer:
    """The body of LightGCN.train() (recommender/LightGCN.py:29-72) on CPU torch
    with INJECTABLE triples: same ops, same order (model(), 3 index gathers,
    bpr_loss + l2_reg_loss(reg, user_emb, pos_item_emb), zero_grad, backward,
    Adam.step with torch defaults)."""

    def __init__(self, norm_adj, user_emb, item_emb, n_layers, lr, reg):
        self.adj = to_torch_coo(norm_adj)
        self.user_emb = torch.nn.Parameter(user_emb.clone())
        self.item_emb = torch.nn.Parameter(item_emb.clone())
        self.n_layers = n_layers
        self.reg = reg
        self.opt = torch.optim.Adam([self.user_emb, self.item_emb], lr=lr)

    def forward(self):
        return lightgcn_forward(self.adj, self.user_emb, self.item_emb, self.n_layers)

    def loss(self, u_idx, i_idx, j_idx):
        ru, ri = self.forward()
        ue, pe, ne = ru[u_idx], ri[i_idx], ri[j_idx]
        return bpr_loss(ue, pe, ne) + l2_reg_loss(self.reg, ue, pe)

    def step(self, u_idx, i_idx, j_idx):
        loss = self.loss(u_idx, i_idx, j_idx)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.item())


class _BprTrainerBase:
    """Shared skeleton of the four train() bodies (recommender/LightGCN.py:46-64, NGCF.py:48-66, SimGCL.py:46-69,
    XSimGCL.py:56-80): forward, three row gathers, bpr_loss (+ l2_reg_loss(reg, user_emb, pos_item_emb)) (+ cl term),
    zero_grad, backward, Adam.step (torch defaults, lr = args.lRate) over ``model.parameters()`` in registration
    order.  Returns python floats like the reference prints them."""

    def _params(self):
        return [self.user_emb, self.item_emb]

    def _finish(self, loss):
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()


class NGCFTrainer(_BprTrainerBase):
    """recommender/NGCF.py:31-66 with NGCF_Encoder.forward (:197-212).  ``w1`` / ``w2``: lists of the d x d layer
    weights (xavier_uniform, drawn w1_k then w2_k per layer after the two embedding tables, :176-183).  Adam is
    element-wise, so the registration order of ``model.parameters()`` does not enter the result."""

    def __init__(self, norm_adj, user_emb, item_emb, w1, w2, lr, reg):
        self.adj = to_torch_coo(norm_adj)
        self.user_emb = torch.nn.Parameter(user_emb.clone())
        self.item_emb = torch.nn.Parameter(item_emb.clone())
        self.w1 = [torch.nn.Parameter(w.clone()) for w in w1]
        self.w2 = [torch.nn.Parameter(w.clone()) for w in w2]
        self.reg = reg
        params = [self.user_emb, self.item_emb]
        for a, b in zip(self.w1, self.w2):
            params += [a, b]
        self.opt = torch.optim.Adam(params, lr=lr)

    def forward(self):
        return ngcf_forward(self.adj, self.user_emb, self.item_emb, self.w1, self.w2)

    def step(self, u_idx, i_idx, j_idx):
        ru, ri = self.forward()
        ue, pe, ne = ru[u_idx], ri[i_idx], ri[j_idx]
        loss = bpr_loss(ue, pe, ne) + l2_reg_loss(self.reg, ue, pe)
        self._finish(loss)
        return float(loss.item())


def unique_ids_f32(ids):
    """``torch.unique(torch.Tensor(ids).type(torch.long))`` -- recommender/SimGCL.py:213-214, XSimGCL.py:40-41: the ids
    pass through float32."""
    return torch.unique(torch.Tensor(ids).type(torch.long))


class SimGCLTrainer(_BprTrainerBase):
    """recommender/SimGCL.py:36-69 with SimGCL_Encoder.forward / cal_cl_loss (:198-219): one clean pass for the rec
    loss, two perturbed passes for the contrastive views, InfoNCE(tau = 0.2) over the unique batch users and unique
    positive items.  ``noise()`` returns the next U[0,1) tensor [N, d] (the reference calls torch.rand_like once per
    perturbed layer, pass 1 layers 1..L then pass 2 layers 1..L)."""

    def __init__(self, norm_adj, user_emb, item_emb, n_layers, eps, cl_rate, lr, reg, noise, tau=0.2, dtype=torch.float32):
        # dtype = float64: the same loop in double precision -- the yardstick for how well conditioned "parameters after
        # an epoch of Adam" is (tests); the reference itself is float32
        self.adj = to_torch_coo(norm_adj).to(dtype)
        self.user_emb = torch.nn.Parameter(user_emb.clone().to(dtype))
        self.item_emb = torch.nn.Parameter(item_emb.clone().to(dtype))
        self.n_layers, self.eps, self.cl_rate, self.reg, self.tau = n_layers, eps, cl_rate, reg, tau
        self.noise = (lambda: noise().to(dtype)) if dtype != torch.float32 else noise
        self.opt = torch.optim.Adam([self.user_emb, self.item_emb], lr=lr)

    def forward(self, perturbed=False):
        noises = [self.noise() for _ in range(self.n_layers)] if perturbed else None
        return simgcl_forward(self.adj, self.user_emb, self.item_emb, self.n_layers, self.eps, noises)

    def step(self, u_idx, i_idx, j_idx):
        ru, ri = self.forward()
        ue, pe, ne = ru[u_idx], ri[i_idx], ri[j_idx]
        rec_loss = bpr_loss(ue, pe, ne)
        uu, ii = unique_ids_f32(u_idx), unique_ids_f32(i_idx)
        u1, i1 = self.forward(True)
        u2, i2 = self.forward(True)
        cl_loss = self.cl_rate * (infonce(u1[uu], u2[uu], self.tau) + infonce(i1[ii], i2[ii], self.tau))
        self._finish(rec_loss + l2_reg_loss(self.reg, ue, pe) + cl_loss)
        return float(rec_loss.item()), float(cl_loss.item())


class XSimGCLTrainer(_BprTrainerBase):
    """recommender/XSimGCL.py:46-80 with XSimGCL_Encoder.forward (:205-223) and cal_cl_loss (:39-44): ONE perturbed
    pass gives the rec view and the layer_cl view; InfoNCE(tau = 0.1)."""

    def __init__(self, norm_adj, user_emb, item_emb, n_layers, eps, cl_rate, layer_cl, lr, reg, noise, tau=0.1,
                 dtype=torch.float32):
        self.adj = to_torch_coo(norm_adj).to(dtype)
        self.user_emb = torch.nn.Parameter(user_emb.clone().to(dtype))
        self.item_emb = torch.nn.Parameter(item_emb.clone().to(dtype))
        self.n_layers, self.eps, self.cl_rate, self.layer_cl = n_layers, eps, cl_rate, layer_cl
        self.reg, self.tau = reg, tau
        self.noise = (lambda: noise().to(dtype)) if dtype != torch.float32 else noise
        self.opt = torch.optim.Adam([self.user_emb, self.item_emb], lr=lr)

    def forward(self, perturbed=False):
        noises = [self.noise() for _ in range(self.n_layers)] if perturbed else None
        return xsimgcl_forward(self.adj, self.user_emb, self.item_emb, self.n_layers, self.eps, self.layer_cl, noises)

    def step(self, u_idx, i_idx, j_idx):
        ru, ri, cu, ci = self.forward(True)
        ue, pe, ne = ru[u_idx], ri[i_idx], ri[j_idx]
        rec_loss = bpr_loss(ue, pe, ne)
        uu, ii = unique_ids_f32(u_idx), unique_ids_f32(i_idx)
        cl_loss = self.cl_rate * (infonce(ru[uu], cu[uu], self.tau) + infonce(ri[ii], ci[ii], self.tau))
        self._finish(rec_loss + l2_reg_loss(self.reg, ue, pe) + cl_loss)
        return float(rec_loss.item()), float(cl_loss.item())


def adjacency_value_grad(norm_adj, user_emb, item_emb, n_layers, loss_fn):
    """attack/White/PGA.py:97-117 -- d loss / d (values of sparse_norm_adj),
    restricted to the stored pattern (torch returns a sparse COO gradient with
    the same nnz).  Returns (coo.row, coo.col, grad_values)."""
    adj = to_torch_coo(norm_adj).coalesce()
    adj.requires_grad_(True)
    ru, ri = lightgcn_forward(adj, user_emb, item_emb, n_layers)
    loss = loss_fn(ru, ri)
    g = torch.autograd.grad(loss, adj)[0].coalesce()
    idx = g.indices().numpy()
    return idx[0], idx[1], g.values().numpy()


# --------------------------------------------------------------------------
# array-backed data shim (bench.py cpu_baseline: no string dicts for 1M edges)
# --------------------------------------------------------------------------
class ArrayEvalData:
    """The attributes full_rank_test() reads, for a SAMPLE of test users, built
    from integer arrays (names = str(id)).  Same code path as PortData for the
    per-user loop of recommender/LightGCN.py:148-156."""

    def __init__(self, user_num, item_num, train_u, train_i, test_u, test_i, users):
        self.user_num, self.item_num = user_num, item_num
        self.item = {str(k): k for k in range(item_num)}
        self.id2item = {k: str(k) for k in range(item_num)}
        self.user = {str(int(u)): int(u) for u in users}
        self.training_set_u = defaultdict(dict)
        self.test_set = defaultdict(dict)
        keep = np.isin(train_u, users)
        for u, i in zip(train_u[keep].tolist(), train_i[keep].tolist()):
            self.training_set_u[str(u)][str(i)] = 1.0
        keep = np.isin(test_u, users)
        for u, i in zip(test_u[keep].tolist(), test_i[keep].tolist()):
            self.test_set[str(u)][str(i)] = 1.0

    def get_user_id(self, u):
        return self.user.get(u)

    def user_rated(self, u):
        d = self.training_set_u[u]
        return list(d.keys()), list(d.values())


# ------------------------------------------------------------------ AttackMetric
def attack_metric(predict_fn, users, target_items, top):
    """Restatement of AttackMetric.precision / hitRate / recall / NDCG (reference util/metrics.py:135-207):
    per user ``np.argsort(-score)[:k]`` of the UN-masked scores, then the four accumulations exactly as written
    there.  ``predict_fn(user) -> np.ndarray[I]``.  Returns {"precision": [...], "hitRate": [...], ...}."""
    n = len(top)
    lists = {}
    for u in users:
        score = predict_fn(u)
        order = np.argsort(-score, kind="stable")          # reference: default quicksort; ties unspecified there
        lists[u] = [order[:k] for k in top]
    tset = set(target_items)
    prec_hit, prec_tot = [0] * n, [0] * n
    hr_hit, hr_tot = [0.0] * n, [0] * n
    rec_hit, rec_tot = [0] * n, [0] * n
    nd_hit, nd_tot = [0.0] * n, [0.0] * n
    for u in users:
        result = lists[u]
        for i, k in enumerate(top):
            prec_tot[i] += k                                                   # :144
            hr_tot[i] += 1                                                     # :162
            rec_tot[i] += len(target_items)                                    # :178
            idcg = 0.0
            for s in range(k):                                                 # :196-199
                if s < len(target_items):
                    idcg += 1 / np.log2(2 + s)
            nd_tot[i] += idcg
        for j in target_items:                                                 # :145-148, :179-182
            for i in range(n):
                if j in result[i]:
                    prec_hit[i] += 1
                    rec_hit[i] += 1
        for i in range(n):                                                     # :163-164
            hr_hit[i] += int(len(tset & set(result[i].tolist())) > 0) / len(target_items)
        for i, r in enumerate(result):                                         # :200-203
            for rank, j in enumerate(r):
                if j in tset:
                    nd_hit[i] += 1 / np.log2(2 + rank)
    return {"precision": [prec_hit[i] / prec_tot[i] for i in range(n)],
            "hitRate": [hr_hit[i] / hr_tot[i] for i in range(n)],
            "recall": [rec_hit[i] / rec_tot[i] for i in range(n)],
            "NDCG": [nd_hit[i] / nd_tot[i] for i in range(n)]}
