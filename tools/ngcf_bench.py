"""NGCF at the Gowalla shape (BASELINE configs[1] graph, NGCF encoder of recommender/NGCF.py:197-212, L = 3, d = 64): the
fused NGCFEngine (one propagation per layer + the fused dense kernels of csrc/ngcf.cu, CUDA-graph replay) against the
reference-shaped loop (torch autograd over the agcf SpMM, torch.mm for the d x d products, torch.optim.Adam) on the same
graph / parameters.  One JSON line each."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch

from bench import make_data
from arlib_b200 import ops
from arlib_b200.engine import DeviceTrainSet, NGCFEngine
from arlib_b200.graph import DeviceGraph

name = sys.argv[1] if len(sys.argv) > 1 else "gowalla"
D = make_data(name, 0.5)
U, I, E, L, d, B = D["U"], D["I"], D["E"], D["L"], 64, D["B"]
N = U + I
dev = torch.device("cuda:0")
half = sp.csr_matrix((np.ones(E, dtype=np.float32), (D["tu"], D["ti"] + U)), shape=(N, N), dtype=np.float32)
g = DeviceGraph.from_dataloader_adj(half + half.T, dev)
ts = DeviceTrainSet.from_arrays(D["tu"], D["ti"], U, I, dev)
gen = torch.Generator().manual_seed(2018)
a = (6.0 / (N + d)) ** 0.5
table0 = ((torch.rand(N, d, generator=gen) * 2 - 1) * a).to(dev)
W0 = ((torch.rand(L, 2 * d, d, generator=gen) * 2 - 1) * (6.0 / (2 * d)) ** 0.5).to(dev)

# ---- fused engine
table, W = table0.clone(), W0.clone()
eng = NGCFEngine(g, table, W, U, 0.005, 1e-4, B, E)
eng.sample_epoch(ts, 2018, 0)
K = min(400, E // B)
eng.run_steps(0, 3, use_graph=False)
eng.run_steps(0, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.run_steps(0, K); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
# per-kernel: the dense forward / backward on their own
P, X, Y = torch.randn(N, d, device=dev), torch.randn(N, d, device=dev), torch.empty(N, d, device=dev)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
fwd_us = t(lambda: ops.ngcf_dense_forward(P, X, W[0], Y))
dP, dE = torch.empty_like(P), torch.empty_like(P)
part, dW = torch.empty((296, 2 * d * d), device=dev), torch.empty((2 * d, d), device=dev)
WT = W[0].t().contiguous()
bwd_us = t(lambda: ops.ngcf_dense_backward(P, Y, P, X, WT, dP, dE, part, dW))
spmm_us = t(lambda: ops.spmm(g, X, Y=Y))
flop_f = 2.0 * N * 2 * d * d
print(json.dumps({"workload": name, "engine": "ngcf-fused", "L": L, "d": d, "steps_timed": K, "ms_per_step": ms,
                  "triples_per_s": B / (ms * 1e-3), "launches_per_step": eng.launches_per_step, "last_loss": float(eng.out4[K - 1, 0]),
                  "dense_forward_us": fwd_us, "dense_forward_tflops_fp32": flop_f / (fwd_us * 1e-6) / 1e12,
                  "dense_backward_us": bwd_us, "dense_backward_tflops_fp32": 2 * flop_f / (bwd_us * 1e-6) / 1e12,
                  "spmm_us": spmm_us}), flush=True)

# ---- reference-shaped loop on autograd (what train() runs with a caller's optimizer)
import types
from arlib_b200.encoder import spmm_autograd
from arlib_b200.util.loss import bpr_l2_fused
import torch.nn.functional as F
pu = torch.nn.Parameter(table0[:U].clone()); pi = torch.nn.Parameter(table0[U:].clone())
Ws = [torch.nn.Parameter(W0[k, :d].clone()) for k in range(L)] + [torch.nn.Parameter(W0[k, d:].clone()) for k in range(L)]
opt = torch.optim.Adam([pu, pi] + Ws, lr=0.005)
u_all, i_all, j_all = eng.tu[:E].long(), eng.ti[:E].long(), eng.tj[:E].long()
def step(b):
    ego = torch.cat([pu, pi], 0)
    layers = [ego]
    for k in range(L):
        tt = torch.mm(ego, Ws[k])
        ego = F.leaky_relu(spmm_autograd(g, tt) + tt + torch.mm(spmm_autograd(g, ego) * ego, Ws[L + k]))
        layers.append(ego)
    out = torch.mean(torch.stack(layers, dim=1), dim=1)
    sl = slice(b * B, (b + 1) * B)
    loss = bpr_l2_fused(out[:U], out[U:], u_all[sl], i_all[sl], j_all[sl], 1e-4)
    opt.zero_grad(); loss.backward(); opt.step()
for b in range(5): step(b)
torch.cuda.synchronize(); t0 = time.perf_counter()
n_ref = 50
for b in range(5, 5 + n_ref): step(b)
torch.cuda.synchronize()
ms_ref = (time.perf_counter() - t0) / n_ref * 1e3
print(json.dumps({"workload": name, "engine": "ngcf-reference-shaped-loop", "steps_timed": n_ref, "ms_per_step": ms_ref,
                  "triples_per_s": B / (ms_ref * 1e-3)}), flush=True)
