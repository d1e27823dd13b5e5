"""Micro-benchmark of agcf_spmm_csr_f32 on the Gowalla-shaped graph (tuning aid).
usage: AGCF_SPMM_VARIANT=k python tools/spmm_variants.py [workload] [alpha]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse as sp
from bench import make_data
from arlib_b200 import ops
from arlib_b200.graph import DeviceGraph

name = sys.argv[1] if len(sys.argv) > 1 else "gowalla"
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
D = make_data(name, alpha)
U, I, E, d = D["U"], D["I"], D["E"], int(os.environ.get("SPMM_D", D["d"]))
N = U + I
dev = torch.device("cuda:0")
half = sp.csr_matrix((np.ones(E, dtype=np.float32), (D["tu"], D["ti"] + U)), shape=(N, N), dtype=np.float32)
g = DeviceGraph.from_dataloader_adj(half + half.T, dev)
if os.environ.get("SPMM_HOTCOL"):
    # diagnostic: every gather reads row (k mod HOTCOL) -- the whole gather stream hits L1 / a few L2 lines, what is left
    # is the kernel's non-gather work (descriptors, index stream, shuffles, combine, epilogue)
    g.col.copy_(torch.arange(g.nnz, device=dev, dtype=torch.int32) % int(os.environ["SPMM_HOTCOL"]))
X = torch.randn(N, d, device=dev)
Y = torch.empty_like(X); A = torch.empty_like(X)
# a batch-like node mask: 2048 edges -> users, pos items; 2048 uniform neg items
rng = np.random.default_rng(0)
pick = rng.integers(0, E, 2048)
nodes = np.unique(np.concatenate([D["tu"][pick], U + D["ti"][pick], U + rng.integers(0, I, 2048)]))
words = np.zeros((N + 31) // 32, dtype=np.uint32)
np.bitwise_or.at(words, nodes >> 5, (np.uint32(1) << (nodes & 31).astype(np.uint32)))
mask = torch.from_numpy(words.view(np.int32)).to(dev)
Xs = torch.zeros_like(X); Xs[torch.from_numpy(nodes).to(dev)] = X[torch.from_numpy(nodes).to(dev)]

def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

part = os.environ.get("SPMM_PART")            # "world": time each rank's row partition on this one GPU (no peer stores)
if part:
    w = int(part)
    b = g.row_ranges(w)
    for r in range(w):
        gp = g.partition(b[r], b[r + 1])
        tp = timeit(lambda: ops.spmm(gp, X, Y=Y, acc_in=X, acc_out=A))
        print("partition %d/%d rows [%d,%d) nnz %d work items %d max segments %d: %.1f us" % (r, w, b[r], b[r + 1], gp.local_nnz, gp.n_vrows, gp.max_segments, tp))
    sys.exit(0)
t_full = timeit(lambda: ops.spmm(g, X, Y=Y, acc_in=X, acc_out=A))
t_row = timeit(lambda: ops.spmm(g, X, acc_in=A, acc_out=A, row_mask=mask))
t_col = timeit(lambda: ops.spmm(g, Xs, Y=Y, addend=Xs, col_mask=mask))
live = np.diff(g.rowptr.cpu().numpy())[nodes].sum() / g.nnz
print("d=%d variant %s %s a=%s: full %.1f us  row-masked %.1f us  col-masked %.1f us  (live nnz %.3f, %d work items)"
      % (d, os.environ.get("AGCF_SPMM_VARIANT", "0"), name, alpha, t_full, t_row, t_col, live, g.n_vrows))
