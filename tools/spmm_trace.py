"""Per-CTA timeline of one agcf_spmm_csr_f32 launch (needs the -DAGCF_SPMM_TRACE build of the library:
ARLIB_B200_LIB=.../libagcf_trace.so).  Tuning aid: shows whether the long rows are the critical path."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse as sp
from bench import make_data
from arlib_b200 import ops, _lib
from arlib_b200.graph import DeviceGraph

name = sys.argv[1] if len(sys.argv) > 1 else "gowalla"
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
D = make_data(name, alpha)
U, I, E, d = D["U"], D["I"], D["E"], int(os.environ.get("TRACE_D", D["d"]))
N = U + I
dev = torch.device("cuda:0")
half = sp.csr_matrix((np.ones(E, dtype=np.float32), (D["tu"], D["ti"] + U)), shape=(N, N), dtype=np.float32)
g = DeviceGraph.from_dataloader_adj(half + half.T, dev)
X = torch.randn(N, d, device=dev); Y = torch.empty_like(X); A = torch.empty_like(X)
for _ in range(20):
    ops.spmm(g, X, Y=Y, acc_in=X, acc_out=A)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * (6 * 16384))()
lib.agcf_debug_spmm_trace.argtypes = [ctypes.c_void_p]
assert lib.agcf_debug_spmm_trace(buf) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 6).astype(np.int64)
lens = g.vrows[:, 1].cpu().numpy()
rpb = 8 * (32 // min(d // 4, 32))
n_cta = (g.n_vrows + rpb - 1) // rpb
t = t[:min(n_cta, 16384)]
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e3, (t[:, 3] - t0) / 1e3
meta, accd = (t[:, 1] - t[:, 0]) / 1e3, (t[:, 2] - t[:, 1]) / 1e3
epi = (t[:, 3] - t[:, 2]) / 1e3
print("%s a=%s d=%d: %d CTAs (%d work items, longest row %d segments), makespan %.1f us" % (name, alpha, d, n_cta, g.n_vrows, g.max_segments, end.max()))
dur = end - start
print("CTAs: dur mean %.1f max %.1f us" % (dur.mean(), dur.max()))
step = max(1, len(t) // 24)
for lo in range(0, len(t), step):
    hi = min(len(t), lo + step)
    print("CTA %5d-%5d  first item len %3d  start %6.1f-%6.1f  dur mean %5.1f max %5.1f  end max %6.1f  warp0: meta %4.2f gather %4.2f epilogue %4.2f us"
          % (lo, hi, int(lens[min(lo * rpb, len(lens) - 1)]), start[lo:hi].min(), start[lo:hi].max(), dur[lo:hi].mean(), dur[lo:hi].max(), end[lo:hi].max(),
             meta[lo:hi].mean(), accd[lo:hi].mean(), epi[lo:hi].mean()))
# concurrency over time
ev = np.concatenate([np.stack([start, np.ones_like(start)], 1), np.stack([end, -np.ones_like(end)], 1)])
ev = ev[np.argsort(ev[:, 0])]
conc = np.cumsum(ev[:, 1])
for q in (0.1, 0.25, 0.5, 0.75, 0.9, 0.97):
    k = np.searchsorted(ev[:, 0], q * end.max())
    print("t=%.0f%% (%.1f us): %d CTAs resident" % (q * 100, q * end.max(), conc[min(k, len(conc) - 1)]))
