"""Per-op device time of one multi-GPU training step (eager launches, CUDA events around every op and barrier).
torchrun --nproc-per-node N tools/dist_breakdown.py [rows|dshard]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
from bench import make_data, xavier_tables, LR, REG

mode = sys.argv[1] if len(sys.argv) > 1 else "rows"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from arlib_b200 import ops
import arlib_b200.engine as engmod
from arlib_b200.dist import DistContext
from arlib_b200.engine import DeviceTrainSet, LightGCNEngine
from arlib_b200.graph import DeviceGraph

D = make_data("gowalla", 0.5)
U, I, E, L, d, B = D["U"], D["I"], D["E"], D["L"], D["d"], D["B"]
N = U + I
half = sp.csr_matrix((np.ones(E, dtype=np.float32), (D["tu"], D["ti"] + U)), shape=(N, N), dtype=np.float32)
g = DeviceGraph.from_dataloader_adj(half + half.T, dev)
ue, ie = xavier_tables(U, I, d)
comm = DistContext(dev)
eng = LightGCNEngine(g, torch.cat([ue, ie]).to(dev), U, L, LR, REG, B, E, comm=comm, mode=mode)
eng.sample_epoch(DeviceTrainSet.from_arrays(D["tu"], D["ti"], U, I, dev), 2018, 0)
for k in range(5):
    eng.run_steps(k, 1, use_graph=False)
torch.cuda.synchronize(); dist.barrier()
rec = collections.defaultdict(list)
order = []

def wrap(name, fn):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        rec[name].append((e0, e1))
        return r
    return f

for name in ("spmm", "bpr_forward", "bpr_partial", "bpr_finish", "bpr_backward", "zero_rows", "adam_step", "increment"):
    setattr(engmod.ops, name, wrap(name, getattr(ops, name)))
comm.barrier = wrap("barrier", comm.barrier)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
t0.record()
for k in range(n):
    eng.run_steps(5 + k, 1, use_graph=False)
t1.record()
torch.cuda.synchronize()
if rank == 0:
    print("mode %s world %d: eager step %.1f us (launch-bound upper bound)" % (eng.mode, world, t0.elapsed_time(t1) / n * 1e3))
    tot = 0.0
    for name, evs in rec.items():
        ms = [a.elapsed_time(b) * 1e3 for a, b in evs]
        per_step = sum(ms) / n
        tot += per_step
        print("  %-13s %3d/step  mean %6.1f us  per step %6.1f us" % (name, len(ms) // n, np.mean(ms), per_step))
    print("  sum of device times per step: %.1f us" % tot)
    sp_ms = [a.elapsed_time(b) * 1e3 for a, b in rec["spmm"]]
    print("  spmm by position in the step:", " ".join("%.1f" % np.mean(sp_ms[i::2 * L]) for i in range(2 * L)))
dist.barrier(); dist.destroy_process_group()
