"""Times the evaluation pipeline (agcf_score_topk per chunk + metrics) on the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_data, xavier_tables
from arlib_b200.evaluator import FullRankEvaluator
from arlib_b200 import ops
D = make_data(sys.argv[1] if len(sys.argv) > 1 else "gowalla", 0.5)
U, I, d = D["U"], D["I"], D["d"]
dev = torch.device("cuda:0")
ue, ie = xavier_tables(U, I, d)
ue, ie = ue.to(dev), ie.to(dev)
ev = FullRankEvaluator.from_arrays(U, I, D["tu"], D["ti"], D["su"], D["si"], dev)
for impl in (1, 0):
    for _ in range(5): ev.topk(ue, ie, 50, impl=impl)
    torch.cuda.synchronize()
    reps = 20 if impl == 1 else 5
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    t0 = time.perf_counter(); evs[0].record()
    for k in range(reps):
        v, i = ev.topk(ue, ie, 50, impl=impl)
        evs[k + 1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    per = sorted(evs[k].elapsed_time(evs[k + 1]) for k in range(reps))
    print("impl %d ctas/sm %s: %.2f ms per eval (device, mean of %d; median %.2f, min %.2f, max %.2f), %.2f ms wall, %d users"
          % (impl, os.environ.get("AGCF_STAGE2_CTAS_PER_SM", "8"), evs[0].elapsed_time(evs[-1]) / reps, reps, per[reps // 2], per[0], per[-1],
             wall, ev.user_rows.numel()))
# heavy-tailed item norms (what training produces): does the TF32 margin admit too many groups?
torch.manual_seed(0)
scale = 1.0 + 20.0 * torch.rand(I, 1, device=dev) ** 8
ie2 = ie * scale
rows = ev.user_rows[:4096].contiguous()
for impl in (1, 0):
    v, i, f = ops.score_topk(ue, ie2, 50, user_rows=rows, mask_rowptr=ev.mask_rowptr, mask_items=ev.mask_items, impl=impl, return_flags=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): ops.score_topk(ue, ie2, 50, user_rows=rows, mask_rowptr=ev.mask_rowptr, mask_items=ev.mask_items, impl=impl)
    b.record(); torch.cuda.synchronize()
    print("heavy-tailed norms impl %d: mean candidate groups %.1f (max %d), %.2f ms / 4096 users" % (impl, f.float().mean().item(), f.max().item(), a.elapsed_time(b) / 3))
