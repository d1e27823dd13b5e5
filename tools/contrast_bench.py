"""BASELINE.json configs[2]: SimGCL / XSimGCL (noise-perturbed propagation + InfoNCE) on a synthetic
Yelp2018-shaped graph, through the drop-in recommender classes (reference training loop, torch.optim.Adam,
agcf SpMM forward/backward with the noise fused, fused InfoNCE kernels, on-device Philox sampler).
Prints one JSON line per model: ms per training step and triples/s over N timed batches, + full-rank eval."""
import json, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

from arlib_b200.util.DataLoader import DataLoader
from arlib_b200.util.synth import SHAPES, synth_edges

name = sys.argv[1] if len(sys.argv) > 1 else "yelp2018"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
models = [m for m in sys.argv[3].split(",") if m != "none"] if len(sys.argv) > 3 else ["XSimGCL", "SimGCL", "NGCF", "LightGCN"]
U, I, E = SHAPES[name]
tu, ti, su, si = synth_edges(U, I, E, 0.5, 0.5, 0)
t0 = time.perf_counter()
data = DataLoader.from_arrays(tu, ti, su, si, name=name) if models else None
t_load = time.perf_counter() - t0
dev = torch.device("cuda:0")

for model in models:
    import importlib
    cls = getattr(importlib.import_module("arlib_b200.recommender." + model), model)
    args = types.SimpleNamespace(topK="50", emb_size=64, n_layers=2, batch_size=2048, lRate=0.005, reg=1e-4, maxEpoch=1,
                                 seed=2018, sampler="device", model_name=model, fused=False)   # the reference-shaped loop
    torch.manual_seed(2018)
    rec = cls(args, data)
    m = rec.model.cuda()
    opt = torch.optim.Adam(m.parameters(), lr=args.lRate)
    # time N steps of the class's own epoch loop by cutting the batch generator short
    orig = rec._epoch_batches
    times = {}

    def limited(dev_, n=n_steps + 5):
        for k, b in enumerate(orig(dev_)):
            if k == 5:
                torch.cuda.synchronize(); times["t0"] = time.perf_counter()
            if k >= n:
                break
            yield b
        torch.cuda.synchronize(); times["t1"] = time.perf_counter()

    rec._epoch_batches = limited
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        rec.train(Epoch=1, optimizer=opt, evalNum=1)
    step_ms = (times["t1"] - times["t0"]) / n_steps * 1e3
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        rec_list, measure = rec.test()
    torch.cuda.synchronize(); t_test = time.perf_counter() - t0
    print(json.dumps({"workload": name, "model": model, "U": U, "I": I, "E": E, "batch": 2048, "steps_timed": n_steps,
                      "ms_per_step": step_ms, "triples_per_s": 2048 / (step_ms * 1e-3),
                      "test_s": t_test, "test_users": len(data.test_set), "dataloader_s": t_load,
                      "measure": [x.strip() for x in measure]}), flush=True)

# ---- the fused engines on the same graph (what train() runs when the recommender owns the optimizer)
from arlib_b200.engine import ContrastiveEngine, DeviceTrainSet, LightGCNEngine
from arlib_b200.graph import DeviceGraph
import scipy.sparse as sp

N = U + I
half = sp.csr_matrix((np.ones(E, dtype=np.float32), (tu, ti + U)), shape=(N, N), dtype=np.float32)
g = DeviceGraph.from_dataloader_adj(half + half.T, dev)
ts = DeviceTrainSet.from_arrays(tu, ti, U, I, dev)
gen = torch.Generator().manual_seed(2018)
a = (6.0 / (N + 64)) ** 0.5
for kind in ("xsimgcl", "simgcl", "lightgcn"):
    table = ((torch.rand(N, 64, generator=gen) * 2 - 1) * a).to(dev)
    if kind == "lightgcn":
        eng = LightGCNEngine(g, table, U, 2, 0.005, 1e-4, 2048, E)
    else:
        eng = ContrastiveEngine(g, table, U, kind, 2, 0.1, 0.2, 0.1 if kind == "xsimgcl" else 0.2, 0.005, 1e-4, 2048, E)
    eng.sample_epoch(ts, 2018, 0)
    K = min(int(os.environ.get('CONTRAST_STEPS', '500')), E // 2048)
    eng.run_steps(0, 3, use_graph=False)
    eng.run_steps(0, K)                                    # capture + warm replay
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run_steps(0, K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    extra = {}
    if kind != "lightgcn":
        rec, cl = eng.losses(0, K)
        extra = {"rec_loss_last": float(rec[-1]), "cl_loss_last": float(cl[-1])}
    print(json.dumps(dict({"workload": name, "engine": kind, "steps_timed": K, "ms_per_step": ms,
                           "triples_per_s": 2048 / (ms * 1e-3), "launches_per_step": eng.launches_per_step}, **extra)),
          flush=True)
