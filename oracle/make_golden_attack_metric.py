"""Freeze golden vectors for AttackMetric from the UNMODIFIED reference class (util/metrics.py:125-207).
Run in the builder container (needs /root/reference):  python oracle/make_golden_attack_metric.py
Writes tests/golden/attack_metric.npz (inputs + the four metric lists for two cutoff sets)."""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("ARLIB_REFERENCE", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "util", "metrics.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from oracle import port
    rng = np.random.default_rng(2018)
    U, I, d = 60, 500, 32
    ue = rng.normal(size=(U, d)).astype(np.float32)
    ie = rng.normal(size=(I, d)).astype(np.float32)
    targets = np.array([7, 123, 499, 250], dtype=np.int64)
    ie[targets] *= 2.0
    users = {"u%d" % k: int(k) for k in rng.permutation(U)}
    scores = ue @ ie.T
    model = types.SimpleNamespace(data=types.SimpleNamespace(user=users), predict=lambda name: scores[users[name]])
    out = {"user_emb": ue, "item_emb": ie, "targets": targets, "user_names": np.array(list(users)), "user_ids": np.array(list(users.values()))}
    for tag, top in (("a", [10]), ("b", [5, 20, 50])):
        am = ref.AttackMetric(model, targets.tolist(), top)
        mine = port.attack_metric(model.predict, list(users), targets.tolist(), top)
        out["top_" + tag] = np.array(top)
        for name in ("precision", "hitRate", "recall", "NDCG"):
            r = np.array(getattr(am, name)(), dtype=np.float64)
            assert np.array_equal(r, np.array(mine[name])), (name, r, mine[name])      # the port reproduces it bit for bit
            out["%s_%s" % (name, tag)] = r
    path = os.path.join(ROOT, "tests", "golden", "attack_metric.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
