"""The closed form of find_k_largest's index set (incl. its tie rule) used to check
the device top-K -- verified against the heap itself (util/algorithm.py:155-167)."""
import numpy as np

from oracle import port


def test_tie_rule_closed_form_matches_heap():
    rng = np.random.default_rng(0)
    for trial in range(3000):
        n = int(rng.integers(1, 60))
        K = int(rng.integers(1, 20))
        levels = int(rng.integers(1, 6))
        s = rng.integers(0, levels, n).astype(np.float32)
        if n < K:
            continue
        ids, vals = port.find_k_largest_py(K, s)
        assert set(ids) == port.topk_reference_set(K, s), (K, s)
        assert list(vals) == sorted(vals, reverse=True)


def test_numba_and_python_heaps_agree_on_sets():
    rng = np.random.default_rng(1)
    for trial in range(50):
        s = rng.integers(0, 4, 200).astype(np.float32)
        a, _ = port.find_k_largest(50, s)
        b, _ = port.find_k_largest_py(50, s)
        assert set(a) == set(b)


def test_no_ties_equals_stable_argsort():
    rng = np.random.default_rng(2)
    s = rng.standard_normal(1000).astype(np.float32)
    ids, _ = port.find_k_largest_py(50, s)
    assert ids == np.argsort(-s, kind="stable")[:50].tolist()


def test_attack_metric_port_reproduces_reference_golden():
    """oracle.port.attack_metric vs outputs of the unmodified reference AttackMetric frozen by
    oracle/make_golden_attack_metric.py."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attack_metric.npz"))
    scores = z["user_emb"] @ z["item_emb"].T
    users = {str(n): int(i) for n, i in zip(z["user_names"], z["user_ids"])}
    for tag in ("a", "b"):
        got = port.attack_metric(lambda name: scores[users[name]], list(users), z["targets"].tolist(), z["top_" + tag].tolist())
        for name in ("precision", "hitRate", "recall", "NDCG"):
            assert np.array_equal(np.array(got[name]), z["%s_%s" % (name, tag)]), name
