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
