import numpy as np


def assert_topk_matches(sel_idx, scores, K, tol):
    """``sel_idx`` (K ids) is a valid top-K of ``scores`` up to score perturbations of
    size tol: every id strictly above the boundary band is selected, nothing below it is."""
    sel = set(int(i) for i in sel_idx)
    assert len(sel) == min(K, scores.shape[0]), "duplicate or missing ids"
    kth = np.sort(scores)[-K]
    must = set(np.flatnonzero(scores > kth + tol).tolist())
    may = set(np.flatnonzero(scores >= kth - tol).tolist())
    assert must <= sel, "missed items clearly inside the top-K"
    assert sel <= may, "selected items clearly outside the top-K"


def boundary_gap(scores, K):
    s = np.sort(scores)
    return float(s[-K] - s[-K - 1])
