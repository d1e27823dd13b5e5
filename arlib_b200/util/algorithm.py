"""Top-K helper with the reference's signature (util/algorithm.py:155-167).

``find_k_largest(K, candidates)`` is part of the surface attacks import
(attack/White/DLAttack.py:84, attack/Black/GTA.py:173).  It takes a HOST numpy
vector, so this drop-in keeps the reference's heap semantics on the host for that
call; the recommender's own ``test()`` never comes through here -- it runs the
fused device kernel (agcf_score_topk), which reproduces the same selection rule.
"""
import heapq


def find_k_largest(K, candidates):
    """min-heap of (score, id) over the first K entries, then replace the root
    whenever a later score is strictly larger; result sorted by score descending."""
    heap = [(s, k) for k, s in enumerate(candidates[:K])]
    heapq.heapify(heap)
    for k, s in enumerate(candidates[K:]):
        if s > heap[0][0]:
            heapq.heapreplace(heap, (s, k + K))
    heap.sort(key=lambda e: e[0], reverse=True)
    return [e[1] for e in heap], [e[0] for e in heap]
