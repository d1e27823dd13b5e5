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


def masked_score_topk(user_emb, item_emb, K, interactions=None, impl=None):
    """Dense score + mask + top-k of the white-/black-box attacks in ONE fused device call.

    The attacks build a dense CPU (U+F) x I score matrix in 2048-row chunks, overwrite the entries of the
    interaction matrix with -10e8 and call ``torch.topk`` (attack/White/CLeaR.py:75-81, BiLevelAttackBatch.py:76-92,
    attack/Black/GTA.py:182-196, attack/Gray/FedRecAttack.py:78-91).  Here the same result comes from
    agcf_score_topk (tcgen05 scoring, exact fp32 rescoring, scores never written to memory):

        values, indices = masked_score_topk(Pu, Pi, topk, uiAdj2)      # == torch.topk(masked Pu @ Pi.T, topk)

    ``interactions``: scipy sparse [n_users, n_items] whose NON-ZERO entries are masked (``.nonzero()`` semantics:
    explicitly stored zeros are not), or None.  Returns device tensors (fp32 [n, K], int64 [n, K]) sorted by score
    descending; equal scores are ordered by item id (torch.topk leaves that order unspecified)."""
    import numpy as np
    import scipy.sparse as sp
    import torch
    from .. import ops
    from ..evaluator import DEFAULT_IMPL
    user_emb = user_emb.detach().to(torch.float32).contiguous()
    item_emb = item_emb.detach().to(torch.float32).contiguous()
    if not item_emb.is_cuda:
        raise TypeError("masked_score_topk runs on the device the embeddings live on (CUDA); got %s" % item_emb.device)
    n_u = user_emb.shape[0]
    rp = it = None
    if interactions is not None:
        csr = sp.csr_matrix(interactions, copy=True)
        if csr.shape[0] != n_u or csr.shape[1] != item_emb.shape[0]:
            raise ValueError("interactions must be [%d, %d]" % (n_u, item_emb.shape[0]))
        csr.eliminate_zeros()
        csr.sort_indices()
        rp = torch.from_numpy(csr.indptr.astype(np.int32)).to(item_emb.device)
        it = torch.from_numpy(csr.indices.astype(np.int32) if csr.nnz else np.zeros(1, np.int32)).to(item_emb.device)
    impl = DEFAULT_IMPL if impl is None else impl
    if item_emb.shape[1] > 128:
        impl = 0
    vals = torch.empty((n_u, K), dtype=torch.float32, device=item_emb.device)
    idx = torch.empty((n_u, K), dtype=torch.int32, device=item_emb.device)
    chunk = 16384
    for lo in range(0, n_u, chunk):
        hi = min(n_u, lo + chunk)
        rows = torch.arange(lo, hi, dtype=torch.int32, device=item_emb.device)
        v, i = ops.score_topk(user_emb, item_emb, K, user_rows=rows, mask_rowptr=rp, mask_items=it, impl=impl)
        vals[lo:hi], idx[lo:hi] = v, i
    return vals, idx.long()
