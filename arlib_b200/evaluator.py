"""Full-rank evaluation on device: scoring + masked top-K + ranking metrics.

Reference: recommender/LightGCN.py:137-161 (test), :86-90 (predict),
util/algorithm.py:155-167, util/metrics.py:87-114.  One agcf_score_topk call (per
user chunk) replaces the per-user GEMV + D2H + Python mask loop + numba heap; one
agcf_rank_metrics call replaces the per-user Python metric loops.  Only the final
per-user numbers cross to the host, where the reference's own summation order and
string format are reproduced.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import ops
from .util.metrics import format_measure

# users per agcf_score_topk call.  ARLIB_B200_EVAL_CHUNK fixes it; by default it is what a ~4 GB workspace holds for the
# item table at hand (the stage-2 workspace grows with users x item groups: ~40 KB per user at 41 k items -> every named
# shape goes through ONE call; ~380 KB per user at 1 M items -> ~10 k users per call).  Measured on B200
# (profiles/r2_summary.md): Gowalla shape, 27 324 test users: one call 1.04 ms, two calls of <= 16 384 users 1.15 ms;
# Amazon-book shape, 51 551 users: one call 4.77 ms, two calls 4.97 ms.
_CHUNK_ENV = os.environ.get("ARLIB_B200_EVAL_CHUNK")
USER_CHUNK = int(_CHUNK_ENV) if _CHUNK_ENV else 32768      # (kept for callers that slice by a fixed chunk)
WS_BUDGET_BYTES = 4 << 30


def user_chunk(n_items, d, K):
    """Users per call for an [n_items, d] item table (see above)."""
    if _CHUNK_ENV:
        return int(_CHUNK_ENV)
    lib = ops._lib.load()
    a, b = int(lib.agcf_score_topk_ws_bytes(1024, n_items, d, K)), int(lib.agcf_score_topk_ws_bytes(2048, n_items, d, K))
    if a < 0 or b <= a:
        return 32768
    per_user = (b - a) / 1024.0
    return int(min(131072, max(4096, (WS_BUDGET_BYTES // per_user) // 1024 * 1024)))


# stage-1 implementation of agcf_score_topk: 1 = TF32 tcgen05 GEMM (d <= 128), 0 = fp32 CUDA-core GEMM
DEFAULT_IMPL = int(os.environ.get("ARLIB_B200_SCORE_IMPL", "1"))


def _csr_from_lists(lists, n_rows_hint=None):
    counts = np.fromiter((len(x) for x in lists), dtype=np.int64, count=len(lists))
    rowptr = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    flat = np.zeros(max(int(rowptr[-1]), 1), dtype=np.int32)
    for k, x in enumerate(lists):
        if len(x):
            flat[rowptr[k]:rowptr[k + 1]] = x
    return rowptr.astype(np.int32), flat


class FullRankEvaluator:
    """Device-side view of (data.test_set, data.training_set_u) for test()."""

    def __init__(self, data, device):
        self.device = device
        item = data.item
        self.users = list(data.test_set.keys())                       # reference iteration order
        uid = np.array([data.user[u] for u in self.users], dtype=np.int32)
        # mask CSR indexed by GLOBAL user id: sorted train item ids (data.user_rated, LightGCN.py:151-153)
        n_users = max(data.user_num, int(uid.max()) + 1 if len(uid) else 0)
        edges = getattr(data, 'pristine_edges', lambda: None)()
        if edges is not None:
            # untouched since the vectorised load: training_set_u is exactly these edges -- sorted unique item ids
            # of the TEST users' rows with array operations instead of a per-user dict walk
            eu, ei = edges
            is_test = np.zeros(n_users, dtype=bool)
            is_test[uid] = True
            keep = is_test[eu]
            key = np.sort(eu[keep].astype(np.int64) * data.item_num + ei[keep])
            if key.size:
                key = key[np.concatenate(([True], key[1:] != key[:-1]))]
            ku, ki = key // data.item_num, key % data.item_num
            mrp = np.zeros(n_users + 1, dtype=np.int64)
            np.cumsum(np.bincount(ku, minlength=n_users), out=mrp[1:])
            mrp, mit = mrp.astype(np.int32), (ki.astype(np.int32) if ki.size else np.zeros(1, np.int32))
        else:
            mask_lists = [()] * n_users
            for u in self.users:
                ids = np.fromiter((item[i] for i in data.training_set_u[u]), dtype=np.int32)
                ids.sort()
                mask_lists[data.user[u]] = ids
            mrp, mit = _csr_from_lists(mask_lists)
        # test CSR indexed by test-user POSITION: sorted ids of test items that exist in train
        t_lists, totals = [], []
        for u in self.users:
            names = data.test_set[u]
            ids = np.fromiter((item[i] for i in names if i in item), dtype=np.int32)
            ids.sort()
            t_lists.append(ids)
            totals.append(len(names))
        trp, tit = _csr_from_lists(t_lists)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.user_rows = to(uid)
        self.mask_rowptr, self.mask_items = to(mrp), to(mit)
        self.t_rowptr, self.t_items = to(trp), to(tit)
        self.test_total = to(np.array(totals, dtype=np.int32))
        self.test_total_host = totals
        self.id2item = np.array([data.id2item[k] for k in range(len(data.id2item))], dtype=object)
        self._ws = None

    @classmethod
    def from_arrays(cls, n_users, n_items, train_u, train_i, test_u, test_i, device):
        """Benchmark / test constructor from integer arrays (names = str(id))."""
        self = cls.__new__(cls)
        self.device = device

        def csr(rows, cols, n):
            order = np.lexsort((cols, rows))
            r, c = np.asarray(rows)[order], np.asarray(cols)[order]
            rp = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.bincount(r, minlength=n), out=rp[1:])
            return rp.astype(np.int32), (c.astype(np.int32) if c.size else np.zeros(1, np.int32))

        tusers = np.unique(test_u)
        pos = np.full(n_users, -1, dtype=np.int64)
        pos[tusers] = np.arange(tusers.shape[0])
        mrp, mit = csr(train_u, train_i, n_users)
        trp, tit = csr(pos[test_u], test_i, tusers.shape[0])
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.users = [str(int(u)) for u in tusers]
        self.user_rows = to(tusers.astype(np.int32))
        self.mask_rowptr, self.mask_items = to(mrp), to(mit)
        self.t_rowptr, self.t_items = to(trp), to(tit)
        totals = np.diff(trp).astype(np.int32)
        self.test_total = to(totals)
        self.test_total_host = totals.tolist()
        self.id2item = np.array([str(k) for k in range(n_items)], dtype=object)
        self._ws = None
        return self

    # ---------------------------------------------------------------- kernels
    def topk(self, user_emb, item_emb, K, impl=None):
        """(values, ids) [n_test_users, K] on device, reference selection rule."""
        impl = DEFAULT_IMPL if impl is None else impl
        if item_emb.shape[1] > 128:
            impl = 0                      # the tcgen05 path is compiled for d in {32, 64, 128}
        user_emb = user_emb.detach().contiguous()
        item_emb = item_emb.detach().contiguous()
        n = self.user_rows.numel()
        vals = torch.empty((n, K), dtype=torch.float32, device=self.device)
        idx = torch.empty((n, K), dtype=torch.int32, device=self.device)
        chunk = user_chunk(item_emb.shape[0], item_emb.shape[1], K)
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            need = ops._lib.load().agcf_score_topk_ws_bytes(hi - lo, item_emb.shape[0], item_emb.shape[1], K)
            if need < 0:
                ops._lib.check(int(need), "agcf_score_topk_ws_bytes")
            ws, keep = self._workspace(lo, hi, item_emb, int(need))
            ops.score_topk(user_emb, item_emb, K, user_rows=self.user_rows[lo:hi], mask_rowptr=self.mask_rowptr,
                           mask_items=self.mask_items, impl=impl, ws=ws, out=(vals[lo:hi], idx[lo:hi]), keep_mask_bits=keep)
        return vals, idx

    def _workspace(self, lo, hi, item_emb, need):
        """(workspace of the user chunk that starts at ``lo``, keep_mask_bits).  Every chunk owns its workspace, so the
        train-item mask bits stage 0 leaves there survive until the chunk's next evaluation: keep_mask_bits is True when
        the LAST call on that workspace scored exactly these users against an item table of this shape (the masks and
        user_rows of an evaluator never change) -- then the memset + bit scatter are skipped (AGCF_TOPK_KEEP_MASK_BITS)."""
        chunks = self.__dict__.setdefault("_ws_chunks", {})
        ws, old_key = chunks.get(lo, (None, None))
        if ws is None or ws.numel() < need:
            ws, old_key = torch.empty(need, dtype=torch.uint8, device=self.device), None
        key = (int(hi), int(item_emb.shape[0]), int(item_emb.shape[1]), int(ws.data_ptr()))
        chunks[lo] = (ws, key)
        self._ws = ws                                  # (bench.py times the scoring stage alone on it)
        return ws, key == old_key

    def topk_sharded(self, user_emb, item_emb, K, rank, world, impl=None):
        """Item-sharded scoring: this rank scores every test user against its
        item block with the fused top-K, the P per-shard lists are all-gathered (NCCL) and
        merged on device (agcf_topk_merge).  Identical result on every rank."""
        import torch.distributed as dist
        impl = DEFAULT_IMPL if impl is None else impl
        if item_emb.shape[1] > 128:
            impl = 0
        n_items = item_emb.shape[0]
        i0 = n_items * rank // world
        i1 = n_items * (rank + 1) // world
        block = item_emb[i0:i1].detach().contiguous()
        n = self.user_rows.numel()
        ue = user_emb.detach().contiguous()
        vals = torch.empty((n, K), dtype=torch.float32, device=self.device)
        idx = torch.empty((n, K), dtype=torch.int32, device=self.device)
        chunk = user_chunk(block.shape[0], block.shape[1], K)
        for lo in range(0, n, chunk):                        # the stage-2 workspace grows with the users of a call
            hi = min(n, lo + chunk)
            ops.score_topk(ue, block, K, user_rows=self.user_rows[lo:hi], mask_rowptr=self.mask_rowptr,
                           mask_items=self.mask_items, item_offset=i0, impl=impl, out=(vals[lo:hi], idx[lo:hi]))
        all_v = torch.empty((world, n, K), dtype=torch.float32, device=self.device)
        all_i = torch.empty((world, n, K), dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(all_v, vals)
        dist.all_gather_into_tensor(all_i, idx)
        return ops.topk_merge(all_v, all_i)

    def topk_user_sharded(self, user_emb, item_emb, K, rank, world, impl=None, gather=True):
        """User-sharded evaluation (multi-GPU default when the item table fits one GPU, which it does at every named
        shape): rank r scores test users [r n/P, (r+1) n/P) against ALL items with the fused top-K -- every per-user
        cost (mask bits, GEMM rows, candidate selection, exact re-scoring, selection) divides by P and nothing is
        exchanged until the P blocks of [n/P, K] results are all-gathered (NCCL, ``gather=False`` keeps them local:
        the ranking metrics are per-user sums).  Item sharding (topk_sharded) replicates the per-user work on every
        rank and only pays when the item table itself must be split."""
        import torch.distributed as dist
        impl = DEFAULT_IMPL if impl is None else impl
        if item_emb.shape[1] > 128:
            impl = 0
        n = self.user_rows.numel()
        per = (n + world - 1) // world
        lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
        ue, ie = user_emb.detach().contiguous(), item_emb.detach().contiguous()
        vals = torch.full((per, K), float("-inf"), dtype=torch.float32, device=self.device)
        idx = torch.full((per, K), -1, dtype=torch.int32, device=self.device)
        chunk = user_chunk(ie.shape[0], ie.shape[1], K)
        for a in range(lo, hi, chunk):
            b = min(hi, a + chunk)
            need = ops._lib.load().agcf_score_topk_ws_bytes(b - a, ie.shape[0], ie.shape[1], K)
            if need < 0:
                ops._lib.check(int(need), "agcf_score_topk_ws_bytes")
            ws, keep = self._workspace(a, b, ie, int(need))
            ops.score_topk(ue, ie, K, user_rows=self.user_rows[a:b], mask_rowptr=self.mask_rowptr,
                           mask_items=self.mask_items, impl=impl, ws=ws, out=(vals[a - lo:b - lo], idx[a - lo:b - lo]),
                           keep_mask_bits=keep)
        if not gather:
            return vals[:hi - lo], idx[:hi - lo], (lo, hi)
        all_v = torch.empty((world * per, K), dtype=torch.float32, device=self.device)
        all_i = torch.empty((world * per, K), dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(all_v, vals)
        dist.all_gather_into_tensor(all_i, idx)
        return all_v[:n], all_i[:n]

    def per_user_metrics(self, idx, cutoffs):
        """[n_users, n_cutoffs, 3] float64 (hits, dcg, idcg) on device."""
        K = idx.shape[1]
        key = (K, tuple(int(c) for c in cutoffs))
        consts = self.__dict__.setdefault("_metric_consts", {})
        if key not in consts:
            # built once per (K, cutoffs): torch.tensor(list, device=cuda) is a synchronous pageable copy, and one per
            # call drained the launch queue in front of every evaluation
            inv_log = torch.tensor([1.0 / math.log(r + 2) for r in range(max(K, max(cutoffs)))], dtype=torch.float64,
                                   device=self.device)
            consts[key] = (torch.tensor(list(key[1]), dtype=torch.int32, device=self.device), inv_log)
        cut, inv_log = consts[key]
        return ops.rank_metrics(idx, self.t_rowptr, self.t_items, self.test_total, cut, inv_log)

    # ------------------------------------------------------------- reference API
    def measure(self, idx, cutoffs):
        """The reference's list of metric strings (util/metrics.py:87-114), summed on the
        host in the reference's order from the per-user device results."""
        per = self.per_user_metrics(idx, cutoffs).cpu().numpy()
        totals = self.test_total_host
        total_num = sum(totals)
        out = []
        n_users = len(totals)
        for c, n in enumerate(cutoffs):
            hits = per[:, c, 0].astype(np.int64).tolist()
            hit_ratio = sum(hits) / total_num
            precision = sum(hits) / (n_users * n)
            rec = [h / t for h, t in zip(hits, totals)]
            recall = sum(rec) / len(rec)
            ratio = (per[:, c, 1] / per[:, c, 2]).tolist()
            total = 0
            for r in ratio:
                total += r
            out += format_measure(n, hit_ratio, precision, recall, total / n_users)
        return out

    def rec_list(self, vals, idx):
        """{user: [(item_name, score), ...]} like recommender/LightGCN.py:155-156"""
        names = self.id2item[idx.cpu().numpy().astype(np.int64)]
        scores = vals.cpu().numpy()
        return {u: list(zip(names[k].tolist(), scores[k].tolist())) for k, u in enumerate(self.users)}

    def test(self, user_emb, item_emb, top_n, max_n, impl=None):
        vals, idx = self.topk(user_emb, item_emb, max_n, impl)
        return self.rec_list(vals, idx), self.measure(idx, top_n)
