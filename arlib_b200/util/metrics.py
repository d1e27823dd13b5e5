"""Ranking metrics -- mirror of the reference's util/metrics.py:4-114.

``ranking_evaluation(origin, res, N)`` keeps the reference's list-of-strings
output format (ARLib.py:172-191 regex-parses it).  The recommender's ``test()``
computes per-user hits / DCG / IDCG on device (agcf_rank_metrics) and formats
the same strings through ``format_measure``; this host version is kept for
callers that pass their own rec_list dict.
"""
import math

import numpy as np


class RecommendMetric(object):
    @staticmethod
    def hits(origin, res):
        return {u: len(set(origin[u].keys()) & set(p[0] for p in res[u])) for u in origin}

    @staticmethod
    def hit_ratio(origin, hits):
        total = sum(len(origin[u]) for u in origin)
        return sum(hits[u] for u in hits) / total

    @staticmethod
    def precision(hits, N):
        return sum(hits[u] for u in hits) / (len(hits) * N)

    @staticmethod
    def recall(hits, origin):
        vals = [hits[u] / len(origin[u]) for u in hits]
        return sum(vals) / len(vals)

    @staticmethod
    def F1(prec, recall):
        return 2 * prec * recall / (prec + recall) if (prec + recall) != 0 else 0

    @staticmethod
    def NDCG(origin, res, N):
        total = 0
        for u in res:
            dcg = 0
            idcg = 0
            for rank, item in enumerate(res[u]):
                if item[0] in origin[u]:
                    dcg += 1.0 / math.log(rank + 2)
            for rank in range(len(list(origin[u].keys())[:N])):
                idcg += 1.0 / math.log(rank + 2)
            total += dcg / idcg
        return total / len(res)


def format_measure(n, hit_ratio, precision, recall, ndcg):
    """the 5 strings the reference emits per cutoff (util/metrics.py:99-113)"""
    return ['Top ' + str(n) + '\n', 'Hit Ratio:' + str(hit_ratio) + '\n', 'Precision:' + str(precision) + '\n',
            'Recall:' + str(recall) + '\n', 'NDCG:' + str(ndcg) + '\n']


def ranking_evaluation(origin, res, N):
    measure = []
    for n in N:
        predicted = {u: res[u][:n] for u in res}
        if len(origin) != len(predicted):
            print('The Lengths of test set and predicted set do not match!')
            exit(-1)
        hits = RecommendMetric.hits(origin, predicted)
        measure += format_measure(n, RecommendMetric.hit_ratio(origin, hits), RecommendMetric.precision(hits, n),
                                  RecommendMetric.recall(hits, origin), RecommendMetric.NDCG(origin, predicted, n))
    return measure


class AttackMetric(object):
    """Mirror of the reference's AttackMetric (util/metrics.py:125-207): how often the attacker's target items
    reach the top-k lists of ALL users (train items are NOT masked there).

    The reference calls ``predict`` and a full ``np.argsort`` per user and per method (4 x U GEMVs and sorts,
    minutes at Gowalla shape -- SURVEY.md 8f-1).  Here one pass of the fused score + top-K kernel
    (agcf_score_topk without a mask, K = max(top)) over all users yields every list; the four statistics
    are then exact integer / float64 reductions of the [U, K] id matrix.  Lists are ordered by (score desc,
    item id asc); the reference's quicksort argsort leaves the order of exactly tied scores unspecified.
    """

    def __init__(self, recommendModel, targetItem, top=[10]):
        self.recommendModel = recommendModel
        self.targetItem = targetItem
        self.top = top
        self._lists = None

    def _topk_lists(self):
        import torch
        from .. import ops
        rec = self.recommendModel
        ue, ie = getattr(rec, "user_emb", None), getattr(rec, "item_emb", None)
        if not (torch.is_tensor(ue) and torch.is_tensor(ie) and ue.is_cuda):
            raise TypeError("AttackMetric needs a recommender exposing CUDA user_emb / item_emb tensors "
                            "(arlib_b200 has no CPU path)")
        ids = np.fromiter(rec.data.user.values(), dtype=np.int64, count=len(rec.data.user))
        K = min(int(max(self.top)), ie.shape[0])
        ue = ue.detach().float().contiguous()
        ie = ie.detach().float().contiguous()
        impl = 1 if ie.shape[1] <= 128 else 0
        out = np.empty((ids.shape[0], K), dtype=np.int64)
        rows = torch.from_numpy(ids.astype(np.int32)).to(ue.device)
        for lo in range(0, ids.shape[0], 16384):
            hi = min(ids.shape[0], lo + 16384)
            _, idx = ops.score_topk(ue, ie, K, user_rows=rows[lo:hi].contiguous(), impl=impl)
            out[lo:hi] = idx.cpu().numpy()
        return out

    def _member(self):
        """[U, K] bool: is the item at that rank one of the targets (computed once, reused by the 4 methods)"""
        if self._lists is None:
            lists = self._topk_lists()
            self._lists = np.isin(lists, np.asarray(list(self.targetItem), dtype=np.int64))
        return self._lists

    def precision(self):
        m = self._member()
        return [float(m[:, :k].sum()) / (m.shape[0] * k) for k in self.top]

    def hitRate(self):
        m = self._member()
        nt = len(self.targetItem)
        return [float(sum(int(bool(x)) / nt for x in m[:, :k].any(axis=1))) / m.shape[0] for k in self.top]

    def recall(self):
        m = self._member()
        return [float(m[:, :k].sum()) / (m.shape[0] * len(self.targetItem)) for k in self.top]

    def NDCG(self):
        m = self._member()
        res = []
        for k in self.top:
            w = 1.0 / np.log2(2.0 + np.arange(min(k, m.shape[1])))
            idcg = float(sum(1 / np.log2(2 + s) for s in range(k) if s < len(self.targetItem)))
            res.append(float((m[:, :k] * w[None, :]).sum()) / (m.shape[0] * idcg))
        return res
