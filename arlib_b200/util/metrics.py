"""Ranking metrics -- mirror of the reference's util/metrics.py:4-114.

``ranking_evaluation(origin, res, N)`` keeps the reference's list-of-strings
output format (ARLib.py:172-191 regex-parses it).  The recommender's ``test()``
computes per-user hits / DCG / IDCG on device (agcf_rank_metrics) and formats
the same strings through ``format_measure``; this host version is kept for
callers that pass their own rec_list dict.
"""
import math


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
