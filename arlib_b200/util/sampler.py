"""Pairwise batch sampler with the reference's signature (util/sampler.py:4-30).

This is the HOST sampler (Python ``random``): it consumes the RNG exactly like the
reference (in-place shuffle of data.training_data, then one ``choice`` per draw over
the item names in dict order), so a run seeded like the reference sees the same
triples.  The default training path uses the on-device Philox sampler
(agcf_bpr_sample_epoch) instead; select this one with args.sampler='host' or
ARLIB_B200_SAMPLER=host.
"""
from random import shuffle, choice


def next_batch_pairwise(data, batch_size):
    rows = data.training_data
    shuffle(rows)
    total = len(rows)
    lo = 0
    while lo < total:
        hi = min(lo + batch_size, total)
        names = list(data.item.keys())
        u_idx, i_idx, j_idx = [], [], []
        for k in range(lo, hi):
            user, item = rows[k][0], rows[k][1]
            i_idx.append(data.item[item])
            u_idx.append(data.user[user])
            neg = choice(names)
            while neg in data.training_set_u[user]:
                neg = choice(names)
            j_idx.append(data.item[neg])
        lo = hi
        yield u_idx, i_idx, j_idx
