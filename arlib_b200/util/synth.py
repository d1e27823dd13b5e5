"""Synthetic power-law implicit-feedback graphs of named dataset shapes
(SURVEY.md 8d): the benchmark / test input generator.  Product-side (bench.py and
tests feed the SAME arrays to the CUDA path and to the oracle)."""
import numpy as np

SHAPES = {
    # name: (users, items, edges)
    "ml-100k": (943, 1682, 100000),
    "gowalla": (29858, 40981, 1027370),
    "yelp2018": (31668, 38048, 1561406),
    "ml-1m": (6040, 3706, 1000209),
    "amazon-book": (52643, 91599, 2984108),
}


def synth_edges(user_num, item_num, n_edges, alpha_u=0.5, alpha_i=0.5, seed=0, test_frac=0.1):
    """Deterministic synthetic implicit-feedback graph of a named shape.

    user activity ~ rank^-alpha_u, item popularity ~ rank^-alpha_i (ranks
    randomly permuted), pairs drawn independently and de-duplicated until
    exactly ``n_edges`` unique train pairs exist; every user and item gets >= 1
    edge; a held-out test set of test_frac*n_edges unique pairs not in train.
    Returns int64 arrays (train_u, train_i, test_u, test_i), train shuffled.
    """
    rng = np.random.default_rng(seed)
    pu = (np.arange(1, user_num + 1, dtype=np.float64) ** -alpha_u)
    pi = (np.arange(1, item_num + 1, dtype=np.float64) ** -alpha_i)
    pu = pu[rng.permutation(user_num)]
    pi = pi[rng.permutation(item_num)]
    pu /= pu.sum()
    pi /= pi.sum()
    cu, ci = np.cumsum(pu), np.cumsum(pi)
    n_test = int(round(test_frac * n_edges))
    want = n_edges + n_test
    # coverage edges first: one per user, one per item
    cov_u = np.concatenate([np.arange(user_num), np.minimum(np.searchsorted(cu, rng.random(item_num)), user_num - 1)])
    cov_i = np.concatenate([np.minimum(np.searchsorted(ci, rng.random(user_num)), item_num - 1), np.arange(item_num)])
    keys = np.unique(cov_u.astype(np.int64) * item_num + cov_i)
    n_cov = keys.shape[0]
    cov_keys = keys.copy()
    while keys.shape[0] < want:
        need = want - keys.shape[0]
        m = int(need * 1.3) + 1024
        u = np.minimum(np.searchsorted(cu, rng.random(m)), user_num - 1).astype(np.int64)
        i = np.minimum(np.searchsorted(ci, rng.random(m)), item_num - 1).astype(np.int64)
        keys = np.unique(np.concatenate([keys, u * item_num + i]))
    # coverage pairs must stay in train; the rest is split at random
    is_cov = np.isin(keys, cov_keys, assume_unique=True)
    rest = keys[~is_cov]
    rest = rest[rng.permutation(rest.shape[0])]
    n_train_rest = n_edges - n_cov
    train = np.concatenate([cov_keys, rest[:n_train_rest]])
    test = rest[n_train_rest:n_train_rest + n_test]
    train = train[rng.permutation(train.shape[0])]
    return train // item_num, train % item_num, test // item_num, test % item_num


def synth_edges_device(user_num, item_num, n_edges, alpha_u=0.5, alpha_i=0.5, seed=0, device="cuda", chunk=1 << 26):
    """The same power-law model drawn ON THE DEVICE (torch), for graphs the host generator and the reference's
    dict-of-dicts loader cannot hold (BASELINE.json configs[4]: 10 M users x 1 M items; SURVEY.md 8 fixes E = 200 M):
    user activity ~ rank^-alpha_u and item popularity ~ rank^-alpha_i over randomly permuted ranks, pairs drawn
    independently by inverse-CDF search and de-duplicated, one coverage edge per user and per item.  Returns int64
    device tensors (u, i) of ~n_edges UNIQUE pairs (exactly n_edges when enough distinct pairs were drawn; the pairs
    are sorted by user, then item)."""
    import torch
    gen = torch.Generator(device=device).manual_seed(seed)
    dev = torch.device(device)

    def cdf(n, alpha):
        p = torch.arange(1, n + 1, dtype=torch.float64, device=dev) ** (-alpha)
        p = p[torch.randperm(n, generator=gen, device=dev)]
        return torch.cumsum(p / p.sum(), 0)

    cu, ci = cdf(user_num, alpha_u), cdf(item_num, alpha_i)

    def draw(c, n, m):
        r = torch.rand(m, dtype=torch.float64, device=dev, generator=gen)
        return torch.searchsorted(c, r).clamp_(max=n - 1)

    cov = torch.cat([torch.arange(user_num, device=dev) * item_num + draw(ci, item_num, user_num),
                     draw(cu, user_num, item_num) * item_num + torch.arange(item_num, device=dev)])
    keys = torch.unique(cov)
    while keys.numel() < n_edges:
        need = n_edges - keys.numel()
        parts = [keys]
        m = int(need * 1.15) + 1024
        for lo in range(0, m, chunk):
            k = min(chunk, m - lo)
            parts.append(draw(cu, user_num, k) * item_num + draw(ci, item_num, k))
        keys = torch.unique(torch.cat(parts))
        del parts
    if keys.numel() > n_edges:                      # drop a random surplus, never a coverage edge
        is_cov = torch.isin(keys, cov)
        extra = torch.nonzero(~is_cov).flatten()
        drop = extra[torch.randperm(extra.numel(), generator=gen, device=dev)[:keys.numel() - n_edges]]
        keep = torch.ones(keys.numel(), dtype=torch.bool, device=dev)
        keep[drop] = False
        keys = keys[keep]
    return keys // item_num, keys % item_num
