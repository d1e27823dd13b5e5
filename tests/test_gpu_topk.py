"""GPU parity: fused scoring + masked top-K (+ tie rule, merge, predict, metrics)
against find_k_largest / ranking_evaluation of the oracle."""
import numpy as np
import pytest
import torch

from _util import assert_topk_matches
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["2", "1", "0"], ids=["stage2-group-major-items", "stage2-group-major", "stage2-warp-per-user"])
def stage2_impl(request, monkeypatch):
    """every test runs on all three stage-2 implementations (csrc/score.cu): they must agree bit for bit"""
    monkeypatch.setenv("AGCF_STAGE2_IMPL", request.param)
    return request.param
DEV = "cuda:0"


def _mask_csr(U, lists):
    rp = np.zeros(U + 1, dtype=np.int32)
    rp[1:] = np.cumsum([len(x) for x in lists])
    flat = np.concatenate([np.sort(np.asarray(x, dtype=np.int32)) for x in lists]) if rp[-1] else np.zeros(1, np.int32)
    return torch.from_numpy(rp).to(DEV), torch.from_numpy(flat.astype(np.int32)).to(DEV)


@pytest.mark.parametrize("U,I,d,K", [(70, 1000, 64, 50), (130, 333, 32, 10), (64, 4100, 128, 50), (5, 31, 64, 3),
                                      (33, 2049, 256, 100), (40, 70000, 32, 20), (24, 100001, 64, 50)])
def test_topk_exact_scores_and_sets(U, I, d, K):
    # (70 000 items: the candidate selection keeps a user's 2 188 group maxima in shared memory; 100 001 items: 3 126
    # groups no longer fit eight to a CTA and are read from L2 in every pass -- csrc/score.cu, cand_stage)
    from arlib_b200 import ops
    rng = np.random.default_rng(U + I)
    ue = torch.randn(U, d)
    ie = torch.randn(I, d)
    lists = [rng.choice(I, size=int(rng.integers(0, min(I - K, 60))), replace=False) for _ in range(U)]
    mrp, mit = _mask_csr(U, lists)
    vals, idx = ops.score_topk(ue.to(DEV), ie.to(DEV), K, mask_rowptr=mrp, mask_items=mit, impl=0)
    vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
    exact = ops.score_rows(ue.to(DEV), torch.arange(U, dtype=torch.int32, device=DEV), ie.to(DEV)).cpu().numpy()
    ref = (ue.double() @ ie.double().T).numpy()
    assert np.abs(exact - ref).max() < 1e-4
    for r in range(U):
        s = exact[r].copy()
        s[lists[r]] = -10e8
        # against the device's own exact fp32 scores the selection must be THE reference set
        assert set(idx[r].tolist()) == port.topk_reference_set(K, s)
        assert np.array_equal(vals[r], s[idx[r]])
        assert np.all(np.diff(vals[r]) <= 0)
        # and against float64 scores it is a valid top-K up to fp32 rounding
        s64 = ref[r].copy(); s64[lists[r]] = -10e8
        assert_topk_matches(idx[r], s64, K, 1e-4)


def test_topk_warp_kernel_overflow_goes_through_block_kernel(monkeypatch):
    """Stage 2 runs one warp per user with room for a bounded number of candidate groups; users beyond
    it are handed to the block-per-user kernel.  Forcing a tiny bound must not change a single bit."""
    from arlib_b200 import ops
    rng = np.random.default_rng(3)
    U, I, d, K = 300, 5000, 64, 50
    ue, ie = torch.randn(U, d, device=DEV), torch.randn(I, d, device=DEV)
    lists = [rng.choice(I, size=int(rng.integers(0, 40)), replace=False) for _ in range(U)]
    mrp, mit = _mask_csr(U, lists)
    for impl in (0, 1):
        monkeypatch.delenv("AGCF_STAGE2_CMAX", raising=False)
        v0, i0, f0 = ops.score_topk(ue, ie, K, mask_rowptr=mrp, mask_items=mit, impl=impl, return_flags=True)
        monkeypatch.setenv("AGCF_STAGE2_CMAX", "40")      # K = 50 groups are always candidates -> every user overflows
        v1, i1, f1 = ops.score_topk(ue, ie, K, mask_rowptr=mrp, mask_items=mit, impl=impl, return_flags=True)
        assert int(f0.min()) >= K and torch.equal(f0, f1)
        assert torch.equal(i0, i1) and torch.equal(v0, v1)
    monkeypatch.delenv("AGCF_STAGE2_CMAX", raising=False)


def test_topk_tie_rule_kat():
    """Heavy exact ties: integer-valued embeddings make many scores identical; the set
    must follow the reference heap's rule, not lowest/highest-id-first."""
    from arlib_b200 import ops
    rng = np.random.default_rng(7)
    U, I, d, K = 40, 200, 32, 20
    ue = torch.from_numpy(rng.integers(0, 2, (U, d)).astype(np.float32))
    ie = torch.from_numpy(rng.integers(0, 2, (I, d)).astype(np.float32))
    lists = [rng.choice(I, size=5, replace=False) for _ in range(U)]
    mrp, mit = _mask_csr(U, lists)
    vals, idx = ops.score_topk(ue.to(DEV), ie.to(DEV), K, mask_rowptr=mrp, mask_items=mit)
    scores = (ue @ ie.T).numpy()
    differs_from_lowest_first = 0
    for r in range(U):
        s = scores[r].copy(); s[lists[r]] = -10e8
        heap_ids, _ = port.find_k_largest_py(K, s)
        assert set(idx[r].cpu().tolist()) == set(heap_ids)
        lowest_first = set(np.argsort(-s, kind="stable")[:K].tolist())
        differs_from_lowest_first += set(heap_ids) != lowest_first
    assert differs_from_lowest_first > 0          # the KAT really exercises the odd rule


def test_topk_more_masked_than_free_and_k_ge_items():
    from arlib_b200 import ops
    U, I, d = 3, 40, 64
    ue, ie = torch.randn(U, d), torch.randn(I, d)
    lists = [np.arange(35), np.arange(0), np.arange(5, 40)]       # user 0 / 2: only 5 free items, K = 8
    mrp, mit = _mask_csr(U, lists)
    vals, idx = ops.score_topk(ue.to(DEV), ie.to(DEV), 8, mask_rowptr=mrp, mask_items=mit)
    scores = ops.score_rows(ue.to(DEV), torch.arange(U, dtype=torch.int32, device=DEV), ie.to(DEV)).cpu().numpy()
    for r in range(U):
        s = scores[r].copy(); s[lists[r]] = -10e8
        ids, _ = port.find_k_largest_py(8, s)
        assert set(idx[r].cpu().tolist()) == set(ids)
    vals, idx = ops.score_topk(ue.to(DEV), ie.to(DEV), 64)          # K > I: all items, then -1 padding
    assert sorted(idx[1].cpu().tolist()[:I]) == list(range(I)) and set(idx[1].cpu().tolist()[I:]) == {-1}


def test_golden_topk_sets_and_metric_strings(golden, golden_rows):
    """The reference's own top-50 lists and metric strings on ml-100k (frozen)."""
    from arlib_b200.evaluator import FullRankEvaluator
    from arlib_b200.util.DataLoader import DataLoader
    train, test = golden_rows
    data = DataLoader.from_rows([list(r) for r in train], (), test)
    ev = FullRankEvaluator(data, torch.device(DEV))
    fu = torch.from_numpy(golden["final_user_emb"]).to(DEV)
    fi = torch.from_numpy(golden["final_item_emb"]).to(DEV)
    vals, idx = ev.topk(fu, fi, 50)
    assert [int(u) for u in ev.users] == golden["topk_users"].tolist()
    names = golden["item_names"]
    scores = golden["final_user_emb"] @ golden["final_item_emb"].T
    near_ties = 0
    for k in range(idx.shape[0]):
        got = set(names[idx[k].cpu().numpy()].tolist())
        want = set(golden["topk_items"][k].tolist())
        if got != want:
            # only legal if the reference's own K-th / (K+1)-th scores are within fp32 rounding
            near_ties += 1
            diff = got ^ want
            ids = [int(np.flatnonzero(names == n)[0]) for n in diff]
            uid = data.user[ev.users[k]]
            sc = scores[uid][ids]
            assert np.ptp(sc) < 1e-6, "top-50 set differs beyond rounding for user %s" % ev.users[k]
    assert near_ties == 0, "boundary near-ties on the golden fixture: %d" % near_ties
    measure = ev.measure(idx, [50])
    want = [str(x) for x in golden["measure"]]
    assert measure[0] == want[0]
    for a, b in zip(measure[1:], want[1:]):
        ka, va = a.strip().split(":"); kb, vb = b.strip().split(":")
        assert ka == kb and abs(float(va) - float(vb)) < 1e-3
    assert measure[1:4] == want[1:4]              # hit ratio / precision / recall are integer-derived: exact strings


def test_topk_merge_matches_unsharded():
    from arlib_b200 import ops
    U, I, d, K = 50, 900, 64, 25
    ue, ie = torch.randn(U, d).to(DEV), torch.randn(I, d).to(DEV)
    full_v, full_i = ops.score_topk(ue, ie, K)
    parts_v, parts_i = [], []
    for lo, hi in ((0, 300), (300, 600), (600, 900)):
        v, i = ops.score_topk(ue, ie[lo:hi].contiguous(), K, item_offset=lo)
        parts_v.append(v); parts_i.append(i)
    mv, mi = ops.topk_merge(torch.stack(parts_v), torch.stack(parts_i))
    assert torch.equal(mi, full_i) and torch.equal(mv, full_v)


def test_rank_metrics_match_ranking_evaluation():
    from arlib_b200 import ops
    import math
    rng = np.random.default_rng(3)
    U, I, K = 60, 400, 50
    idx = np.stack([rng.permutation(I)[:K] for _ in range(U)]).astype(np.int32)
    origin, res = {}, {}
    t_lists, totals = [], []
    for r in range(U):
        seen = rng.choice(I, size=int(rng.integers(1, 30)), replace=False)
        unseen = int(rng.integers(0, 3))
        origin[str(r)] = {str(i): 1.0 for i in seen}
        for x in range(unseen):
            origin[str(r)]["unseen%d" % x] = 1.0
        res[str(r)] = [(str(i), 0.0) for i in idx[r]]
        t_lists.append(np.sort(seen).astype(np.int32)); totals.append(len(seen) + unseen)
    rp = np.zeros(U + 1, dtype=np.int32); rp[1:] = np.cumsum([len(x) for x in t_lists])
    cut = [10, 50]
    inv_log = torch.tensor([1.0 / math.log(r + 2) for r in range(K)], dtype=torch.float64, device=DEV)
    out = ops.rank_metrics(torch.from_numpy(idx).to(DEV), torch.from_numpy(rp).to(DEV),
                           torch.from_numpy(np.concatenate(t_lists)).to(DEV),
                           torch.tensor(totals, dtype=torch.int32, device=DEV),
                           torch.tensor(cut, dtype=torch.int32, device=DEV), inv_log).cpu().numpy()
    want = port.ranking_evaluation(origin, res, cut)
    for c, n in enumerate(cut):
        hits = out[:, c, 0]
        hr = hits.sum() / sum(totals)
        ndcg = 0
        for r in range(U):
            ndcg += out[r, c, 1] / out[r, c, 2]
        assert want[5 * c + 1] == 'Hit Ratio:' + str(hr) + '\n'
        assert want[5 * c + 4] == 'NDCG:' + str(ndcg / U) + '\n'


@pytest.mark.parametrize("U,I,d,K", [(300, 3000, 64, 50), (129, 257, 64, 10), (1000, 5000, 32, 20), (260, 1500, 128, 50),
                                      (5, 40, 64, 3)])
def test_tcgen05_stage1_gives_the_same_topk_as_the_fp32_path(U, I, d, K):
    """impl=1 (TF32 tcgen05 group-max GEMM + margin + exact re-score) must return exactly
    what impl=0 (fp32 CUDA-core stage 1) returns: the tensor-core rounding never leaks out."""
    from arlib_b200 import ops
    rng = np.random.default_rng(U * 7 + I)
    ue = torch.randn(U, d).to(DEV)
    ie = (torch.randn(I, d) * torch.rand(I, 1) * 2).to(DEV)          # varied item norms
    lists = [rng.choice(I, size=int(rng.integers(0, min(I - K, 80))), replace=False) for _ in range(U)]
    mrp, mit = _mask_csr(U, lists)
    rows = torch.from_numpy(rng.permutation(U).astype(np.int32)).to(DEV)
    for user_rows in (None, rows):
        v0, i0, f0 = ops.score_topk(ue, ie, K, user_rows=user_rows, mask_rowptr=mrp, mask_items=mit, impl=0, return_flags=True)
        v1, i1, f1 = ops.score_topk(ue, ie, K, user_rows=user_rows, mask_rowptr=mrp, mask_items=mit, impl=1, return_flags=True)
        assert torch.equal(i0, i1) and torch.equal(v0, v1)
        # the TF32 margin admits a few more candidate groups, not an explosion
        assert float(f1.float().mean()) <= 3.0 * float(f0.float().mean()) + 4


def test_attack_metric_matches_reference_formulas():
    """AttackMetric (reference util/metrics.py:125-207) from ONE fused score + top-K pass over
    all users instead of 4 x U predict() + argsort; values equal the oracle restatement of the reference
    (itself checked against the unmodified reference class when the oracle was written)."""
    import types
    from arlib_b200 import ops
    from arlib_b200.util.metrics import AttackMetric
    rng = np.random.default_rng(11)
    U, I, d = 300, 2000, 64
    ue = torch.randn(U, d, device=DEV)
    ie = torch.randn(I, d, device=DEV)
    users = {"u%d" % k: k for k in rng.permutation(U)}
    rec = types.SimpleNamespace(data=types.SimpleNamespace(user=users), user_emb=ue, item_emb=ie)
    exact = ops.score_rows(ue, torch.arange(U, dtype=torch.int32, device=DEV), ie).cpu().numpy()
    targets = [5, 77, 1203, 1999]
    # make the targets competitive so that the statistics are not all zero
    ie_boost = ie.clone(); ie_boost[targets] *= 2.5
    rec.item_emb = ie_boost
    exact = ops.score_rows(ue, torch.arange(U, dtype=torch.int32, device=DEV), ie_boost).cpu().numpy()
    for top in ([10], [5, 20, 50]):
        am = AttackMetric(rec, targets, top)
        ref = port.attack_metric(lambda name: exact[users[name]], list(users), targets, top)
        for name in ("precision", "hitRate", "recall", "NDCG"):
            got = getattr(am, name)()
            assert max(ref[name]) > 0
            np.testing.assert_allclose(got, ref[name], rtol=1e-12, atol=0)
    with pytest.raises(TypeError):
        AttackMetric(types.SimpleNamespace(data=rec.data, user_emb=ue.cpu(), item_emb=ie.cpu()), targets).precision()


def test_masked_score_topk_equals_the_attacks_dense_topk():
    """attack/White/CLeaR.py:75-81 -- dense scores, interaction entries := -10e8, torch.topk."""
    import scipy.sparse as sp
    from arlib_b200.util.algorithm import masked_score_topk
    rng = np.random.default_rng(4)
    U, I, d, K = 500, 1300, 64, 50
    Pu = torch.from_numpy(rng.standard_normal((U, d)).astype(np.float32))
    Pi = torch.from_numpy(rng.standard_normal((I, d)).astype(np.float32))
    dense = sp.random(U, I, density=0.05, random_state=1, dtype=np.float32).toarray()
    dense[3, :] = 1.0                                  # a user with everything masked
    dense[4, :I - 7] = 1.0                             # fewer free items than K
    inter = sp.csr_matrix(dense)
    rows_of = np.repeat(np.arange(U), np.diff(inter.indptr))
    stored_zero = (np.arange(inter.nnz) % 11 == 0) & (rows_of != 3) & (rows_of != 4)
    inter.data[stored_zero] = 0.0                      # explicitly stored zeros are NOT masked (.nonzero())
    scores = Pu.double() @ Pi.double().T
    nz = inter.nonzero()
    scores[nz[0], nz[1]] = -10e8
    ref_v, ref_i = torch.topk(scores, K)
    vals, idx = masked_score_topk(Pu.cuda(), Pi.cuda(), K, inter)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (U, K)
    idx, vals = idx.cpu(), vals.cpu()
    for u in range(U):
        if u == 3:
            assert bool((vals[u] == -10e8).all())
            continue
        assert_topk_matches(idx[u].numpy(), scores[u].numpy(), K, 1e-4)
        free = int((ref_v[u] > -1e8).sum())
        np.testing.assert_allclose(vals[u, :free].numpy(), ref_v[u, :free].numpy(), rtol=1e-5, atol=1e-5)
    vals2, idx2 = masked_score_topk(Pu.cuda(), Pi.cuda(), K, None)
    ref_v2, _ = torch.topk(Pu.double() @ Pi.double().T, K)
    np.testing.assert_allclose(vals2.cpu().numpy(), ref_v2.numpy(), rtol=1e-5, atol=1e-5)


def test_stage2_implementations_are_bit_identical(monkeypatch):
    from arlib_b200 import ops
    rng = np.random.default_rng(8)
    U, I, d, K = 3000, 9000, 64, 50
    ue = torch.from_numpy(rng.standard_normal((U, d)).astype(np.float32)).to(DEV)
    ie = torch.from_numpy((rng.standard_normal((I, d)) * (1 + 3 * rng.random((I, 1)) ** 4)).astype(np.float32)).to(DEV)
    lists = [np.sort(rng.choice(I, rng.integers(0, 200), replace=False)) for _ in range(U)]
    mrp, mit = _mask_csr(U, lists)
    outs = {}
    for impl2 in ("2", "1", "0"):
        monkeypatch.setenv("AGCF_STAGE2_IMPL", impl2)
        for impl1 in (1, 0):
            outs[(impl2, impl1)] = ops.score_topk(ue, ie, K, mask_rowptr=mrp, mask_items=mit, impl=impl1, return_flags=True)
    ref = outs[("0", 0)]
    for key, got in outs.items():
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]), key
    assert torch.equal(outs[("1", 1)][2], outs[("0", 1)][2])               # same candidate-group counts
    assert torch.equal(outs[("2", 1)][2], outs[("0", 1)][2])


def test_item_compacting_stage2_with_heavy_ties_and_few_free_items(monkeypatch):
    """The default stage 2 keeps only the exact scores >= T_u per user (<= 128 entries); users it cannot hold -- exact
    ties by the hundred, or fewer than K unmasked items so that masked -1e9 entries are wanted -- fall through to the
    block kernel.  Every variant must return the same bits."""
    from arlib_b200 import ops
    rng = np.random.default_rng(21)
    U, I, d, K = 64, 8000, 64, 50
    ue = torch.from_numpy(rng.integers(0, 2, (U, d)).astype(np.float32)).to(DEV)        # integer scores: ties everywhere
    ie = torch.from_numpy(rng.integers(0, 2, (I, d)).astype(np.float32)).to(DEV)
    ue[40:] = torch.randn(U - 40, d, device=DEV)
    lists = [np.sort(rng.choice(I, rng.integers(0, 100), replace=False)) for _ in range(U)]
    lists[50] = np.arange(I - 10)                                                        # 10 free items < K
    lists[51] = np.arange(3, I)
    mrp, mit = _mask_csr(U, lists)
    outs = {}
    for impl2 in ("2", "1", "0"):
        monkeypatch.setenv("AGCF_STAGE2_IMPL", impl2)
        for impl1 in (1, 0):
            outs[(impl2, impl1)] = ops.score_topk(ue, ie, K, mask_rowptr=mrp, mask_items=mit, impl=impl1)
    ref = outs[("0", 0)]
    for key, got in outs.items():
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]), key
    scores = ops.score_rows(ue, torch.arange(U, dtype=torch.int32, device=DEV), ie).cpu().numpy()
    idx = ref[1].cpu().numpy()
    for r in (0, 7, 39, 45, 50, 51):
        s_ = scores[r].copy(); s_[lists[r]] = -10e8
        assert set(idx[r].tolist()) == port.topk_reference_set(K, s_)


def test_group_max_stage_alone_and_the_tf32_error_bound():
    """agcf_score_group_max: stage 1 on its own.  impl 0 equals the masked fp32 group maxima of the exact scores bit for
    bit; impl 1 (tcgen05 TF32) stays within delta = 1.01 * 2^-9 * |u| * max|v| of them -- the premise of the top-K
    exactness proof in csrc/score.cu."""
    from arlib_b200 import ops
    rng = np.random.default_rng(4)
    U, I, d = 300, 5000, 64
    ue = torch.from_numpy(rng.standard_normal((U, d)).astype(np.float32)).to(DEV)
    ie = torch.from_numpy((rng.standard_normal((I, d)) * (1 + 2 * rng.random((I, 1)) ** 3)).astype(np.float32)).to(DEV)
    lists = [np.sort(rng.choice(I, rng.integers(0, 80), replace=False)) for _ in range(U)]
    mrp, mit = _mask_csr(U, lists)
    g0 = ops.score_group_max(ue, ie, mask_rowptr=mrp, mask_items=mit, impl=0)
    g1 = ops.score_group_max(ue, ie, mask_rowptr=mrp, mask_items=mit, impl=1)
    exact = ops.score_rows(ue, torch.arange(U, dtype=torch.int32, device=DEV), ie)
    for r, l in enumerate(lists):
        exact[r, torch.from_numpy(l.astype(np.int64)).to(DEV)] = -1.0e9
    pad = (-I) % 32
    ref = torch.cat([exact, torch.full((U, pad), -3.4e38, device=DEV)], 1).view(U, -1, 32).max(2).values
    assert g0.shape == ref.shape and torch.equal(g0, ref)
    delta = 1.01 * 2.0 ** -9 * ue.norm(dim=1, keepdim=True) * ie.norm(dim=1).max()
    assert bool(((g1 - ref).abs() <= delta).all())
    assert float((g1 - ref).abs().max()) > 0                     # it IS approximate: stage 2 is what makes the result exact


def test_evaluator_keeps_the_mask_bits_between_evaluations():
    """FullRankEvaluator re-uses its workspace: from the second evaluation on stage 0 (memset + train-item bit scatter) is
    skipped (AGCF_TOPK_KEEP_MASK_BITS).  Every evaluation must still equal a call on a fresh workspace, also after a
    call with other users / another table shape went through the same workspace in between."""
    from arlib_b200 import ops
    from arlib_b200.evaluator import FullRankEvaluator
    rng = np.random.default_rng(11)
    U, I, d, K = 700, 3000, 64, 50
    tu = rng.integers(0, U, 20000); ti = rng.integers(0, I, 20000)
    su = rng.integers(0, U, 2000); si = rng.integers(0, I, 2000)
    ev = FullRankEvaluator.from_arrays(U, I, tu, ti, su, si, torch.device(DEV))
    for trial in range(4):
        ue = torch.randn(U, d, device=DEV); ie = torch.randn(I, d, device=DEV)
        v, i = ev.topk(ue, ie, K)
        v0, i0 = ops.score_topk(ue, ie, K, user_rows=ev.user_rows, mask_rowptr=ev.mask_rowptr, mask_items=ev.mask_items, impl=1)
        assert torch.equal(i, i0) and torch.equal(v, v0), trial
        if trial == 1:
            # another shape through the same workspace: the key changes, the next evaluation rebuilds its bits
            ie2 = torch.randn(I - 500, d, device=DEV)
            v2, i2 = ev.topk(ue, ie2, K)
            w2, j2 = ops.score_topk(ue, ie2, K, user_rows=ev.user_rows, mask_rowptr=ev.mask_rowptr, mask_items=ev.mask_items, impl=1)
            assert torch.equal(i2, j2) and torch.equal(v2, w2)
