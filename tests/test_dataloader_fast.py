"""CPU: the vectorised ingest path (C parser + pandas.factorize + group-wise dict building) builds
exactly the object the reference's row-by-row loops build (util/FileIO.py:21-31, util/DataLoader.py:32-108):
same ids, same dict contents AND insertion orders (attacks iterate these dicts), same scipy matrices."""
import os
import types

import numpy as np
import pytest


def _write(path, rows):
    with open(path, "w") as fh:
        fh.writelines("%s %s %s\n" % tuple(r) for r in rows)


def _rows(seed, n, n_users, n_items, prefix=""):
    rng = np.random.default_rng(seed)
    rows = [[prefix + "u%d" % rng.integers(n_users), "i%d" % rng.integers(n_items), int(rng.integers(1, 6))] for _ in range(n)]
    rows += [rows[3][:2] + [1], rows[3][:2] + [4], rows[10][:2] + [2]]          # duplicates: the last weight wins
    return rows


def _same_nested(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert list(a[k].items()) == list(b[k].items()), k


@pytest.mark.parametrize("seed", [0, 1])
def test_vectorised_ingest_equals_row_by_row(tmp_path, seed):
    from arlib_b200.util.DataLoader import DataLoader
    from arlib_b200.util.FileIO import FileIO, TripleRows
    d = str(tmp_path) + "/"
    os.makedirs(d + "ds", exist_ok=True)
    train = _rows(seed, 5000, 300, 400)
    val = _rows(seed + 10, 300, 330, 420)                     # some users / items unseen in train
    test = _rows(seed + 20, 900, 330, 420)
    for name, rows in (("train.txt", train), ("val.txt", val), ("test.txt", test)):
        _write(d + "ds/" + name, rows)
    loaded = FileIO.load_data_set(d + "ds/train.txt")
    assert isinstance(loaded, TripleRows) and loaded.parsed_columns() is not None
    assert loaded == [[r[0], r[1], float(r[2])] for r in train]
    args = types.SimpleNamespace(data_path=d, dataset="ds", training_data="/train.txt", val_data="/val.txt", test_data="/test.txt")
    fast = DataLoader(args)
    as_float = lambda rows: [[r[0], r[1], float(r[2])] for r in rows]
    slow = DataLoader.from_rows(as_float(train), as_float(val), as_float(test), "ds")     # plain lists: the per-row loops
    assert fast._edges is not None and getattr(slow, "_edges", None) is None
    for attr in ("user", "item", "id2user", "id2item"):
        assert list(getattr(fast, attr).items()) == list(getattr(slow, attr).items()), attr
    for attr in ("training_set_u", "training_set_i", "val_set", "test_set"):
        _same_nested(getattr(fast, attr), getattr(slow, attr))
    assert fast.val_set_item == slow.val_set_item and fast.test_set_item == slow.test_set_item
    assert (fast.user_num, fast.item_num) == (slow.user_num, slow.item_num)
    assert fast.training_data == slow.training_data and type(fast.training_data[0]) is list
    for a, b in zip(fast.edge_arrays(), slow.edge_arrays()):
        assert np.array_equal(a, b)
    for attr in ("ui_adj", "norm_adj", "interaction_mat"):
        x, y = getattr(fast, attr).tocsr(), getattr(slow, attr).tocsr()
        x.sort_indices(); y.sort_indices()
        assert np.array_equal(x.indptr, y.indptr) and np.array_equal(x.indices, y.indices)
        assert np.array_equal(x.data.view(np.uint32), y.data.view(np.uint32)), attr
    # defaultdict semantics survive (the sampler touches fake users that have no entry)
    assert fast.training_set_u["nobody"] == {} and "nobody" in fast.training_set_u
    # an attack appends rows and a user: the cached edge arrays are dropped, not reused
    fast.training_data.append(["fake0", train[0][1], 1.0])
    fast.user["fake0"] = len(fast.user)
    u, i = fast.edge_arrays()
    assert len(u) == len(fast.training_data) and u[-1] == fast.user["fake0"]


def test_irregular_files_fall_back_to_the_reference_loop(tmp_path):
    from arlib_b200.util.FileIO import FileIO
    p = str(tmp_path / "ragged.txt")
    with open(p, "w") as fh:
        fh.write("a b 1\nc d 2 extra\n  e f 3  \n")
    assert FileIO.load_data_set(p) == [["a", "b", 1.0], ["c", "d", 2.0], ["e", "f", 3.0]]


def test_from_arrays_is_the_same_object_as_from_rows():
    from arlib_b200.util.DataLoader import DataLoader
    rng = np.random.default_rng(3)
    tu, ti = rng.integers(0, 50, 2000), rng.integers(0, 80, 2000)
    su, si = rng.integers(0, 60, 300), rng.integers(0, 90, 300)
    a = DataLoader.from_arrays(tu, ti, su, si)
    tr = [[str(int(u)), str(int(i)), 1.0] for u, i in zip(tu, ti)]
    te = [[str(int(u)), str(int(i)), 1.0] for u, i in zip(su, si)]
    b = DataLoader.from_rows(tr, te[:1000], te)
    assert list(a.user.items()) == list(b.user.items()) and list(a.item.items()) == list(b.item.items())
    _same_nested(a.training_set_u, b.training_set_u)
    _same_nested(a.test_set, b.test_set)
    assert a.training_data == b.training_data
    assert (a.interaction_mat != b.interaction_mat).nnz == 0


def test_device_mirrors_from_pristine_edges_equal_the_dict_walk():
    """The sampler's rejection lists and the evaluation masks are derived from the cached edge arrays while the data
    object is untouched, and from the (possibly stale -- SURVEY.md App. B) dicts once an attack resized it."""
    import torch
    from arlib_b200.engine import DeviceTrainSet
    from arlib_b200.evaluator import FullRankEvaluator
    from arlib_b200.util.DataLoader import DataLoader
    rng = np.random.default_rng(5)
    tu, ti = rng.integers(0, 200, 6000), rng.integers(0, 300, 6000)          # with duplicate pairs
    su, si = rng.integers(0, 220, 800), rng.integers(0, 320, 800)
    data = DataLoader.from_arrays(tu, ti, su, si)
    assert data.pristine_edges() is not None
    fast_ts, fast_ev = DeviceTrainSet(data, "cpu"), FullRankEvaluator(data, "cpu")
    edges = data._edges
    data._edges = None                                                        # force the dict-walking constructors
    slow_ts, slow_ev = DeviceTrainSet(data, "cpu"), FullRankEvaluator(data, "cpu")
    for k in ("e_user", "e_item", "rej_rowptr", "rej_items"):
        assert torch.equal(getattr(fast_ts, k), getattr(slow_ts, k)), k
    for k in ("user_rows", "mask_rowptr", "mask_items", "t_rowptr", "t_items", "test_total"):
        assert torch.equal(getattr(fast_ev, k), getattr(slow_ev, k)), k
    # fake-user injection (attack/White/CLeaR.py:179-197): rows + a user appended, training_set_u left stale
    data._edges = edges
    data.training_data.append(["fake", data.id2item[0], 1.0])
    data.user["fake"] = len(data.user)
    data.user_num += 1
    assert data.pristine_edges() is None
    ts = DeviceTrainSet(data, "cpu")
    assert ts.n_edges == len(data.training_data) and int(ts.rej_rowptr[-1]) == int(slow_ts.rej_rowptr[-1])   # fake user rejects nothing


def test_long_mantissa_weights_parse_like_float(tmp_path):
    """pandas' default fast float parser is not bit-identical to the reference's ``float(weight)``
    (util/FileIO.py:28); the loader asks for the round-trip parser."""
    import random as _r
    from arlib_b200.util.FileIO import FileIO
    rng = _r.Random(5)
    weights = [repr(rng.random() * 10 ** rng.randint(-8, 8)) for _ in range(20000)] + ["1", "0.1", "1e-7", "3.0000000000000004"]
    path = tmp_path / "w.txt"
    path.write_text("".join("u%d i%d %s\n" % (k % 97, k % 89, w) for k, w in enumerate(weights)))
    rows = FileIO.load_data_set(str(path))
    assert [r[2] for r in rows] == [float(w) for w in weights]
