"""Interaction data model -- drop-in for the reference's util/DataLoader.py:7-177.

Same constructor, attributes and methods (attacks read and MUTATE these in place:
``user_num``, ``user``, ``id2user``, ``training_data``, ``interaction_mat``,
``ui_adj``, ``norm_adj`` -- attack/White/CLeaR.py:179-197, PGA.py:183-184), so the
object stays a plain Python / scipy structure.  Device mirrors (CSR of the
normalized adjacency, edge arrays, rejection lists) are derived from it on demand
by arlib_b200.graph / arlib_b200.engine and never stored here, which keeps the
object picklable and deep-copyable (ARLib.py:128,241).
"""
from collections import defaultdict
from itertools import islice

import numpy as np
import scipy.sparse as sp

from .FileIO import FileIO, no_gc, rows_from_columns


class DataLoader(object):
    def __init__(self, args):
        base = args.data_path + args.dataset
        self._setup(FileIO.load_data_set(base + args.training_data),
                    FileIO.load_data_set(base + args.val_data),
                    FileIO.load_data_set(base + args.test_data), args.dataset)

    @classmethod
    def from_rows(cls, training_data, val_data=(), test_data=(), name="synthetic"):
        """Build from in-memory [user, item, weight] rows (no files) -- used by the
        benchmark and the tests with synthetic data."""
        self = cls.__new__(cls)
        self._setup(list(training_data), list(val_data), list(test_data), name)
        return self

    @classmethod
    def from_arrays(cls, train_u, train_i, test_u=(), test_i=(), name="synthetic"):
        """Integer (user, item) arrays -> rows with string names str(id)."""
        def rows(u, i):
            names = lambda a: np.asarray(a, dtype=np.int64).astype(str).astype(object)
            return rows_from_columns(names(u), names(i), np.ones(len(u), dtype=np.float64))
        tr, te = rows(train_u, train_i), rows(test_u, test_i)
        self = cls.__new__(cls)
        self._setup(tr, te[:1000], te, name)
        return self

    # ------------------------------------------------------------------ build
    def _setup(self, training_data, val_data, test_data, name):
        self.training_data = training_data
        self.val_data = val_data
        self.test_data = test_data
        self.dataName = name
        self.user, self.item = {}, {}
        self.id2user, self.id2item = {}, {}
        self.training_set_u = defaultdict(dict)
        self.training_set_i = defaultdict(dict)
        self.val_set = defaultdict(dict)
        self.val_set_item = set()
        self.test_set = defaultdict(dict)
        self.test_set_item = set()
        self._edges = None
        cols = getattr(training_data, 'parsed_columns', lambda: None)()
        with no_gc():
            if cols is not None:
                self._index_interactions_columns(cols)
            else:
                self._index_interactions()
        self.user_num = len(self.training_set_u)
        self.item_num = len(self.training_set_i)
        self.ui_adj = self._bipartite_adjacency()
        self.norm_adj = self.normalize_graph_mat(self.ui_adj)
        self.interaction_mat = self._interaction_matrix()

    def _index_interactions(self):
        """ids by first appearance in train (util/DataLoader.py:32-40); val/test rows of
        users unseen in train are dropped (:41-55)."""
        user, item = self.user, self.item
        for row in self.training_data:
            u, i, r = row[0], row[1], row[2]
            if u not in user:
                user[u] = len(user)
                self.id2user[user[u]] = u
            if i not in item:
                item[i] = len(item)
                self.id2item[item[i]] = i
            self.training_set_u[u][i] = r
            self.training_set_i[i][u] = r
        for rows, bucket, seen in ((self.val_data, self.val_set, self.val_set_item),
                                   (self.test_data, self.test_set, self.test_set_item)):
            for row in rows:
                if row[0] in user:
                    bucket[row[0]][row[1]] = row[2]
                    seen.add(row[1])

    @staticmethod
    def _grouped_dicts(codes, n_keys, key_names, inner_names, weights):
        """{key_names[k]: {inner: weight, ...}} with the insertion orders of the row-by-row loop: outer keys by first
        appearance (= code order), inner keys by first appearance inside the group, later duplicates overwrite the
        value -- built from one stable sort and one dict(zip(...)) per group instead of one dict update per row."""
        order = np.argsort(codes, kind='stable')
        bounds = np.searchsorted(codes[order], np.arange(n_keys + 1))
        inner = iter(inner_names[order].tolist())
        w = iter(weights[order].tolist())
        out = defaultdict(dict)
        for name, cnt in zip(key_names, np.diff(bounds).tolist()):
            out[name] = dict(zip(islice(inner, cnt), islice(w, cnt)))
        return out

    def _index_interactions_columns(self, cols):
        """Same result as _index_interactions (asserted attribute by attribute in tests/test_dataloader_fast.py), from
        the parsed columns of the training rows: ids with pandas.factorize (first appearance, exactly the reference's
        numbering util/DataLoader.py:32-40), the dict-of-dicts group-wise."""
        import pandas as pd
        users, items, weights = cols
        ucode, unames = pd.factorize(users)
        icode, inames = pd.factorize(items)
        unames, inames = unames.tolist(), inames.tolist()
        self.user.update(zip(unames, range(len(unames))))
        self.item.update(zip(inames, range(len(inames))))
        self.id2user.update(enumerate(unames))
        self.id2item.update(enumerate(inames))
        self.training_set_u = self._grouped_dicts(ucode, len(unames), unames, items, weights)
        self.training_set_i = self._grouped_dicts(icode, len(inames), inames, users, weights)
        self._edges = (len(self.training_data), len(unames), len(inames), ucode.astype(np.int64), icode.astype(np.int64))
        user = self.user
        for rows, bucket, seen in ((self.val_data, self.val_set, self.val_set_item),
                                   (self.test_data, self.test_set, self.test_set_item)):
            c = getattr(rows, 'parsed_columns', lambda: None)()
            if c is None:
                for row in rows:
                    if row[0] in user:
                        bucket[row[0]][row[1]] = row[2]
                        seen.add(row[1])
                continue
            keep = np.fromiter((u in user for u in c[0].tolist()), dtype=bool, count=len(c[0]))
            ku, ki, kw = c[0][keep], c[1][keep], c[2][keep]
            code, names = pd.factorize(ku)
            bucket.update(self._grouped_dicts(code, len(names), names.tolist(), ki, kw))
            seen.update(ki.tolist())

    def edge_arrays(self):
        """(user ids, item ids) of training_data as int64 arrays, in list order (cached from the vectorised build for
        as long as nobody resized training_data; an in-place shuffle permutes rows, which no consumer depends on)."""
        n = len(self.training_data)
        e = getattr(self, '_edges', None)
        if e is not None and e[:3] == (n, len(self.user), len(self.item)):
            return e[3], e[4]
        u = np.fromiter((self.user[r[0]] for r in self.training_data), dtype=np.int64, count=n)
        i = np.fromiter((self.item[r[1]] for r in self.training_data), dtype=np.int64, count=n)
        return u, i

    def pristine_edges(self):
        """(user ids, item ids) from the vectorised build if training_data / user / item still have the sizes they
        were built with -- i.e. training_set_u is still exactly the set of these edges -- else None.  Lets the
        device mirrors (sampler rejection lists, evaluation masks) be derived with array operations; after an
        attack appended rows they go back to reading the dicts (whose staleness is a quirk callers rely on)."""
        e = getattr(self, '_edges', None)
        if e is not None and e[:3] == (len(self.training_data), len(self.user), len(self.item)) \
                and len(self.training_set_u) == len(self.user):
            return e[3], e[4]
        return None

    def _bipartite_adjacency(self, self_connection=False):
        """util/DataLoader.py:57-71 -- [[0,R],[R^T,0]] with fp32 ones."""
        n = self.user_num + self.item_num
        u, i = self.edge_arrays()
        half = sp.csr_matrix((np.ones_like(u, dtype=np.float32), (u, i + self.user_num)),
                             shape=(n, n), dtype=np.float32)
        adj = half + half.T
        if self_connection:
            adj += sp.eye(n)
        return adj

    def normalize_graph_mat(self, adj_mat):
        """util/DataLoader.py:73-87 -- kept on the host with the reference's exact numpy
        expressions: the degree vector must be bit-identical (np.power(x,-0.5) differs
        from 1/sqrt(x) in the last bit), see DESIGN.md "adjacency"."""
        shape = adj_mat.get_shape()
        rowsum = np.array(adj_mat.sum(1))
        if shape[0] == shape[1]:
            d_inv = np.power(rowsum, -0.5).flatten()
            d_inv[np.isinf(d_inv)] = 0.
            d_mat = sp.diags(d_inv)
            return d_mat.dot(adj_mat).dot(d_mat)
        d_inv = np.power(rowsum, -1).flatten()
        d_inv[np.isinf(d_inv)] = 0.
        return sp.diags(d_inv).dot(adj_mat)

    def convert_to_laplacian_mat(self, adj_mat):
        """util/DataLoader.py:89-96"""
        shape = adj_mat.get_shape()
        n = shape[0] + shape[1]
        rows, cols = adj_mat.nonzero()
        half = sp.csr_matrix((adj_mat.data, (rows, cols + shape[0])), shape=(n, n), dtype=np.float32)
        return self.normalize_graph_mat(half + half.T)

    def _interaction_matrix(self):
        """util/DataLoader.py:98-108 -- U x I CSR of ones (duplicates sum)."""
        u, i = self.edge_arrays()
        return sp.csr_matrix((np.ones(len(u), dtype=np.float32), (u, i)),
                             shape=(self.user_num, self.item_num), dtype=np.float32)

    # ---------------------------------------------------------------- queries
    def get_user_id(self, u):
        return self.user.get(u)

    def get_item_id(self, i):
        return self.item.get(i)

    def training_size(self):
        return len(self.user), len(self.item), len(self.training_data)

    def val_size(self):
        return len(self.val_set), len(self.val_set_item), len(self.val_data)

    def test_size(self):
        return len(self.test_set), len(self.test_set_item), len(self.test_data)

    def contain(self, u, i):
        return u in self.user and i in self.training_set_u[u]

    def contain_user(self, u):
        return u in self.user

    def contain_item(self, i):
        return i in self.item

    def user_rated(self, u):
        d = self.training_set_u[u]
        return list(d.keys()), list(d.values())

    def item_rated(self, i):
        d = self.training_set_i[i]
        return list(d.keys()), list(d.values())

    def row(self, u):
        names, vals = self.user_rated(self.id2user[u])
        vec = np.zeros(len(self.item))
        for n, v in zip(names, vals):
            vec[self.item[n]] = v
        return vec

    def col(self, i):
        names, vals = self.item_rated(self.id2item[i])
        vec = np.zeros(len(self.user))
        for n, v in zip(names, vals):
            vec[self.user[n]] = v
        return vec

    def matrix(self):
        return self._interaction_matrix()
