"""Device mirror of the normalized adjacency: int32 CSR + fp32 values + the row
processing plan of the SpMM kernel.

Reference objects this mirrors: ``data.norm_adj`` (scipy, util/DataLoader.py:73-87),
``LGCN_Encoder.sparse_norm_adj`` (torch sparse COO, recommender/LightGCN.py:210,
212-215, 247-252).  Bit-exactness contract: the degree vectors are
computed on the host with the reference's own numpy expression; the O(nnz) products
``fl(fl(d_i*w)*d_j)`` run on device (agcf_norm_adj_csr) and reproduce scipy's
``(D.A).D`` association bit for bit.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib

LONG_ROW_THRESHOLD = int(os.environ.get("ARLIB_B200_LONG_ROW", "256"))


def _plan_rows(indptr: np.ndarray):
    """row_order = rows by degree descending (stable); n_long = #rows above the
    long-row threshold (those get a whole CTA in the SpMM kernel)."""
    deg = np.diff(indptr)
    order = np.argsort(-deg, kind="stable").astype(np.int32)
    n_long = int(np.count_nonzero(deg > LONG_ROW_THRESHOLD))
    return order, n_long


def balanced_row_ranges(indptr, world):
    """Boundaries b[0..world] with b[0]=0, b[world]=n_rows and ~nnz/world non-zeros per range."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.shape[0] - 1
    nnz = int(indptr[-1])
    bounds = [0]
    for p in range(1, world):
        target = nnz * p // world
        b = int(np.searchsorted(indptr, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return bounds


class DeviceGraph:
    """CSR (rowptr, col, val) on the device + SpMM plan.  Immutable after build."""

    def __init__(self, rowptr, col, val, n_rows, row_order, n_long):
        self.rowptr, self.col, self.val = rowptr, col, val
        self.n_rows = int(n_rows)                    # global row count (= rows of X / Y)
        self.nnz = int(col.numel())
        self.row_order, self.n_long = row_order, int(n_long)
        self.n_local_rows = int(row_order.numel())   # rows this graph computes (all, or one rank's partition)
        self._coo_idx = None
        self._build_plan()

    def _build_plan(self):
        """Slot-ordered copy of the CSR for the SpMM kernel: rows physically permuted into
        processing order (degree descending), so a task's rowptr / col / val reads are
        contiguous and independent of the row-id lookup (no dependent-load chain)."""
        order = self.row_order.long()
        deg = (self.rowptr[1:] - self.rowptr[:-1]).long()
        pdeg = deg[order]
        prowptr = torch.zeros(self.n_local_rows + 1, dtype=torch.int64, device=self.device)
        prowptr[1:] = torch.cumsum(pdeg, 0)
        # source position of every nnz in slot order: start of its row + offset inside the row
        starts = self.rowptr[:-1].long()[order]
        local_nnz = int(prowptr[-1])
        src = torch.repeat_interleave(starts - prowptr[:-1], pdeg) + torch.arange(local_nnz, device=self.device)
        self.local_nnz = local_nnz
        self.p_rowptr = prowptr.to(torch.int32)
        self.p_src = src
        self.p_col = self.col[src].contiguous()
        self.p_val = self.val[src].contiguous()

    def refresh_plan_values(self):
        """re-gather the slot-ordered values after ``val`` was modified in place"""
        self.p_val = self.val[self.p_src].contiguous()

    @property
    def device(self):
        return self.val.device

    # -------------------------------------------------------------- partitioning
    def row_ranges(self, world):
        """world+1 boundaries of contiguous row ranges with ~equal nnz (power-law rows:
        balance by non-zeros, not by rows)."""
        return balanced_row_ranges(self.rowptr.cpu().numpy(), world)

    def partition(self, r0, r1):
        """Graph that COMPUTES only rows [r0, r1) (one rank of the row-partitioned
        multi-GPU path); X / Y keep global row ids, col ids are global."""
        deg = (self.rowptr[r0 + 1:r1 + 1] - self.rowptr[r0:r1]).cpu().numpy()
        order = (np.argsort(-deg, kind="stable") + r0).astype(np.int32)
        n_long = int(np.count_nonzero(deg > LONG_ROW_THRESHOLD))
        return DeviceGraph(self.rowptr, self.col, self.val, self.n_rows, torch.from_numpy(order).to(self.device), n_long)

    # ---------------------------------------------------------------- builders
    @classmethod
    def _from_csr_host(cls, csr: sp.csr_matrix, device, values=None):
        if csr.shape[0] != csr.shape[1]:
            raise ValueError("adjacency must be square")
        if csr.nnz >= 2 ** 31 or csr.shape[0] >= 2 ** 31:
            raise ValueError("graph exceeds int32 indexing")
        order, n_long = _plan_rows(csr.indptr)
        dev = torch.device(device)
        rowptr = torch.from_numpy(csr.indptr.astype(np.int32)).to(dev)
        col = torch.from_numpy(csr.indices.astype(np.int32)).to(dev)
        if values is None:
            val = torch.from_numpy(np.ascontiguousarray(csr.data, dtype=np.float32)).to(dev)
        else:
            val = values
        return cls(rowptr, col, val, csr.shape[0], torch.from_numpy(order).to(dev), n_long)

    @classmethod
    def from_scipy(cls, mat, device="cuda"):
        """Upload an already-normalized scipy matrix as is (values untouched)."""
        csr = sp.csr_matrix(mat)
        csr.sum_duplicates()
        csr.sort_indices()
        return cls._from_csr_host(csr, device)

    @classmethod
    def normalized(cls, adj, d_row, d_col, device="cuda"):
        """values = fl(fl(d_row[i]*w)*d_col[j]) computed ON DEVICE from the raw
        weights of ``adj`` (scipy, canonical CSR) and host-computed degree vectors."""
        csr = sp.csr_matrix(adj)
        csr.sum_duplicates()
        csr.sort_indices()
        dev = torch.device(device)
        w = torch.from_numpy(np.ascontiguousarray(csr.data, dtype=np.float32)).to(dev)
        dr = torch.from_numpy(np.ascontiguousarray(d_row, dtype=np.float32)).to(dev)
        dc = torch.from_numpy(np.ascontiguousarray(d_col, dtype=np.float32)).to(dev)
        val = torch.empty_like(w)
        g = cls._from_csr_host(csr, device, values=val)
        lib = _lib.load()
        _lib.check(lib.agcf_norm_adj_csr(g.rowptr.data_ptr(), g.col.data_ptr(), w.data_ptr(), dr.data_ptr(),
                                         dc.data_ptr(), val.data_ptr(), g.n_rows, g.nnz, _lib.stream_ptr()),
                   "agcf_norm_adj_csr")
        g.refresh_plan_values()
        return g

    @classmethod
    def from_ui_adj(cls, ui_adj, device="cuda"):
        """``_init_uiAdj`` (recommender/LightGCN.py:212-215): 1/np.sqrt of row AND
        column sums, no inf guard, fractional weights allowed."""
        d_row = np.array((1 / np.sqrt(ui_adj.sum(1)))).flatten()
        d_col = np.array((1 / np.sqrt(ui_adj.sum(0)))).flatten()
        return cls.normalized(ui_adj, d_row, d_col, device)

    @classmethod
    def from_dataloader_adj(cls, adj, device="cuda"):
        """``DataLoader.normalize_graph_mat`` (util/DataLoader.py:73-81): np.power(rowsum,
        -0.5) with inf -> 0, the same vector on both sides."""
        rowsum = np.array(adj.sum(1))
        d = np.power(rowsum, -0.5).flatten()
        d[np.isinf(d)] = 0.
        return cls.normalized(adj, d, d, device)

    @classmethod
    def from_coo_tensor(cls, t: torch.Tensor):
        """From a torch sparse COO tensor (someone assigned model.sparse_norm_adj)."""
        t = t.detach().coalesce()
        idx = t.indices()
        n = t.shape[0]
        counts = torch.bincount(idx[0], minlength=n)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=t.device)
        rowptr[1:] = torch.cumsum(counts, 0)
        order, n_long = _plan_rows(rowptr.cpu().numpy())
        return cls(rowptr.to(torch.int32), idx[1].to(torch.int32).contiguous(), t.values().float().contiguous(),
                   n, torch.from_numpy(order).to(t.device), n_long)

    # ------------------------------------------------------------------ views
    def coo_indices(self):
        """int64 [2, nnz] COO indices (row-major order) -- built once on demand."""
        if self._coo_idx is None:
            lib = _lib.load()
            row = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
            _lib.check(lib.agcf_csr_expand_rows(self.rowptr.data_ptr(), row.data_ptr(), self.n_rows, self.nnz,
                                                _lib.stream_ptr()), "agcf_csr_expand_rows")
            self._coo_idx = torch.stack([row.long(), self.col.long()])
        return self._coo_idx

    def to_coo_tensor(self):
        """torch sparse COO tensor sharing ``val`` -- the ``sparse_norm_adj`` attribute
        attacks touch (attack/White/PGA.py:98,117).  Unlike the reference's it IS
        coalesced (it is built from canonical CSR)."""
        return torch.sparse_coo_tensor(self.coo_indices(), self.val, (self.n_rows, self.n_rows),
                                       is_coalesced=True)

    def to_scipy(self):
        return sp.csr_matrix((self.val.cpu().numpy(), self.col.cpu().numpy(), self.rowptr.cpu().numpy()),
                             shape=(self.n_rows, self.n_rows))
