"""Graph-CF encoders on the agcf propagation kernels, as torch autograd functions.

Reference: recommender/LightGCN.py:202-240 (LGCN_Encoder), SimGCL.py:169-219,
XSimGCL.py:179-223, NGCF.py:163-212.  The modules keep the reference's attribute
surface (``embedding_dict`` ParameterDict with 'user_emb' / 'item_emb',
``sparse_norm_adj``, ``_init_uiAdj``, ``attack_emb``, ``forward``) so attacks can
drive them unchanged; the arithmetic underneath is agcf_spmm_csr_f32 /
agcf_sddmm_csr_f32.  Because the normalized adjacency is symmetric the backward
pass is the same SpMM kernel on the same CSR (no transposed copy).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .graph import DeviceGraph


# ------------------------------------------------------------------ table packing
def adjacent_table(a: torch.Tensor, b: torch.Tensor):
    """If ``a`` [U,d] and ``b`` [I,d] are contiguous, back-to-back views of one
    storage, return the [U+I,d] tensor over that memory (zero-copy), else None."""
    if a is None or b is None or not (a.is_contiguous() and b.is_contiguous()):
        return None
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1] or a.dtype != b.dtype:
        return None
    if a.untyped_storage().data_ptr() != b.untyped_storage().data_ptr():
        return None
    if b.storage_offset() != a.storage_offset() + a.numel():
        return None
    out = torch.empty(0, dtype=a.dtype, device=a.device)
    out.set_(a.untyped_storage(), a.storage_offset(), (a.shape[0] + b.shape[0], a.shape[1]), (a.shape[1], 1))
    return out


def pack_table(a: torch.Tensor, b: torch.Tensor):
    """[a; b] as one table: zero-copy when adjacent, else one concat kernel."""
    t = adjacent_table(a, b)
    if t is not None:
        return t
    return ops.concat_rows(a.contiguous(), b.contiguous())


# ------------------------------------------------------------------ propagation
class _Propagate(torch.autograd.Function):
    """F = mean over layers of A^k E0 (LightGCN, include_layer0=True) or of the
    perturbed layers 1..L (SimGCL / XSimGCL, include_layer0=False).

    inputs : user_emb, item_emb, adj (sparse COO leaf or None -- only consulted for
             requires_grad), enc (module: graph, n_layers), noises (list or None),
             eps, include_layer0, cl_layer (0 = no extra view)
    outputs: user_all, item_all [, user_cl, item_cl]
    """

    @staticmethod
    def forward(ctx, user_emb, item_emb, adj, graph, n_layers, noises, eps, include_layer0, cl_layer):
        g: DeviceGraph = graph
        nu, ni = user_emb.shape[0], item_emb.shape[0]
        if nu + ni != g.n_rows:
            raise RuntimeError("embedding rows (%d+%d) do not match the adjacency (%d); call _init_uiAdj after "
                               "resizing the model" % (nu, ni, g.n_rows))
        e0 = pack_table(user_emb.detach(), item_emb.detach())
        need_adj_grad = adj is not None and adj.requires_grad and ctx.needs_input_grad[2]
        L = n_layers
        n_mean = L + 1 if include_layer0 else L
        acc = torch.empty_like(e0)
        layers = [e0] if need_adj_grad else None
        x = e0
        cl = e0
        for k in range(1, L + 1):
            last = k == L
            keep_y = (not last) or need_adj_grad or cl_layer == k
            y = torch.empty_like(e0) if keep_y else None
            if k == 1:
                acc_in = e0 if include_layer0 else None
            else:
                acc_in = acc
            ops.spmm(g, x, Y=y, acc_in=acc_in, acc_out=acc, acc_div=float(n_mean) if last else 1.0,
                     noise=None if noises is None else noises[k - 1], eps=eps)
            if cl_layer == k:
                cl = y
            if not last:
                x = y
                if need_adj_grad:
                    layers.append(y)
        ctx.graph, ctx.L, ctx.n_mean, ctx.nu = g, L, n_mean, nu
        ctx.include_layer0, ctx.cl_layer, ctx.need_adj_grad = include_layer0, cl_layer, need_adj_grad
        ctx.layers = layers
        outs = (acc[:nu], acc[nu:])
        if cl_layer:
            outs = outs + (cl[:nu], cl[nu:])
        return outs

    @staticmethod
    def backward(ctx, *grads):
        g, L, nu = ctx.graph, ctx.L, ctx.nu
        n = g.n_rows
        gu, gi = grads[0], grads[1]
        d = (gu if gu is not None else gi if gi is not None else grads[2] if grads[2] is not None else grads[3]).shape[1]
        dev = g.device

        def table(a, b):
            if a is None and b is None:
                return None
            if a is None:
                a = torch.zeros((nu, d), dtype=torch.float32, device=dev)
            if b is None:
                b = torch.zeros((n - nu, d), dtype=torch.float32, device=dev)
            return pack_table(a, b)

        G = table(gu, gi)                                   # d loss / d F (mean output), unscaled
        Gcl = table(grads[2], grads[3]) if ctx.cl_layer else None
        if G is None:
            G = torch.zeros((n, d), dtype=torch.float32, device=dev)
        # H_k = d loss / d E_k * n_mean.  H_L = G; H_{k-1} = A H_k + G (k-1 >= 1, or k-1 = 0 with layer 0 in
        # the mean); the cl view adds n_mean * Gcl at its layer.  dE0 = H_0 / n_mean.
        scale = float(ctx.n_mean)
        gval = None
        if ctx.need_adj_grad:
            gval = torch.zeros(g.nnz, dtype=torch.float32, device=dev)
        H = G
        if Gcl is not None and ctx.cl_layer == L:
            H = G + Gcl * scale
        dE0 = torch.empty_like(G)
        for k in range(L, 0, -1):
            if gval is not None:
                ops.sddmm(g, H, ctx.layers[k - 1], gval, accumulate=True)
            if k > 1:
                nxt = torch.empty_like(G)
                ops.spmm(g, H, Y=nxt, addend=G)
                if Gcl is not None and ctx.cl_layer == k - 1:
                    nxt += Gcl * scale
                H = nxt
            else:
                ops.spmm(g, H, acc_in=G if ctx.include_layer0 else None, acc_out=dE0, acc_div=scale)
        if Gcl is not None and ctx.cl_layer == 0:
            dE0 += Gcl
        adj_grad = None
        if gval is not None:
            gval /= scale
            adj_grad = torch.sparse_coo_tensor(g.coo_indices(), gval, (n, n), is_coalesced=True)
        return dE0[:nu], dE0[nu:], adj_grad, None, None, None, None, None, None


class _SpMM(torch.autograd.Function):
    """Y = A X as a differentiable op (NGCF's two propagations per layer).  ``adj`` is the encoder's sparse COO leaf
    (or None): when it requires grad the backward also returns dL/dA on the stored pattern -- <gY[i], X[j]> per
    non-zero (agcf_sddmm_csr_f32), what torch.sparse.mm's autograd gives attack/White/PGA.py:98,117."""

    @staticmethod
    def forward(ctx, x, adj, graph):
        ctx.graph = graph
        ctx.need_adj_grad = adj is not None and adj.requires_grad and ctx.needs_input_grad[1]
        x = x.detach().contiguous()
        if ctx.need_adj_grad:
            ctx.save_for_backward(x)
        y = torch.empty_like(x)
        ops.spmm(graph, x, Y=y)
        return y

    @staticmethod
    def backward(ctx, gy):
        g = ctx.graph
        gy = gy.contiguous()
        gx = torch.empty_like(gy)
        ops.spmm(g, gy, Y=gx)
        adj_grad = None
        if ctx.need_adj_grad:
            (x,) = ctx.saved_tensors
            gval = torch.empty(g.nnz, dtype=torch.float32, device=gy.device)
            ops.sddmm(g, gy, x, gval)
            adj_grad = torch.sparse_coo_tensor(g.coo_indices(), gval, (g.n_rows, g.n_rows), is_coalesced=True)
        return gx, adj_grad, None


def spmm_autograd(graph, x, adj=None):
    return _SpMM.apply(x, adj, graph)


# --------------------------------------------------------------------- modules
def unique_ids_like_reference(ids, dev):
    """``torch.unique(torch.Tensor(ids).type(torch.long))`` of recommender/SimGCL.py:213-214 / XSimGCL.py:40-41: the
    ids pass through float32 (exact below 2**24).  Accepts the reference's Python lists or the
    device LongTensors of the Philox sampler."""
    if torch.is_tensor(ids):
        return torch.unique(ids.to(dev).to(torch.float32).to(torch.long))
    return torch.unique(torch.Tensor(ids).type(torch.long)).to(dev)


class TorchGraphInterface(object):
    """recommender/LightGCN.py:243-252 -- kept for callers that import it."""

    @staticmethod
    def convert_sparse_mat_to_tensor(X):
        return DeviceGraph.from_scipy(X).to_coo_tensor()


class GraphEncoderBase(nn.Module):
    """Shared plumbing of the four encoders: parameter table, device graph and the
    ``sparse_norm_adj`` attribute."""

    def __init__(self, data, emb_size):
        super().__init__()
        self.data = data
        self.latent_size = emb_size
        self.emb_size = emb_size
        self.norm_adj = data.norm_adj
        self._graph = self._graph_from_data(data)
        self._adj_tensor = None
        self.embedding_dict = self._init_model()

    # graph -----------------------------------------------------------------
    @staticmethod
    def _device():
        if not torch.cuda.is_available():
            raise RuntimeError("arlib_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def _graph_from_data(self, data):
        dev = self._device()
        ui_adj = getattr(data, "ui_adj", None)
        if ui_adj is not None and ui_adj.shape == data.norm_adj.shape and ui_adj.dtype == 'float32':
            # recompute the values on device from the raw weights (bit-identical to data.norm_adj)
            return DeviceGraph.from_dataloader_adj(ui_adj, dev)
        return DeviceGraph.from_scipy(data.norm_adj, dev)

    @property
    def sparse_norm_adj(self):
        """torch sparse COO view of the device graph (recommender/LightGCN.py:210);
        materialized on first access, cached so that ``.requires_grad = True`` sticks."""
        if self._adj_tensor is None:
            self._adj_tensor = self._graph.to_coo_tensor()
        return self._adj_tensor

    @sparse_norm_adj.setter
    def sparse_norm_adj(self, value):
        self._graph = DeviceGraph.from_coo_tensor(value.to(self._device()))
        self._adj_tensor = value

    def _init_uiAdj(self, ui_adj):
        """recommender/LightGCN.py:212-215: re-normalize an (N' x N') symmetric scipy
        adjacency (fractional weights allowed) and make it the propagation matrix."""
        if ui_adj.dtype != 'float32':
            ui_adj = ui_adj.astype('float32')
        # same pattern as the previous call (PGA's per-batch loop): the device indices, the work plan and the COO index
        # view are re-used, only the weights are re-normalized (DeviceGraph.normalized)
        self._graph = DeviceGraph.from_ui_adj(ui_adj, self._device(), reuse=getattr(self, "_uiadj_graph", None))
        self._uiadj_graph = self._graph
        self._adj_tensor = None

    # parameters --------------------------------------------------------------
    def _init_model(self):
        """xavier_uniform on the HOST generator, user table first then item table --
        the same draws as recommender/LightGCN.py:222-228 under torch.manual_seed --
        stored as two views of ONE [N,d] device table so propagation needs no concat."""
        init = nn.init.xavier_uniform_
        u = init(torch.empty(self.data.user_num, self.latent_size))
        i = init(torch.empty(self.data.item_num, self.latent_size))
        table = torch.cat([u, i], 0).to(self._device())
        return nn.ParameterDict({
            'user_emb': nn.Parameter(table[:self.data.user_num]),
            'item_emb': nn.Parameter(table[self.data.user_num:]),
        })

    def parameter_table(self):
        """The [N,d] tensor both embedding parameters are views of; re-packs them
        into one fresh table first if some caller broke the adjacency (deepcopy,
        re-assignment), keeping the Parameter objects (and optimizers on them) valid."""
        pu, pi = self.embedding_dict['user_emb'], self.embedding_dict['item_emb']
        t = adjacent_table(pu.data, pi.data)
        if t is None:
            t = ops.concat_rows(pu.data.contiguous(), pi.data.contiguous())
            pu.data = t[:pu.shape[0]]
            pi.data = t[pu.shape[0]:]
        return t

    def attack_emb(self, users_emb_grad, items_emb_grad):
        """recommender/LightGCN.py:217-220"""
        with torch.no_grad():
            self.embedding_dict['user_emb'] += users_emb_grad
            self.embedding_dict['item_emb'] += items_emb_grad

    def _adj_for_autograd(self):
        a = self._adj_tensor
        return a if (a is not None and a.requires_grad) else None


class LGCN_Encoder(GraphEncoderBase):
    """recommender/LightGCN.py:202-240"""

    def __init__(self, data, emb_size, n_layers):
        super().__init__(data, emb_size)
        self.layers = n_layers

    def forward(self):
        return _Propagate.apply(self.embedding_dict['user_emb'], self.embedding_dict['item_emb'],
                                self._adj_for_autograd(), self._graph, self.layers, None, 0.0, True, 0)


class SimGCL_Encoder(GraphEncoderBase):
    """recommender/SimGCL.py:169-219"""

    def __init__(self, data, emb_size, eps, n_layers):
        super().__init__(data, emb_size)
        self.eps = eps
        self.n_layers = n_layers
        self.noise_source = None          # test hook: callable(k, like) -> U[0,1) noise tensor

    def _noises(self, like):
        n = self._graph.n_rows
        out = []
        for k in range(self.n_layers):
            if self.noise_source is not None:
                out.append(self.noise_source(k, like).contiguous())
            else:
                out.append(torch.rand((n, self.emb_size), dtype=torch.float32, device=like.device))
        return out

    def forward(self, perturbed=False):
        pu, pi = self.embedding_dict['user_emb'], self.embedding_dict['item_emb']
        noises = self._noises(pu) if perturbed else None
        return _Propagate.apply(pu, pi, self._adj_for_autograd(), self._graph, self.n_layers, noises,
                                float(self.eps), False, 0)

    def cal_cl_loss(self, idx):
        """recommender/SimGCL.py:212-219 (ids go through float32 exactly like
        torch.Tensor(list) there)."""
        from .util.loss import InfoNCE
        dev = self.embedding_dict['user_emb'].device
        u_idx, i_idx = unique_ids_like_reference(idx[0], dev), unique_ids_like_reference(idx[1], dev)
        u1, i1 = self.forward(perturbed=True)
        u2, i2 = self.forward(perturbed=True)
        return InfoNCE(u1[u_idx], u2[u_idx], 0.2) + InfoNCE(i1[i_idx], i2[i_idx], 0.2)


class XSimGCL_Encoder(GraphEncoderBase):
    """recommender/XSimGCL.py:179-223"""

    def __init__(self, data, emb_size, eps, n_layers, layer_cl):
        super().__init__(data, emb_size)
        self.eps = eps
        self.n_layers = n_layers
        self.layer_cl = layer_cl
        self.noise_source = None

    _noises = SimGCL_Encoder._noises

    def forward(self, perturbed=False):
        pu, pi = self.embedding_dict['user_emb'], self.embedding_dict['item_emb']
        if not perturbed:
            return _Propagate.apply(pu, pi, self._adj_for_autograd(), self._graph, self.n_layers, None, 0.0, False, 0)
        cl_layer = self.layer_cl if 1 <= self.layer_cl <= self.n_layers else 0
        outs = _Propagate.apply(pu, pi, self._adj_for_autograd(), self._graph, self.n_layers, self._noises(pu),
                                float(self.eps), False, cl_layer)
        if cl_layer == 0:       # layer_cl outside 1..L: the reference's cl view stays the ego embeddings
            return outs[0], outs[1], pu, pi
        return outs


class NGCF_Encoder(GraphEncoderBase):
    """recommender/NGCF.py:163-212.  ``forward`` (the differentiable model() attacks call, and the reference-shaped
    training loop) runs the reference's expression on torch autograd with the two propagations per layer on the agcf
    SpMM; ``train()`` of the recommender, when it owns the optimizer, runs the fused NGCFEngine instead (one propagation
    per layer + the fused dense kernels of csrc/ngcf.cu).  The 2L weight matrices are views of ONE packed [L, 2d, d]
    buffer (row block k = [W1_k ; W2_k]) -- what the fused layer multiplies with -- like the two embedding tables are
    views of one [N, d] table."""

    def __init__(self, data, emb_size, n_layers):
        self.layers = n_layers
        super().__init__(data, emb_size)

    def _init_model(self):
        emb = super()._init_model()
        init = nn.init.xavier_uniform_
        d = self.latent_size
        host = []
        for k in range(self.layers):              # host-generator draws in the reference's order: w1_k, then w2_k
            host.append(torch.cat([init(torch.empty(d, d)), init(torch.empty(d, d))], 0))
        packed = (torch.stack(host, 0) if host else torch.empty(0, 2 * d, d)).to(self._device())
        w = {}
        for k in range(self.layers):
            w['w1_' + str(k)] = nn.Parameter(packed[k, :d])
            w['w2_' + str(k)] = nn.Parameter(packed[k, d:])
        self.W = nn.ParameterDict(w)
        return emb

    def weight_table(self):
        """The packed [L, 2d, d] tensor the 2L weight parameters are views of; re-packs them into a fresh buffer first
        if some caller broke the sharing (deepcopy, re-assignment), keeping the Parameter objects valid."""
        d, L = self.latent_size, self.layers
        ps = [self.W['w%d_%d' % (a, k)] for k in range(L) for a in (1, 2)]
        if L == 0:
            return torch.empty((0, 2 * d, d), dtype=torch.float32, device=self._device())
        base = ps[0].data
        ok = all(p.data.is_contiguous() and p.data.untyped_storage().data_ptr() == base.untyped_storage().data_ptr()
                 and p.data.storage_offset() == base.storage_offset() + q * d * d for q, p in enumerate(ps))
        if ok:
            out = torch.empty(0, dtype=base.dtype, device=base.device)
            out.set_(base.untyped_storage(), base.storage_offset(), (L, 2 * d, d), (2 * d * d, d, 1))
            return out
        packed = torch.stack([torch.cat([self.W['w1_%d' % k].data, self.W['w2_%d' % k].data], 0) for k in range(L)], 0).contiguous()
        for k in range(L):
            self.W['w1_%d' % k].data = packed[k, :d]
            self.W['w2_%d' % k].data = packed[k, d:]
        return packed

    def forward(self):
        import torch.nn.functional as F
        pu, pi = self.embedding_dict['user_emb'], self.embedding_dict['item_emb']
        ego = torch.cat([pu, pi], 0)
        adj = self._adj_for_autograd()
        layers = [ego]
        for k in range(self.layers):
            t = torch.mm(ego, self.W['w1_' + str(k)])
            ego = F.leaky_relu(spmm_autograd(self._graph, t, adj) + t +
                               torch.mm(spmm_autograd(self._graph, ego, adj) * ego, self.W['w2_' + str(k)]))
            layers.append(ego)
        out = torch.mean(torch.stack(layers, dim=1), dim=1)
        return out[:self.data.user_num], out[self.data.user_num:]
