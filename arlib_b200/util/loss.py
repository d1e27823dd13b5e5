"""Differentiable losses with the reference's signatures (util/loss.py:5-9,25-29,42-49).

These are the torch-level entry points attacks import (attack/Black/GTA.py:205,
attack/White/DLAttack.py).  The recommender's own training loop does not come
through here when it owns the optimizer: it runs the fused CUDA kernels
(agcf_bpr_forward / agcf_bpr_backward) inside the engines.  ``bpr_l2_fused``
exposes the same two kernels as ONE autograd function for the reference-shaped
loops (caller's optimizer, gradient export, NGCF): three index gathers, ~12
element-wise kernels and three index_put backward passes become two launches
forward and three backward.
"""
import torch


def bpr_loss(user_emb, pos_item_emb, neg_item_emb):
    """util/loss.py:5-9"""
    pos = (user_emb * pos_item_emb).sum(dim=1)
    neg = (user_emb * neg_item_emb).sum(dim=1)
    return (-torch.log(10e-8 + torch.sigmoid(pos - neg))).mean()


def l2_reg_loss(reg, *args):
    """util/loss.py:25-29 -- un-squared Frobenius norms."""
    total = 0
    for emb in args:
        total = total + torch.norm(emb, p=2)
    return total * reg


class _BprL2(torch.autograd.Function):
    """loss = bpr_loss(U[u], V[i], V[j]) + l2_reg_loss(reg, U[u], V[i]) on the propagated tables
    (recommender/LightGCN.py:51-54): agcf_bpr_forward; backward = agcf_bpr_group_batches + agcf_bpr_backward
    (sorted, atomic-free segment sums: every distinct node's gradient row is written once)."""

    @staticmethod
    def forward(ctx, user_all, item_all, u, i, j, reg):
        from .. import ops
        from ..encoder import pack_table
        F_ = pack_table(user_all.detach(), item_all.detach())          # zero-copy when the two are views of one table
        nb, n_users = int(u.numel()), int(user_all.shape[0])
        dev = F_.device
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        coef = torch.empty(max(nb, 1), dtype=torch.float32, device=dev)
        ws = torch.zeros(ops.bpr_ws_bytes(nb), dtype=torch.uint8, device=dev)      # its ticket word must be 0 on entry
        ops.bpr_forward(F_, u, i, j, nb, n_users, float(reg), out4, coef, ws)
        ctx.save_for_backward(F_, u, i, j, out4, coef)
        ctx.reg, ctx.n_users = float(reg), n_users
        ctx.mark_non_differentiable(out4)
        return out4[0].clone(), out4

    @staticmethod
    def backward(ctx, grad, _grad_parts):
        from .. import ops
        F_, u, i, j, out4, coef = ctx.saved_tensors
        nb, n_users, dev = int(u.numel()), ctx.n_users, F_.device
        i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        occ, seg_off, seg_node, n_seg = i32(3 * nb), i32(3 * nb + 1), i32(3 * nb), i32(1)
        ops.bpr_group_batches(u, i, j, nb, nb, n_users, occ, seg_off, seg_node, n_seg)
        G = torch.zeros_like(F_)
        ops.bpr_backward(F_, u, i, j, nb, n_users, ctx.reg, 1.0, out4, coef, occ, seg_off, seg_node, n_seg, G)
        G.mul_(grad.to(torch.float32))
        return G[:n_users], G[n_users:], None, None, None, None


def _ids_i32(idx, dev):
    if torch.is_tensor(idx):
        return idx.to(device=dev, dtype=torch.int32).contiguous()
    return torch.as_tensor(idx, dtype=torch.int32).to(dev)


def bpr_l2_fused(rec_user_emb, rec_item_emb, user_idx, pos_idx, neg_idx, reg, return_parts=False):
    """``bpr_loss(U[u], V[i], V[j]) + l2_reg_loss(reg, U[u], V[i])`` -- the loss line of every graph recommender's
    train() (recommender/LightGCN.py:51-54, NGCF.py:53-56, SimGCL.py:50-54, XSimGCL.py:60-64) -- as one fused CUDA
    op, differentiable w.r.t. both tables.  ``return_parts``: also the detached 4-vector {total, bpr term, |U[u]|_F,
    |V[i]|_F}.  Batches the single-CTA grouping cannot sort (3 B > 16384) use the torch expressions (still on the GPU)."""
    dev = rec_user_emb.device
    nb = len(user_idx)
    fits = rec_user_emb.is_cuda and rec_user_emb.dtype == torch.float32 and 0 < 3 * nb <= 16384 \
        and rec_user_emb.shape[1] in (32, 64, 128, 256)
    if not fits:
        if not rec_user_emb.is_cuda:
            raise RuntimeError("arlib_b200 losses run on CUDA tensors only (no CPU fallback)")
        u, i, j = (torch.as_tensor(x, dtype=torch.long, device=dev) if not torch.is_tensor(x) else x.long() for x in
                   (user_idx, pos_idx, neg_idx))
        ue, pe, ne = rec_user_emb[u], rec_item_emb[i], rec_item_emb[j]
        rec = bpr_loss(ue, pe, ne)
        total = rec + l2_reg_loss(reg, ue, pe)
        if return_parts:
            parts = torch.stack([total.detach(), rec.detach(), torch.norm(ue.detach()), torch.norm(pe.detach())])
            return total, parts
        return total
    total, parts = _BprL2.apply(rec_user_emb, rec_item_emb, _ids_i32(user_idx, dev), _ids_i32(pos_idx, dev),
                                _ids_i32(neg_idx, dev), float(reg))
    return (total, parts) if return_parts else total


class _InfoNCE(torch.autograd.Function):
    """agcf_infonce_forward / agcf_infonce_backward as one differentiable op."""

    @staticmethod
    def forward(ctx, view1, view2, temperature):
        from .. import ops
        loss, ws = ops.infonce_forward(view1, view2, temperature)
        ctx.ws, ctx.shape, ctx.temperature = ws, tuple(view1.shape), float(temperature)
        return loss[0]

    @staticmethod
    def backward(ctx, grad):
        from .. import ops
        n, d = ctx.shape
        g = grad.detach().to(torch.float32).reshape(1).contiguous()
        g1 = torch.empty((n, d), dtype=torch.float32, device=g.device) if ctx.needs_input_grad[0] else None
        g2 = torch.empty((n, d), dtype=torch.float32, device=g.device) if ctx.needs_input_grad[1] else None
        if g1 is not None or g2 is not None:
            ops.infonce_backward(n, d, ctx.temperature, ctx.ws, grad_loss=g, grad1=g1, grad2=g2)
        return g1, g2, None


def InfoNCE(view1, view2, temperature):
    """util/loss.py:42-49 on the fused kernels of csrc/contrast.cu -- the n x n logit matrix never reaches memory.
    CUDA tensors only: like every other entry point of this package there is no CPU path (callers that keep tensors on
    the host get an error, not a silent torch fallback)."""
    if not (view1.is_cuda and view2.is_cuda):
        raise RuntimeError("arlib_b200.util.loss.InfoNCE needs CUDA tensors (no CPU fallback)")
    if view1.shape[0] == 0:                      # mean over zero rows: nan like the reference expression
        return (view1.sum() + view2.sum()) * float('nan')
    v1 = view1.to(torch.float32).contiguous()
    v2 = view2.to(torch.float32).contiguous()
    return _InfoNCE.apply(v1, v2, float(temperature))
