"""Differentiable losses with the reference's signatures (util/loss.py:5-9,25-29,42-49).

These are the torch-level entry points attacks import (attack/Black/GTA.py:205,
attack/White/DLAttack.py).  The recommender's own training loop does not come
through here when it owns the optimizer: it runs the fused CUDA kernels
(agcf_bpr_forward / agcf_bpr_backward).  ``bpr_l2_fused`` exposes those kernels
as one autograd function for callers that hand in their own optimizer.
"""
import torch
import torch.nn.functional as F


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
    """util/loss.py:42-49.  CUDA fp32 views (what SimGCL / XSimGCL.cal_cl_loss pass) run the fused kernels of
    csrc/contrast.cu -- the n x n logit matrix never reaches memory -- and raise if the library is missing; tensors a
    caller keeps on the CPU are evaluated with the reference's own torch expression."""
    if view1.is_cuda and view1.shape[0] > 0:
        v1 = view1.to(torch.float32).contiguous()
        v2 = view2.to(torch.float32).contiguous()
        return _InfoNCE.apply(v1, v2, float(temperature))
    view1, view2 = F.normalize(view1, dim=1), F.normalize(view2, dim=1)
    pos = torch.exp((view1 * view2).sum(dim=-1) / temperature)
    ttl = torch.exp(torch.matmul(view1, view2.transpose(0, 1)) / temperature).sum(dim=1)
    return (-torch.log(pos / ttl)).mean()
