"""GPU parity of the fused InfoNCE kernels (csrc/contrast.cu) against the oracle's restatement of
util/loss.py:42-49 (oracle.port.infonce, torch CPU fp32 with autograd) on the same inputs.
Tolerance: 1e-5 relative on the loss, 1e-4 of the gradient scale on the gradients (fp32, different
summation order of an n-term exp sum)."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _views(n, d, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, d, generator=g) * scale
    b = a + 0.3 * torch.randn(n, d, generator=g) * scale        # correlated views, like two perturbed passes
    return a, b


@pytest.mark.parametrize("d", [32, 64, 128, 256])
@pytest.mark.parametrize("n", [1, 5, 63, 64, 65, 300, 2048])
@pytest.mark.parametrize("tau", [0.2, 0.1])
def test_infonce_forward_backward_match_the_reference_expression(n, d, tau):
    from arlib_b200.util.loss import InfoNCE
    if n == 2048 and d == 256 and tau == 0.1:
        pytest.skip("covered by the smaller shapes")
    a, b = _views(n, d, n * 7 + d)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = port.infonce(ar, br, tau)
    (ref * 0.2).backward()                                        # cl_rate of SimGCL / XSimGCL as upstream gradient
    ag, bg = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    got = InfoNCE(ag, bg, tau)
    (got * 0.2).backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    for g_, r_ in ((ag.grad.cpu(), ar.grad), (bg.grad.cpu(), br.grad)):
        scale = float(r_.abs().max()) + 1e-30
        assert float((g_ - r_).abs().max()) <= 1e-4 * scale + 1e-7


def test_infonce_small_norm_rows_and_determinism():
    from arlib_b200.util.loss import InfoNCE
    a, b = _views(777, 64, 3, scale=1e-3)                         # embedding-scale magnitudes (xavier init)
    ref = port.infonce(a.clone(), b.clone(), 0.2)
    outs = []
    for _ in range(2):
        ag, bg = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        loss = InfoNCE(ag, bg, 0.2)
        loss.backward()
        outs.append((loss.detach().clone(), ag.grad.clone(), bg.grad.clone()))
    assert abs(float(outs[0][0]) - float(ref)) <= 1e-5 * abs(float(ref))
    for x, y in zip(outs[0], outs[1]):
        assert torch.equal(x, y)                                  # fixed-order reductions: run-to-run identical


def test_infonce_one_sided_gradient_and_shared_input():
    from arlib_b200.util.loss import InfoNCE
    a, b = _views(200, 64, 11)
    ar = a.clone().requires_grad_(True)
    ref = port.infonce(ar, b, 0.2)
    ref.backward()
    ag = a.to(DEV).requires_grad_(True)
    InfoNCE(ag, b.to(DEV), 0.2).backward()                        # view2 needs no gradient
    assert float((ag.grad.cpu() - ar.grad).abs().max()) <= 1e-4 * float(ar.grad.abs().max())
    # gathered rows of ONE table on both sides (XSimGCL: rec view vs layer_cl view of the same pass)
    t = torch.randn(500, 64)
    idx = torch.unique(torch.randint(0, 500, (300,)))
    mix = torch.randn(idx.numel(), 64)                     # (a second view that is not a rescaling of the first:
    tr = t.clone().requires_grad_(True)                    #  F.normalize would cancel that and leave a ~0 gradient)
    port.infonce(tr[idx], torch.tanh(tr[idx]) + 0.5 * mix, 0.1).backward()
    tg = t.to(DEV).requires_grad_(True)
    InfoNCE(tg[idx.to(DEV)], torch.tanh(tg[idx.to(DEV)]) + 0.5 * mix.to(DEV), 0.1).backward()
    assert float((tg.grad.cpu() - tr.grad).abs().max()) <= 1e-4 * float(tr.grad.abs().max()) + 1e-7


def test_infonce_rejects_unsupported_width():
    from arlib_b200 import _lib
    from arlib_b200.util.loss import InfoNCE
    with pytest.raises(_lib.AgcfError):
        InfoNCE(torch.randn(10, 48, device=DEV), torch.randn(10, 48, device=DEV), 0.2)
