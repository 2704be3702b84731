"""GPU parity of the fused tail kernels (through the C ABI) against the oracle and the
reference's golden vectors.  Tolerances: fp32 kernels vs fp32 oracle, rtol 1e-5 / atol 1e-6 on
probabilities, coordinates and losses; gradients rtol 1e-4 / atol 1e-7 (log/exp rounding)."""
import os

import pytest
import torch

from oracle import dsnt_oracle as D
from tests.golden.make_golden import tail_inputs

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'margipose_golden.pt'),
                  weights_only=False)
TOL = dict(rtol=1e-5, atol=1e-6)
GTOL = dict(rtol=1e-4, atol=1e-7)


def _cuda(ts):
    return [t.detach().cuda() for t in ts]


def test_known_answer_reference_test():
    # /root/reference/tests/test_models.py:39-46 through the CUDA path
    from margipose_b200 import dsntnn as K
    from margipose_b200.models.margipose_model import MargiPoseModel
    size, sigma = (32, 32), 1
    xy = K.make_gauss(torch.Tensor([[[-0.5, 0.5]]]).cuda(), size, sigma, normalize=True)
    zy = K.make_gauss(torch.Tensor([[[0.1, 0]]]).cuda(), size, sigma, normalize=True)
    xz = K.make_gauss(torch.Tensor([[[0, 0.2]]]).cuda(), size, sigma, normalize=True)
    xyz = MargiPoseModel.heatmaps_to_coords(xy, zy, xz)
    torch.testing.assert_close(xyz.cpu(), torch.Tensor([[[-0.5, 0.5, 0.15]]]))
    k = GOLD['known_answer']
    torch.testing.assert_close(xy.cpu(), k['xy'], **TOL)
    torch.testing.assert_close(xyz.cpu(), k['coords'], **TOL)


@pytest.mark.parametrize('case', GOLD['tail'], ids=lambda c: 'x'.join(map(str, c['shape'])))
def test_generic_ops_vs_golden_and_oracle(case):
    from margipose_b200 import dsntnn as K
    z, target, mask = tail_inputs(case['seed'], case['shape'], case['scale'])
    zc = [t.cuda().requires_grad_() for t in z]
    tc, mc = target.cuda(), mask.cuda()
    p = [K.flat_softmax(t) for t in zc]
    t_pl = [tc[..., [0, 1]].contiguous(), tc[..., [2, 1]].contiguous(), tc[..., [0, 2]].contiguous()]
    for k in range(3):
        torch.testing.assert_close(K.dsnt(p[k]).cpu(), case['dsnt'][k], **TOL)
        torch.testing.assert_close(p[k].sum(-1).cpu(), case['prob_rowsum'][k], **TOL)
        torch.testing.assert_close(p[k].sum(-2).cpu(), case['prob_colsum'][k], **TOL)
        torch.testing.assert_close(K.js_reg_losses(p[k], t_pl[k], 1.0).cpu(), case['js'][k], **TOL)
        if case['probs'] is not None:
            torch.testing.assert_close(p[k].cpu(), case['probs'][k], **TOL)
    coords = K.heatmaps_to_coords(*p)
    torch.testing.assert_close(coords.cpu(), case['coords'], **TOL)
    eu3 = K.euclidean_losses(coords, tc)
    torch.testing.assert_close(eu3.cpu(), case['eu3'], **TOL)
    l3 = K.js_reg_losses(p[0], t_pl[0], 1.0) + K.js_reg_losses(p[1], t_pl[1], 1.0) + \
        K.js_reg_losses(p[2], t_pl[2], 1.0) + eu3
    loss3 = K.average_loss(l3, mc)
    torch.testing.assert_close(loss3.cpu(), case['loss3'], **TOL)
    g3 = torch.autograd.grad(loss3, zc)
    for k in range(3):
        torch.testing.assert_close(g3[k].sum(-1).cpu(), case['grad3_rowsum'][k], rtol=1e-4, atol=1e-6)
        if case['grad3'] is not None:
            torch.testing.assert_close(g3[k].cpu(), case['grad3'][k], **GTOL)
    torch.testing.assert_close(K.make_gauss(t_pl[0], case['shape'][-2:], 1.0).sum(-1).cpu(),
                               case['gauss_rowsum'], **TOL)


@pytest.mark.parametrize('case', GOLD['tail'], ids=lambda c: 'x'.join(map(str, c['shape'])))
@pytest.mark.parametrize('from_logits', [True, False])
def test_fused_tail_vs_golden(case, from_logits):
    from margipose_b200 import dsntnn as K
    z, target, mask = tail_inputs(case['seed'], case['shape'], case['scale'])
    zc = [t.cuda().requires_grad_() for t in z]
    tc, mc = target.cuda(), mask.cuda()
    if from_logits:
        pxy, pzy, pxz, coords, l3 = K.fused_tail_from_logits(zc[0], zc[1], zc[2], tc)
    else:
        pxy, pzy, pxz = [K.flat_softmax(t) for t in zc]
        l3, coords = K.fused_tail_losses(pxy, pzy, pxz, tc)
    torch.testing.assert_close(coords.cpu(), case['coords'], **TOL)
    want = case['js'][0] + case['js'][1] + case['js'][2] + case['eu3']
    torch.testing.assert_close(l3.cpu(), want, **TOL)
    loss3 = K.average_loss(l3, mc)
    torch.testing.assert_close(loss3.cpu(), case['loss3'], **TOL)
    g3 = torch.autograd.grad(loss3, zc)
    for k in range(3):
        torch.testing.assert_close(g3[k].sum(-1).cpu(), case['grad3_rowsum'][k], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(g3[k].abs().mean().cpu(), case['grad3_absmean'][k], rtol=1e-4, atol=1e-9)
        if case['grad3'] is not None:
            torch.testing.assert_close(g3[k].cpu(), case['grad3'][k], **GTOL)


@pytest.mark.parametrize('shape', [(4, 17, 32, 32), (2, 17, 48, 48), (2, 3, 128, 128), (2, 3, 64, 128), (2, 5, 64, 64), (1, 2, 32, 64),
                                   (2, 4, 20, 30), (3, 2, 7, 9)])
def test_fused_tail_mixed_2d_3d_and_upstream_grads_vs_oracle(shape):
    """Mixed valid_depth batches (bin/train_3d.py:126-142), upstream heatmap gradients (what the
    next stage's combiner sends back), ragged / non-multiple-of-4 planes, pixelwise_loss=None."""
    from margipose_b200 import dsntnn as K
    g = torch.Generator().manual_seed(7)
    B, J, H, W = shape
    z = [torch.randn(shape, generator=g) * 2 for _ in range(3)]
    target = torch.rand(B, J, 3, generator=g) * 1.6 - 0.8
    mask = (torch.rand(B, J, generator=g) > 0.3).float()
    vd = torch.tensor([(i % 2) for i in range(B)], dtype=torch.int32)
    up = [torch.randn(shape, generator=g) * 1e-3 for _ in range(3)]
    for pixelwise in (True, False):
        zo = [t.clone().requires_grad_() for t in z]
        po = [D.flat_softmax(t) for t in zo]
        lo = D.forward_loss([po[0]], [po[1]], [po[2]], target, mask, vd.tolist(),
                            'jsd' if pixelwise else None)
        extra = sum((p * u).sum() for p, u in zip(po, up))
        go = torch.autograd.grad(lo + extra, zo)
        zc = [t.cuda().requires_grad_() for t in z]
        pxy, pzy, pxz, coords, l = K.fused_tail_from_logits(zc[0], zc[1], zc[2], target.cuda(),
                                                           valid_depth=vd.cuda(), pixelwise=pixelwise)
        lc = K.average_loss(l, mask.cuda())
        extra_c = sum((p * u.cuda()).sum() for p, u in zip((pxy, pzy, pxz), up))
        gc = torch.autograd.grad(lc + extra_c, zc)
        torch.testing.assert_close(lc.cpu(), lo.detach(), **TOL)
        torch.testing.assert_close(coords.cpu(), D.heatmaps_to_coords(*po).detach(), **TOL)
        for k in range(3):
            torch.testing.assert_close(pxy.cpu() if k == 0 else (pzy.cpu() if k == 1 else pxz.cpu()),
                                       po[k].detach(), **TOL)
            torch.testing.assert_close(gc[k].cpu(), go[k], rtol=1e-4, atol=2e-7)


def test_average_loss_edge_cases():
    from margipose_b200 import dsntnn as K
    l = torch.rand(3, 17).cuda()
    zero = torch.zeros(3, 17).cuda()
    assert K.average_loss(l, zero).item() == 0.0            # denominator clamps to 1
    torch.testing.assert_close(K.average_loss(l).cpu(), l.mean().cpu(), **TOL)
    with pytest.raises(AssertionError):
        K.average_loss(l, torch.ones(3, 16).cuda())


def test_cpu_tensors_are_rejected():
    from margipose_b200 import dsntnn as K
    from margipose_b200._lib import MargiposeB200Error
    with pytest.raises(MargiposeB200Error):
        K.flat_softmax(torch.randn(1, 17, 32, 32))
