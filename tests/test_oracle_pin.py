"""Pins oracle/ against the UNMODIFIED reference, where /root/reference exists.

(The GPU box has no /root/reference: these tests skip there and the committed golden
vectors in tests/golden/ -- generated from the reference by make_golden.py -- take over.)
"""
import pytest
import torch

from oracle import dsnt_oracle as D
from oracle import model_oracle as M
from oracle.ref_shim import reference_available, load_reference

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference tree not present')


@pytest.fixture(scope='module')
def ref():
    return load_reference()


def test_known_answer_from_reference_tests(ref):
    # /root/reference/tests/test_models.py:39-46, run against BOTH implementations
    for mk, h2c in [(ref.dsntnn.make_gauss, ref.model.MargiPoseModel.heatmaps_to_coords),
                    (D.make_gauss, D.heatmaps_to_coords)]:
        xy = mk(torch.Tensor([[[-0.5, 0.5]]]), (32, 32), 1, normalize=True)
        zy = mk(torch.Tensor([[[0.1, 0]]]), (32, 32), 1, normalize=True)
        xz = mk(torch.Tensor([[[0, 0.2]]]), (32, 32), 1, normalize=True)
        torch.testing.assert_close(h2c(xy, zy, xz), torch.Tensor([[[-0.5, 0.5, 0.15]]]))


@pytest.mark.parametrize('size', [(32, 32), (16, 48), (64, 64)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_tail_functions_match_reference(ref, size, dtype):
    g = torch.Generator().manual_seed(1)
    z = (torch.randn(3, 17, *size, generator=g, dtype=dtype) * 3).requires_grad_()
    mu = (torch.rand(3, 17, 2, generator=g, dtype=dtype) * 1.6 - 0.8)
    outs = []
    for mod in (ref.dsntnn, D):
        p = mod.flat_softmax(z)
        c = mod.dsnt(p)
        js = mod.js_reg_losses(p, mu, 1.0)
        eu = mod.euclidean_losses(c, mu)
        mask = (torch.arange(3 * 17).reshape(3, 17) % 3 != 0).to(dtype)
        loss = mod.average_loss(js + eu, mask)
        (gz,) = torch.autograd.grad(loss, z)
        gauss = mod.make_gauss(mu, size, 1.0)
        outs.append((p, c, js, eu, loss, gz, gauss))
    tol = dict(rtol=1e-5, atol=1e-7) if dtype == torch.float32 else dict(rtol=1e-11, atol=1e-14)
    for a, b in zip(*outs):
        torch.testing.assert_close(b, a, **tol)


def test_joint_bookkeeping_bit_exact(ref):
    sk = ref.CanonicalSkeletonDesc
    assert list(sk.joint_names) == M.JOINT_NAMES
    assert list(sk.joint_tree) == M.JOINT_TREE
    assert list(sk.hflip_indices) == M.HFLIP_INDICES
    assert sk.n_joints == M.N_JOINTS == 17


@pytest.mark.parametrize('fe,n_stages', [('resnet18', 2), ('resnet34', 1), ('resnet50', 1)])
def test_model_matches_reference(ref, fe, n_stages):
    desc = {'type': 'margipose', 'version': '6.0.1',
            'settings': {'n_stages': n_stages, 'axis_permutation': True,
                         'feature_extractor': fe, 'pixelwise_loss': 'jsd'}}
    torch.manual_seed(3)
    rm = ref.models.create_model(desc)
    om = M.create_oracle(desc)
    assert list(om.state_dict().keys()) == list(rm.state_dict().keys())
    for k, v in rm.state_dict().items():
        assert om.state_dict()[k].shape == v.shape, k
    om.load_state_dict(rm.state_dict())
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 256, 256, generator=g)
    t = torch.rand(2, 17, 3, generator=g) * 1.6 - 0.8
    mask = torch.ones(2, 17)
    res = []
    for m, avg in ((rm, ref.dsntnn.average_loss), (om, D.average_loss)):
        m.train()
        out = m(x)
        l3 = avg(m.forward_3d_losses(out, t), mask)
        l2 = avg(m.forward_2d_losses(out, t), mask)
        (l3 + l2).backward()
        res.append((out, l3, l2, m.xy_heatmaps[-1], m.zy_heatmaps[0], m.xz_heatmaps[-1]))
    for a, b in zip(*res):
        torch.testing.assert_close(b, a, rtol=1e-5, atol=1e-6)
    rg = dict(rm.named_parameters())
    for k, p in om.named_parameters():
        torch.testing.assert_close(p.grad, rg[k].grad, rtol=1e-4, atol=1e-6, msg=k)
    for k, v in rm.state_dict().items():   # running stats updated identically
        torch.testing.assert_close(om.state_dict()[k], v, rtol=1e-6, atol=1e-7, msg=k)
    # eval mode too
    rm.eval(); om.eval()
    with torch.no_grad():
        torch.testing.assert_close(om(x), rm(x), rtol=1e-5, atol=1e-6)


def test_unknown_model_raises(ref):
    with pytest.raises(Exception):
        M.create_oracle({'type': 'nope', 'version': '6.0.1', 'settings': {}})
    with pytest.raises(Exception):
        ref.models.create_model({'type': 'nope', 'version': '6.0.1', 'settings': {}})
