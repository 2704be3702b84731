"""oracle/ against the committed golden vectors the reference produced (CPU; runs anywhere)."""
import os

import pytest
import torch

from oracle import dsnt_oracle as D
from oracle import model_oracle as M
from tests.golden.make_golden import tail_inputs, model_inputs

GOLD = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'margipose_golden.pt'),
                  weights_only=False)
TOL = dict(rtol=1e-5, atol=1e-6)


def test_known_answer():
    k = GOLD['known_answer']
    for name, mu in (('xy', [-0.5, 0.5]), ('zy', [0.1, 0]), ('xz', [0, 0.2])):
        torch.testing.assert_close(D.make_gauss(torch.Tensor([[mu]]), (32, 32), 1), k[name], **TOL)
    c = D.heatmaps_to_coords(k['xy'], k['zy'], k['xz'])
    torch.testing.assert_close(c, k['coords'], **TOL)
    torch.testing.assert_close(c, k['expected'])     # reference tests/test_models.py:46


def test_joint_bookkeeping():
    assert M.JOINT_NAMES == GOLD['joint_names']
    assert M.JOINT_TREE == GOLD['joint_tree']
    assert M.HFLIP_INDICES == GOLD['hflip_indices']


@pytest.mark.parametrize('case', GOLD['tail'], ids=lambda c: 'x'.join(map(str, c['shape'])))
def test_tail(case):
    z, target, mask = tail_inputs(case['seed'], case['shape'], case['scale'])
    z = [t.requires_grad_() for t in z]
    p = [D.flat_softmax(t) for t in z]
    torch.testing.assert_close(D.heatmaps_to_coords(*p), case['coords'], **TOL)
    for k in range(3):
        torch.testing.assert_close(D.dsnt(p[k]), case['dsnt'][k], **TOL)
        torch.testing.assert_close(p[k].sum(-1), case['prob_rowsum'][k], **TOL)
        torch.testing.assert_close(p[k].sum(-2), case['prob_colsum'][k], **TOL)
    l3 = D.losses_3d([p[0]], [p[1]], [p[2]], target)
    l2 = D.losses_2d([p[0]], [p[1]], [p[2]], target)
    torch.testing.assert_close(l3, case['js'][0] + case['js'][1] + case['js'][2] + case['eu3'], **TOL)
    torch.testing.assert_close(l2, case['js'][0] + case['eu2'], **TOL)
    loss3 = D.average_loss(l3, mask)
    torch.testing.assert_close(loss3, case['loss3'], **TOL)
    torch.testing.assert_close(D.average_loss(l2, mask), case['loss2'], **TOL)
    g3 = torch.autograd.grad(loss3, z)
    for k in range(3):
        torch.testing.assert_close(g3[k].sum(-1), case['grad3_rowsum'][k], rtol=1e-4, atol=1e-7)
        if case['grad3'] is not None:
            torch.testing.assert_close(g3[k], case['grad3'][k], rtol=1e-4, atol=1e-8)
            torch.testing.assert_close(p[k], case['probs'][k], **TOL)


@pytest.mark.parametrize('case', GOLD['model'], ids=lambda c: c['name'])
def test_model(case):
    torch.manual_seed(case['weight_seed'])
    om = M.create_oracle(case['desc'])
    assert list(om.state_dict().keys()) == case['state_keys']
    assert sum(p.numel() for p in om.parameters()) == case['n_params']
    x, target, mask = model_inputs(case['input_seed'], case['batch'])
    om.train()
    out = om(x)
    torch.testing.assert_close(out, case['train_coords'], **TOL)
    l3 = D.average_loss(om.forward_3d_losses(out, target), mask)
    l2 = D.average_loss(om.forward_2d_losses(out, target), mask)
    torch.testing.assert_close(l3, case['loss3'], **TOL)
    torch.testing.assert_close(l2, case['loss2'], **TOL)
    for t in range(len(om.xy_heatmaps)):
        torch.testing.assert_close(om.xy_heatmaps[t].sum(-1), case['xy_rowsum'][t], **TOL)
        torch.testing.assert_close(om.zy_heatmaps[t].sum(-1), case['zy_rowsum'][t], **TOL)
        torch.testing.assert_close(om.xz_heatmaps[t].sum(-2), case['xz_colsum'][t], **TOL)
    l3.backward()
    for k, p in om.named_parameters():
        torch.testing.assert_close(p.grad.norm(), case['grad_norms'][k], rtol=1e-3, atol=1e-6, msg=k)
    grads = dict(om.named_parameters())
    for k, v in case['grad_probe'].items():
        torch.testing.assert_close(grads[k].grad.flatten()[:64], v, rtol=1e-3, atol=1e-6, msg=k)
    sd = om.state_dict()
    torch.testing.assert_close(sd['inner.in_cnn.1.running_mean'], case['running_mean_bn1'], **TOL)
    torch.testing.assert_close(sd['inner.in_cnn.1.running_var'], case['running_var_bn1'], **TOL)
    om.eval()
    with torch.no_grad():
        torch.testing.assert_close(om(x), case['eval_coords'], **TOL)
