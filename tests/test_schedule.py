"""Per-batch hyperparameter schedules: margipose_b200.hyperparam_scheduler against
numpy.interp (the arithmetic the reference's scheduler uses), against golden values the UNMODIFIED reference produced,
and -- where /root/reference exists -- against the reference's own make_1cycle, bit for bit; plus the optimiser /
schedule selection of bin/train_3d.py:338-347.  Host logic only (no GPU)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from margipose_b200.hyperparam_scheduler import HyperparameterScheduler, PiecewiseLinear, make_1cycle
from margipose_b200 import train

REF_FILE = '/root/reference/src/margipose/hyperparam_scheduler.py'

# make_1cycle(SGD(lr=0), max_iters=200, lr_max=0.05, momentum=0.9) of the reference (hyperparam_scheduler.py:6-42),
# (lr, momentum) after batch_step() number 1, 2, 45, 90, 91, 135, 180, 181, 199, 200, 201 -- printed with repr()
GOLDEN_1CYCLE = {
    1: (0.005000000000000001, 0.9),
    2: (0.005505617977528091, 0.8994382022471911),
    45: (0.027247191011235957, 0.8752808988764045),
    90: (0.05, 0.85),
    91: (0.0495, 0.8505555555555555),
    135: (0.027500000000000004, 0.875),
    180: (0.005000000000000001, 0.9),
    181: (0.004750250000000001, 0.9),
    199: (0.00025474999999999977, 0.9),
    200: (5.000000000000001e-06, 0.9),
    201: (5.000000000000001e-06, 0.9),
}


def _sgd(lr=0.0):
    return torch.optim.SGD([torch.nn.Parameter(torch.zeros(3))], lr=lr)


def _run(scheduler, n):
    out = {}
    for i in range(1, n + 1):
        scheduler.batch_step()
        g = scheduler.optimizer.param_groups[0]
        out[i] = (g['lr'], g['momentum'])
    return out


def test_piecewise_linear_is_numpy_interp_bit_for_bit():
    rng = np.random.default_rng(0)
    for _ in range(20):
        n = int(rng.integers(2, 7))
        ts = np.sort(rng.uniform(0, 1000, n))
        ys = rng.uniform(-3, 3, n)
        curve = PiecewiseLinear(ts, ys)
        xs = np.concatenate([ts, rng.uniform(-50, 1050, 200), np.arange(0, 1000, 37.0)])
        for x in xs:
            assert curve(x) == float(np.interp(x, ts, ys)), (x, ts, ys)
    with pytest.raises(ValueError):
        PiecewiseLinear([0, 1], [1.0])
    with pytest.raises(ValueError):
        PiecewiseLinear([1, 0], [1.0, 2.0])


def test_one_cycle_matches_the_reference_golden_values():
    got = _run(make_1cycle(_sgd(), 200, lr_max=0.05, momentum=0.9), 201)
    for i, want in GOLDEN_1CYCLE.items():
        assert got[i] == want, 'batch %d: %r != %r' % (i, got[i], want)
    lrs = [got[i][0] for i in range(1, 202)]
    assert max(lrs) == 0.05 and lrs.index(max(lrs)) + 1 == 90          # peak at 45 % of the run
    assert all(0.85 <= got[i][1] <= 0.9 for i in got)


@pytest.mark.skipif(not os.path.exists(REF_FILE), reason='reference tree not present')
@pytest.mark.parametrize('max_iters,lr_max,momentum', [(200, 0.05, 0.9), (1000, 1.0, 0.8), (37, 3e-3, 0), (10, 0.1, 0.95)])
def test_one_cycle_is_the_reference_schedule_bit_for_bit(max_iters, lr_max, momentum):
    spec = importlib.util.spec_from_file_location('_ref_hyperparam_scheduler', REF_FILE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    want = _run(ref.make_1cycle(_sgd(), max_iters, lr_max, momentum), max_iters + 3)
    got = _run(make_1cycle(_sgd(), max_iters, lr_max, momentum), max_iters + 3)
    assert got == want


def test_scheduler_rejects_what_the_reference_rejects():
    with pytest.raises(AssertionError, match='expected 2 milestones'):
        HyperparameterScheduler(_sgd(), ts=[1, 2], hyperparam_milestones={'lr': [1.0]})
    with pytest.raises(AssertionError, match='not an optimizer hyperparameter'):
        HyperparameterScheduler(_sgd(), ts=[1, 2], hyperparam_milestones={'beta': [1.0, 2.0]})


def test_scheduler_resumes_from_its_batch_count():
    a = make_1cycle(_sgd(), 100, lr_max=0.1, momentum=0.9)
    full = _run(a, 60)
    b = make_1cycle(_sgd(), 100, lr_max=0.1, momentum=0.9)
    _run(b, 25)
    c = make_1cycle(_sgd(), 100, lr_max=0.1, momentum=0.9)
    c.load_state_dict(b.state_dict())
    c.batch_step()
    g = c.optimizer.param_groups[0]
    assert (g['lr'], g['momentum']) == full[26]


def test_learning_schedule_selects_like_train_3d():
    """bin/train_3d.py:338-347 / train_helpers.py:57-77 with a CPU stand-in for the flat optimiser."""
    made = []

    def sgd(model, lr, momentum=0.0, nesterov=False):
        opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, nesterov=nesterov)
        made.append(opt)
        return opt

    model = torch.nn.Linear(2, 2)
    s = train.learning_schedule(model, '1cycle', 0.05, max_iters=200, sgd=sgd)
    assert s.optimizer is made[-1] and s.optimizer.param_groups[0]['lr'] == 0
    assert _run(s, 90)[90] == GOLDEN_1CYCLE[90]
    with pytest.raises(ValueError):
        train.learning_schedule(model, '1cycle', 0.05, sgd=sgd)

    s = train.learning_schedule(model, 'sgd_simple', 0.01, sgd=sgd)
    assert s.optimizer.param_groups[0]['lr'] == 0.01 and not hasattr(s, 'batch_step') and not hasattr(s, 'step')

    s = train.learning_schedule(model, 'nesterov', 0.1, lr_milestones=[2, 4], lr_gamma=0.5, sgd=sgd)
    g = s.optimizer.param_groups[0]
    assert g['nesterov'] and g['momentum'] == 0.8
    lrs = []
    for _epoch in range(6):
        lrs.append(g['lr'])
        s.optimizer.step()
        s.step()
    assert lrs == [0.1, 0.1, 0.05, 0.05, 0.025, 0.025]

    s = train.learning_schedule(model, 'sgd', 0.1, lr_milestones=[1], sgd=sgd)
    assert s.optimizer.param_groups[0]['momentum'] == 0
    with pytest.raises(Exception, match='unrecognised optimisation algorithm'):
        train.learning_schedule(model, 'rmsprop', 0.1, sgd=sgd)
