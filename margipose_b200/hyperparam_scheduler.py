"""Per-batch hyperparameter schedules for the training step.

Host-side mirror of the reference's scheduler interface (/root/reference/src/margipose/hyperparam_scheduler.py:6-42,
driven from bin/train_3d.py:155-156 once per batch): `make_1cycle(optimizer, max_iters, lr_max, momentum)` returns an
object with `.optimizer` and `.batch_step()`, and `batch_step()` writes the interpolated values into every
`param_groups[i][name]`.  With `FlatSGD` those values reach the fused SGD kernel through its device-side hyperparameter
block, so the schedule also drives a step replayed from CUDA graphs (optim.py).

The schedule itself is a piecewise-linear curve through (t_k, v_k) knots, clamped outside [t_0, t_n] -- evaluated with
the same arithmetic as numpy.interp (slope * (t - t_k) + v_k in double precision), which the reference calls, so the
values are bit-identical (tests/test_host_logic.py pins them against numpy and, where present, the reference itself).
"""
import bisect


class PiecewiseLinear:
    """y(t) through the knots (ts[k], ys[k]); constant before the first and after the last knot."""

    def __init__(self, ts, ys):
        ts, ys = [float(t) for t in ts], [float(y) for y in ys]
        if len(ts) != len(ys) or not ts:
            raise ValueError('expected as many values as knots (%d knots, %d values)' % (len(ts), len(ys)))
        if any(b < a for a, b in zip(ts, ts[1:])):
            raise ValueError('knots must be non-decreasing')
        self.ts, self.ys = ts, ys

    def __call__(self, t):
        ts, ys = self.ts, self.ys
        t = float(t)
        if t <= ts[0]:
            return ys[0]
        if t >= ts[-1]:
            return ys[-1]
        k = bisect.bisect_right(ts, t) - 1          # ts[k] <= t < ts[k + 1]
        slope = (ys[k + 1] - ys[k]) / (ts[k + 1] - ts[k])
        return slope * (t - ts[k]) + ys[k]


class HyperparameterScheduler:
    """Sets optimiser hyperparameters batch by batch.  `hyperparam_milestones[name][k]` is the value of
    `param_groups[*][name]` at batch count `ts[k]`; in between the value is interpolated linearly."""

    def __init__(self, optimizer, ts, hyperparam_milestones):
        for name, values in hyperparam_milestones.items():
            assert len(values) == len(ts), \
                'expected {} milestones for hyperparameter "{}"'.format(len(ts), name)
            for group in optimizer.param_groups:
                assert name in group, '"{}" is not an optimizer hyperparameter'.format(name)
        self.optimizer = optimizer
        self.ts = list(ts)
        self.hyperparam_milestones = {k: list(v) for k, v in hyperparam_milestones.items()}
        self._curves = {k: PiecewiseLinear(self.ts, v) for k, v in self.hyperparam_milestones.items()}
        self.batch_count = 0

    def values_at(self, batch_count):
        return {name: curve(batch_count) for name, curve in self._curves.items()}

    def batch_step(self):
        """Call once per batch, before the optimiser step of that batch (bin/train_3d.py:155-156)."""
        self.batch_count += 1
        for name, value in self.values_at(self.batch_count).items():
            for group in self.optimizer.param_groups:
                group[name] = value

    # checkpoint / resume: the reference restarts its schedule from zero; carrying the count is an extension
    def state_dict(self):
        return {'batch_count': self.batch_count}

    def load_state_dict(self, state):
        self.batch_count = int(state['batch_count'])


def make_1cycle(optimizer, max_iters, lr_max, momentum=0):
    """The 1-cycle policy (arXiv:1803.09820) as the reference parameterises it: the learning rate climbs from
    lr_max / 10 to lr_max over the first 45 % of the run, returns to lr_max / 10 at 90 % and decays to 1e-3 of that in
    the last tenth; the momentum mirrors it between `momentum` and min(momentum, 0.85)."""
    lr_floor = lr_max * 1e-1
    lr_end = lr_floor * 1e-3
    t_end = max_iters
    t_down = 0.9 * t_end
    t_peak = t_down / 2
    m_hi = momentum
    m_lo = min(m_hi, 0.85)
    return HyperparameterScheduler(optimizer, ts=[1, t_peak, t_down, t_end], hyperparam_milestones={
        'lr': [lr_floor, lr_max, lr_floor, lr_end],
        'momentum': [m_hi, m_lo, m_hi, m_hi],
    })
