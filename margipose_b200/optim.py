"""Flat SGD for the B200 engine: torch.optim.SGD semantics (the optimiser the reference trains
with, /root/reference/src/margipose/bin/train_3d.py:338-340 + train_helpers.py:70-75) as ONE
kernel launch over the flat parameter / gradient buffers instead of ~1100 per-tensor updates.

It is a torch.optim.Optimizer, so LR / momentum schedulers that write `param_groups[0][name]`
(the reference's 1-cycle HyperparameterScheduler, hyperparam_scheduler.py:24-42) keep working,
and `state_dict()` / `load_state_dict()` carry the momentum buffer and the step count (the
reference saves `optimizer.state_dict()` in its checkpoints, train_3d.py:374-382).
"""
import torch

from . import ops


class FlatSGD(torch.optim.Optimizer):
    def __init__(self, model, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError('Nesterov momentum requires a momentum and zero dampening')
        device = next(model.parameters()).device
        if device.type != 'cuda':
            raise ValueError('FlatSGD needs the model on a CUDA device (call model.cuda() first)')
        model._ensure(device)
        self.model = model
        self.bank = model._bank
        frozen = [k for k, p in model.named_parameters() if not p.requires_grad]
        if frozen:
            raise ValueError('FlatSGD updates the whole flat parameter buffer; frozen parameters are not '
                             'supported (%s, ...)' % frozen[0])
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                        nesterov=nesterov)
        super().__init__(list(model.parameters()), defaults)
        self.momentum_buf = torch.zeros_like(self.bank.flat)
        self._steps = 0
        self.grad_scale = 1.0     # e.g. 1 / world_size after a summing all-reduce
        # (lr, momentum, dampening, weight_decay, grad_scale, first_step) as the kernel reads them: a pinned host
        # mirror that refresh_hyper() fills from param_groups and an async copy into device memory in front of
        # every step.  The copy is part of a captured step, so a replayed CUDA graph picks up whatever a
        # scheduler wrote -- including the "first step initialises the momentum buffer" flag.
        self._hyper_host = torch.zeros(8).pin_memory()
        self._hyper = torch.zeros(8, device=device)
        # The mirror is read by the GPU when the queued copy executes, not when it is queued: with steps in flight
        # (TrainStep.submit) a scheduler's next value must not reach the mirror before the previous step's copy has
        # run.  _hyper_copied marks the end of the last queued step; refresh_hyper() waits for it before it writes
        # CHANGED values (unchanged hyperparameters never wait).
        self._hyper_last = None
        self._hyper_copied = torch.cuda.Event()

    def add_param_group(self, param_group):
        if getattr(self, 'param_groups', None):
            raise ValueError('FlatSGD supports exactly one parameter group (one flat buffer, one set of '
                             'hyperparameters)')
        super().add_param_group(param_group)

    def refresh_hyper(self):
        """Host side only: publish the current param_groups hyperparameters to the pinned mirror (call before
        replaying a CUDA graph that contains step(); step() itself does it when run eagerly)."""
        g = self.param_groups[0]
        new = (float(g['lr']), float(g['momentum']), float(g['dampening']), float(g['weight_decay']),
               float(self.grad_scale), 1.0 if self._steps == 0 else 0.0)
        if new == self._hyper_last:
            return
        if not torch.cuda.is_current_stream_capturing():
            # no-op unless a step that reads the old values is still queued.  (While a step is being captured nothing
            # may synchronise -- and nothing is queued: TrainStep._capture synchronises the device first.)
            self._hyper_copied.synchronize()
        h = self._hyper_host
        for i, v in enumerate(new):
            h[i] = v
        self._hyper_last = new

    def _mark_hyper_copied(self):
        """Behind a queued step (eager or replayed): everything that reads the pinned mirror is in front of this."""
        if not torch.cuda.is_current_stream_capturing():
            self._hyper_copied.record(torch.cuda.current_stream(self._hyper.device))

    def zero_grad(self, set_to_none=False):
        self.bank.flat_grad.zero_()
        if set_to_none:
            for s in self.bank.params:
                getattr(s.mod, s.name).grad = None

    def state_dict(self):
        sd = super().state_dict()
        sd['flat'] = {'momentum_buf': self.momentum_buf.detach().clone(), 'steps': self._steps}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop('flat', None)
        super().load_state_dict(state_dict)
        if flat is not None:
            self.momentum_buf.copy_(flat['momentum_buf'])
            self._steps = int(flat['steps'])

    def count_replayed_step(self):
        """Host bookkeeping for a step replayed from a CUDA graph (step() itself did not run)."""
        self._steps += 1
        self.model.mark_params_dirty()
        self._mark_hyper_copied()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.bank is not self.model._bank:
            raise RuntimeError('the model was re-materialised (moved / re-created) after this optimiser '
                               'was built; create the optimiser after model.cuda()')
        g = self.param_groups[0]
        self.refresh_hyper()
        self._hyper.copy_(self._hyper_host, non_blocking=True)
        ops.sgd_step(self.bank.flat, self.bank.flat_grad, self.momentum_buf, g['lr'], g['momentum'],
                     g['dampening'], g['weight_decay'], g['nesterov'], first_step=(self._steps == 0),
                     grad_scale=self.grad_scale, hyper=self._hyper)
        self._steps += 1
        self.model.mark_params_dirty()
        self._mark_hyper_copied()
        return loss
