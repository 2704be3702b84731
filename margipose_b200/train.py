"""One MargiPose training step as a user calls it: the fwd / loss / bwd / step sequence of the
reference's `do_training_pass` (/root/reference/src/margipose/bin/train_3d.py:159-186) with
`forward_loss` (:126-142, all-3D branch) as the loss, on the B200 engine.

    step = TrainStep(model, optimizer, batch=32)
    loss = step(images, targets, joint_mask)        # host or device tensors

All device buffers are static, so after a few eager iterations the whole step (≈730 kernel
launches for the 4-stage ResNet-34 model) is captured once into a CUDA graph and replayed; the
per-step host work is then two async H2D copies, one graph launch and one 4-byte loss read-back.
With torch.distributed initialised, gradients are averaged across ranks with ONE all-reduce over
the flat gradient buffer between the backward graph and the optimiser graph.
"""
import torch

from . import dsntnn as K
from . import parallel


class TrainStep:
    def __init__(self, model, optimizer, batch, height=256, width=256, use_graph=True, warmup=3):
        self.model, self.opt = model, optimizer
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise ValueError('TrainStep needs the model on a CUDA device')
        self.device = dev
        J = model.n_joints
        self.x = torch.zeros(batch, 3, height, width, device=dev)
        self.target = torch.zeros(batch, J, 3, device=dev)
        self.mask = torch.ones(batch, J, device=dev)
        self.loss = torch.zeros((), device=dev)
        self.coords = torch.zeros(batch, J, 3, device=dev)
        self.use_graph = use_graph
        self.warmup = warmup
        self._graphs = None
        self._eager_runs = 0
        self.world = parallel.world()[1]
        model.train()

    # ---- the step, split where the gradient all-reduce goes
    def _fwd_bwd(self):
        self.opt.zero_grad()
        out = self.model(self.x)
        loss = K.average_loss(self.model.forward_3d_losses(out, self.target), self.mask)
        loss.backward()
        self.loss.copy_(loss.detach())
        self.coords.copy_(out.detach())

    def _update(self):
        self.opt.step()

    def _allreduce(self):
        if self.world > 1:
            parallel.allreduce_mean_(self.model.flat_grads)

    def run(self):
        """One step on whatever is in the static buffers (self.x / self.target / self.mask)."""
        if self.use_graph and self._graphs is None and self._eager_runs >= self.warmup:
            self._capture()
        if self._graphs is not None:
            self._graphs[0].replay()
            self._allreduce()
            if hasattr(self.opt, 'refresh_hyper'):
                self.opt.refresh_hyper()   # LR / momentum schedules reach the captured optimiser step
            self._graphs[1].replay()
        else:
            self._fwd_bwd()
            self._allreduce()
            self._update()
            self._eager_runs += 1

    def _capture(self):
        torch.cuda.synchronize(self.device)
        g0, g1 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g0):
            self._fwd_bwd()
        with torch.cuda.graph(g1, pool=g0.pool()):
            self._update()
        self._graphs = (g0, g1)

    def load(self, images, targets, mask=None):
        self.x.copy_(images, non_blocking=True)
        self.target.copy_(targets, non_blocking=True)
        if mask is not None:
            self.mask.copy_(mask, non_blocking=True)

    def __call__(self, images, targets, mask=None):
        """Copies one batch in (pinned host tensors copy asynchronously), runs the step and returns
        the loss as a Python float (a 4-byte device-to-host read, like train_3d.py:167)."""
        self.load(images, targets, mask)
        self.run()
        return self.loss.item()

    def launches_per_step(self):
        """Kernel launches of OUR library in one step (for bench.py's gpu_launches)."""
        eng = self.model.engine_for(self.x.size(0), self.x.size(2), self.x.size(3), True)
        n_stages = len(eng.probs)
        extra = 1 + 2 * n_stages + 1 + 2 + 1   # pack, stage losses fwd+bwd, coords, masked mean fwd+bwd, sgd
        return eng.launches() + extra
