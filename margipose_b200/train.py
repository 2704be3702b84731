"""One MargiPose training step as a user calls it: the fwd / loss / bwd / step sequence of the
reference's `do_training_pass` (/root/reference/src/margipose/bin/train_3d.py:159-186) with
`forward_loss` (:126-142, 3D / 2D / mixed batches) as the loss, on the B200 engine.

    step = TrainStep(model, optimizer, batch=32)
    loss = step(images, targets, joint_mask, valid_depth)      # host or device tensors

All device buffers are static, so after a few eager iterations the whole step (≈700 kernel
launches for the 4-stage ResNet-34 model) is captured once into CUDA graphs and replayed; the
per-step host work is then a few async H2D copies, the graph launches and one 4-byte loss
read-back.  The loss is computed by the fused tail kernels inside the engine's programs: per stage
ONE forward launch (softmax + soft-argmax + xyz + Gaussian + JS x3 + Euclid, accumulating the
per-joint loss) and ONE backward launch (loss gradient + combiner gradient + softmax backward).

With torch.distributed initialised, gradients are summed across ranks bucket by bucket: the flat
gradient buffer is laid out stage by stage, the backward program is cut where a stage's parameter
gradients are complete, and each bucket's all-reduce (NCCL over NVLink) is issued from a side stream
while the next stage's backward runs; only the small stem bucket is exposed.  The 1 / world_size
factor is folded into the SGD kernel (no extra pass over the gradients).
"""
import torch

from . import dsntnn as K
from . import parallel


def forward_loss(model, out_var, target_var, mask_var, valid_depth):
    """Mirror of bin/train_3d.py:126-142: the 3D loss for samples with valid depth, the 2D loss for the
    others, masked-averaged over joints.  `valid_depth`: per-sample flags (list / tensor).  The reference
    stacks per-sample rows in a Python loop; here the flag goes into the fused tail kernels."""
    target_var = target_var.narrow(-1, 0, 3)
    if not torch.is_tensor(valid_depth):
        valid_depth = torch.as_tensor(list(valid_depth))
    vd = valid_depth.to(device=target_var.device, dtype=torch.int32)
    if not 0 in vd.tolist():
        losses = model.forward_3d_losses(out_var, target_var)
    elif not 1 in vd.tolist():
        losses = model.forward_2d_losses(out_var, target_var)
    else:
        losses = model.forward_mixed_losses(out_var, target_var, vd)
    return K.average_loss(losses, mask_var)


def learning_schedule(model, optim_algorithm, lr, max_iters=None, lr_milestones=(), lr_gamma=0.1, sgd=None):
    """The optimiser / schedule selection of bin/train_3d.py:338-347 (+ train_helpers.py:57-77) on the flat optimiser:
    returns an object with `.optimizer` and, depending on the algorithm, `.batch_step()` (call per batch) or
    `.step()` (call per epoch), which is how `do_training_pass` (:145-156) drives it.

      '1cycle'     SGD(lr=0) under the 1-cycle learning-rate / momentum schedule over `max_iters` batches (momentum 0.9)
      'sgd_simple' plain SGD, no schedule
      'sgd'        SGD, learning rate multiplied by `lr_gamma` at the epochs in `lr_milestones`
      'nesterov'   the same with Nesterov momentum 0.8
    RMSprop (the reference's third milestone algorithm) has no flat kernel here and is rejected.
    `sgd`: optimiser factory `sgd(model, lr=..., momentum=..., nesterov=...)`, default `FlatSGD`."""
    from collections import namedtuple
    from .hyperparam_scheduler import make_1cycle
    if sgd is None:
        from .optim import FlatSGD as sgd
    if optim_algorithm == '1cycle':
        if not max_iters:
            raise ValueError("'1cycle' needs max_iters = epochs * batches per epoch")
        return make_1cycle(sgd(model, lr=0), max_iters, lr_max=lr, momentum=0.9)
    if optim_algorithm == 'sgd_simple':
        return namedtuple('DummyScheduler', 'optimizer')(optimizer=sgd(model, lr=lr, momentum=0))
    if optim_algorithm == 'sgd':
        optimiser = sgd(model, lr=lr)
    elif optim_algorithm == 'nesterov':
        optimiser = sgd(model, lr=lr, momentum=0.8, nesterov=True)
    else:
        raise Exception('unrecognised optimisation algorithm: ' + optim_algorithm)
    return torch.optim.lr_scheduler.MultiStepLR(optimiser, milestones=list(lr_milestones), gamma=lr_gamma)


class PendingLoss:
    """Loss of a step queued with `TrainStep.submit`: a pinned host scalar the GPU writes when the step is done."""

    def __init__(self, host, event):
        self.host, self.event = host, event

    def done(self):
        return self.event.query()

    def item(self):
        self.event.synchronize()
        return self.host.item()


class TrainStep:
    def __init__(self, model, optimizer, batch, height=256, width=256, use_graph=True, warmup=3,
                 fused_loss=True, overlap_allreduce=True):
        self.model, self.opt = model, optimizer
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise ValueError('TrainStep needs the model on a CUDA device')
        self.device = dev
        J = model.n_joints
        model.train()
        model._ensure(dev)
        self.eng = model.engine_for(batch, height, width, True)
        self.fused = bool(fused_loss) and self.eng.group
        self.x = self.eng.x_in if self.fused else torch.zeros(batch, 3, height, width, device=dev)
        if self.fused:
            L = self.eng.enable_fused_loss(pixelwise=model._pixelwise_flag())
            self.target, self.mask, self.valid_depth = L.target, L.mask, L.valid_depth
            self.coords = L.coords[-1]
            self.loss = L.out2[0]
        else:
            self.target = torch.zeros(batch, J, 3, device=dev)
            self.mask = torch.ones(batch, J, device=dev)
            self.valid_depth = None
            self.coords = torch.zeros(batch, J, 3, device=dev)
            self.loss = torch.zeros((), device=dev)
        self.use_graph = use_graph
        self.warmup = warmup
        self._graphs = None
        self._eager_runs = 0
        self.world = parallel.world()[1]
        self.overlap = bool(overlap_allreduce) and self.world > 1 and self.fused
        self.comm = torch.cuda.Stream(device=dev) if self.overlap else None
        if self.world > 1 and hasattr(optimizer, 'grad_scale'):
            optimizer.grad_scale = 1.0 / self.world      # the all-reduce sums; the SGD kernel scales
        self._scale_in_opt = self.world > 1 and hasattr(optimizer, 'grad_scale')
        # host -> device double buffering (prefetch()): two staging slots filled on a copy stream
        self._copy_stream = None
        self._slots = []
        self._staged = {}          # id(images tensor) -> slot index
        self._next_slot = 0
        self._loss_ring = []       # pinned host scalars + events of the steps whose loss has not been read yet
        self._n_submitted = 0
        # backward program slices [lo, hi) and the flat-gradient ranges that are final after each
        self._pieces = parallel.bucket_plan(self.eng.bwd_marks, self.eng.L.stage_ranges,
                                            model._bank.flat_grad.numel())

    # ---- pieces of the step
    def _forward_and_loss(self):
        model, eng = self.model, self.eng
        self.opt.zero_grad()
        if self.fused:
            # bf16 operand refresh: the feature extractor's packs first, the columns' (95 % of the weights) on the
            # auxiliary stream beside the feature extractor's forward pass
            bank, cur, aux = model._bank, torch.cuda.current_stream(self.device), eng.aux[0]
            bank.pack(part=0)
            aux.wait_stream(cur)
            with torch.cuda.stream(aux):
                bank.pack(part=1)
            model._packed_version = None
            probs = eng.forward(self.x, fused=True, join_after_stem=aux)
            model.xy_heatmaps = [row[0] for row in probs]
            model.zy_heatmaps = [row[1] for row in probs]
            model.xz_heatmaps = [row[2] for row in probs]
            eng.loss_reduce()
            model._bank.attach_grads()
            return None
        out = model(self.x)
        loss = K.average_loss(model.forward_3d_losses(out, self.target), self.mask)
        self.loss.copy_(loss.detach())
        self.coords.copy_(out.detach())
        return loss

    def _fwd_bwd(self):
        loss = self._forward_and_loss()
        if self.fused:
            self.eng.backward()
        else:
            loss.backward()

    def _update(self):
        self.opt.step()

    def _allreduce_all(self):
        if self.world > 1:
            g = self.model.flat_grads
            if self._scale_in_opt:
                parallel.allreduce_sum_(g)
            else:
                parallel.allreduce_mean_(g)

    def _reduce_ranges(self, ranges, works):
        """All-reduce finished gradient ranges on the side stream, behind everything issued so far."""
        g = self.model.flat_grads
        self.comm.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.comm):
            for lo, hi in ranges:
                works.append(parallel.allreduce_sum_(g[lo:hi], async_op=True))

    def run(self):
        """One step on whatever is in the static buffers (self.x / self.target / self.mask)."""
        if self.use_graph and self._graphs is None and self._eager_runs >= self.warmup:
            self._capture()
        if self._graphs is not None:
            if self.overlap:
                works = []
                for g, (_lo, _hi, ranges) in zip(self._graphs[:-1], self._pieces):
                    g.replay()
                    self._reduce_ranges(ranges, works)
                for w in works:
                    if w is not None:
                        w.wait()
            else:
                self._graphs[0].replay()
                self._allreduce_all()
            if hasattr(self.opt, 'refresh_hyper'):
                self.opt.refresh_hyper()   # LR / momentum schedules reach the captured optimiser step
            self._graphs[-1].replay()
            if hasattr(self.opt, 'count_replayed_step'):
                self.opt.count_replayed_step()
        else:
            if self.overlap:
                works = []
                self._forward_and_loss()
                for lo, hi, ranges in self._pieces:
                    self.eng.backward(lo=lo, hi=hi)
                    self._reduce_ranges(ranges, works)
                for w in works:
                    if w is not None:
                        w.wait()
            else:
                self._fwd_bwd()
                self._allreduce_all()
            self._update()
            self._eager_runs += 1

    def _capture(self):
        torch.cuda.synchronize(self.device)
        graphs = []
        if self.overlap:
            pool = None
            for i, (lo, hi, _ranges) in enumerate(self._pieces):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    if i == 0:
                        self._forward_and_loss()
                    self.eng.backward(lo=lo, hi=hi)
                pool = pool or g.pool()
                graphs.append(g)
        else:
            g0 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g0):
                self._fwd_bwd()
            graphs.append(g0)
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, pool=graphs[0].pool()):
            self._update()
        # the capture ran step() once on the host without executing it: undo its bookkeeping
        if hasattr(self.opt, '_steps'):
            self.opt._steps -= 1
        graphs.append(g1)
        self._graphs = tuple(graphs)

    def load(self, images, targets, mask=None, valid_depth=None):
        self.x.copy_(images, non_blocking=True)
        self.target.copy_(targets[..., :3], non_blocking=True)
        if mask is not None:
            self.mask.copy_(mask, non_blocking=True)
        if valid_depth is not None:
            if not self.fused:
                raise ValueError('per-sample valid_depth flags need the fused loss path')
            self.valid_depth.copy_(torch.as_tensor(valid_depth).to(torch.int32), non_blocking=True)

    def prefetch(self, images, targets, mask=None, valid_depth=None):
        """Starts the host -> device copy of a LATER batch (pinned host tensors) on a copy stream, so it overlaps
        the step that is running; the `__call__` that is given the same `images` tensor picks the staged copy up
        with a device-to-device copy instead of waiting for PCIe.  What a prefetching data loader does; at most
        two batches can be staged at a time."""
        dev = self.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            for _ in range(2):
                slot = dict(x=torch.empty_like(self.x), target=torch.empty_like(self.target),
                            mask=torch.empty_like(self.mask),
                            vd=torch.empty_like(self.valid_depth) if self.valid_depth is not None else None,
                            ready=torch.cuda.Event(), free=torch.cuda.Event(), has_mask=False, has_vd=False)
                slot['free'].record(torch.cuda.current_stream(dev))
                self._slots.append(slot)
        k = self._next_slot
        self._next_slot ^= 1
        self._staged = {key: v for key, v in self._staged.items() if v != k}
        slot = self._slots[k]
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(slot['free'])       # the step that consumed this slot has copied it out
            slot['x'].copy_(images, non_blocking=True)
            slot['target'].copy_(targets[..., :3], non_blocking=True)
            slot['has_mask'] = mask is not None
            if mask is not None:
                slot['mask'].copy_(mask, non_blocking=True)
            slot['has_vd'] = valid_depth is not None
            if valid_depth is not None:
                if not self.fused:
                    raise ValueError('per-sample valid_depth flags need the fused loss path')
                slot['vd'].copy_(torch.as_tensor(valid_depth).to(torch.int32), non_blocking=True)
            slot['ready'].record(self._copy_stream)
        self._staged[id(images)] = k

    def _take_staged(self, images):
        k = self._staged.pop(id(images), None)
        if k is None:
            return False
        slot = self._slots[k]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(slot['ready'])
        self.x.copy_(slot['x'])
        self.target.copy_(slot['target'])
        if slot['has_mask']:
            self.mask.copy_(slot['mask'])
        if slot['has_vd']:
            self.valid_depth.copy_(slot['vd'])
        slot['free'].record(cur)
        return True

    def submit(self, images, targets, mask=None, valid_depth=None, prefetch=None):
        """Queues one step and returns a `PendingLoss` without waiting for the GPU: the batch is copied in (pinned host
        tensors copy asynchronously; a batch announced with `prefetch` is already on the device), the step's launches
        and a 4-byte device-to-host copy of its loss are queued, then the NEXT batch's host -> device copy is started
        (`prefetch`: a tuple (images, targets[, mask[, valid_depth]])).  `PendingLoss.item()` waits for that step only,
        so a loop that reads step i's loss after submitting step i + 1 never leaves the GPU idle while the host turns
        around.  Up to four steps may be pending."""
        if not self._take_staged(images):
            self.load(images, targets, mask, valid_depth)
        self.run()
        if not self._loss_ring:
            self._loss_ring = [PendingLoss(torch.zeros(1, dtype=torch.float32).pin_memory(), torch.cuda.Event())
                               for _ in range(4)]
        pending = self._loss_ring[self._n_submitted % len(self._loss_ring)]
        self._n_submitted += 1
        pending.host.copy_(self.loss.reshape(1), non_blocking=True)
        pending.event.record(torch.cuda.current_stream(self.device))
        if prefetch is not None:
            self.prefetch(*prefetch)
        return pending

    def __call__(self, images, targets, mask=None, valid_depth=None, prefetch=None):
        """One step, returning its loss as a Python float (a 4-byte device-to-host read, like train_3d.py:167):
        `submit(...).item()`.
        valid_depth: optional per-sample flags, 1 = 3D loss, 0 = 2D loss (bin/train_3d.py:126-142).
        prefetch: the NEXT batch as a tuple (images, targets[, mask[, valid_depth]]): its host -> device copy is
        started after this step's launches have been queued and before its loss is waited for."""
        return self.submit(images, targets, mask, valid_depth, prefetch).item()

    def launches_per_step(self):
        """Kernel launches of OUR library in one step (for bench.py's gpu_launches)."""
        eng = self.eng
        n_stages = len(eng.probs)
        if self.fused:
            extra = 1 + 2 + 1               # pack, masked mean fwd + bwd, sgd (stage tails are in the programs)
        else:
            extra = 1 + 2 * n_stages + 1 + 2 + 1   # pack, stage losses fwd+bwd, coords, masked mean fwd+bwd, sgd
        return eng.launches() + extra
