"""Data parallelism for the MargiPose hot path: one process per GPU, batch sharded across ranks,
ONE gradient all-reduce per step over the flat gradient buffer.

The reference has no multi-GPU path (single `--device`, /root/reference/src/margipose/cli.py:11);
BatchNorm batch statistics stay per replica (torch DistributedDataParallel does not synchronise them
either; its default `broadcast_buffers=True` re-broadcasts rank 0's running buffers every forward, which
only affects eval -- here `sync_model` broadcasts them once at start).
Collectives go through torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_batch(global_batch, rank=None, world_size=None):
    """Contiguous, near-equal shard [lo, hi) of a global batch for this rank."""
    if rank is None:
        rank, world_size = world()
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_flat(tensors, src=0):
    """Rank `src`'s flat buffers (parameters, BatchNorm buffers) to everyone, at start-up."""
    _, ws = world()
    if ws > 1:
        for t in tensors:
            dist.broadcast(t, src)


def allreduce_sum_(t, async_op=False):
    """In-place sum across ranks (the 1 / world factor then goes into the SGD kernel's grad_scale).
    async_op=True returns the collective's Work handle (None when there is nothing to do)."""
    _, ws = world()
    if ws == 1 or t.numel() == 0:
        return None
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)


def bucket_plan(bwd_marks, stage_ranges, n_param):
    """Cuts the backward program where parameter gradients become final.  bwd_marks[i]: program index after
    which stage (n_stages - 1 - i) is done (stages run backwards); stage_ranges[t] = [lo, hi) of stage t in the
    flat gradient buffer.  Returns [(program lo, program hi or None, [gradient ranges final after it])]; the
    last piece (ResNet stem backward) carries everything outside the stages (stem, adapter, combiners)."""
    n_stages = len(bwd_marks)
    assert len(stage_ranges) == n_stages
    pieces, lo = [], 0
    for i, hi in enumerate(bwd_marks):
        pieces.append((lo, hi, [tuple(stage_ranges[n_stages - 1 - i])]))
        lo = hi
    stage_lo = min(r[0] for r in stage_ranges)
    stage_hi = max(r[1] for r in stage_ranges)
    rest = [r for r in ((0, stage_lo), (stage_hi, n_param)) if r[1] > r[0]]
    pieces.append((lo, None, rest))
    return pieces


def allreduce_mean_(flat_grad, extra=None):
    """In-place mean of the flat gradient buffer across ranks: one collective for the step.
    `extra` (optional small tensor, e.g. a loss / mask-count pair) is summed alongside."""
    _, ws = world()
    if ws == 1:
        return flat_grad
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    flat_grad.mul_(1.0 / ws)
    if extra is not None:
        dist.all_reduce(extra, op=dist.ReduceOp.SUM)
    return flat_grad


def sync_model(model, src=0):
    """Makes every replica start from rank `src`'s weights and buffers."""
    bank = model._bank
    broadcast_flat([bank.flat, bank.flat_buf, bank.flat_cnt], src)
    model.mark_params_dirty()
