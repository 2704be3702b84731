#!/usr/bin/env python
"""Benchmark of the MargiPose training hot path on B200 (BASELINE.json metric: images/sec fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|torch-gpu]
                    [--config r34x4_256|r50x5_384]

Workload (config.workload): BASELINE.json configs[1] -- 4-stage ResNet-34 MargiPose, 256x256,
17 joints, batch 32 per GPU, one training step = forward + forward_3d_losses/average_loss +
backward (+ one gradient all-reduce when N > 1) + SGD-momentum update, synthetic images/targets,
seeded random-init weights (no network for datasets / checkpoints).

  value : steps run from device-resident inputs (CUDA events, max over ranks)
  e2e   : the same step through the public API (margipose_b200.train.TrainStep.submit, the call behind
          TrainStep.__call__) fed from pinned HOST buffers each step, every step's loss read back to the host
          (step i's loss is waited for after step i + 1 has been queued)
  roofline : the tcgen05 implicit-GEMM conv kernel (mp_conv_igemm: the grouped fprop + dgrad launches
          of one step, each alone on the GPU, replayed from a CUDA graph), algorithmic FLOPs / CUDA-event
          time per launch vs the measured sustained bf16 peak
  cpu_baseline / --impl reference : the reference algorithm's CPU path (the fp32 oracle port
          of /root/reference/src/margipose, the unmodified reference cannot travel to the GPU box)
          on the host cores, on a bounded sample of the same workload
  gpu_library_baseline / --impl torch-gpu : the same-box GPU comparator of SURVEY.md section 8(d): the oracle
          module (stock nn.Conv2d / BatchNorm2d / ATen ops = cuDNN + cuBLAS, cudnn.benchmark as the reference
          sets it, utils.py:23-24) running the SAME step on the B200 in fp32, TF32 and bf16-autocast
  inference : eval-mode forward (bin/infer_single.py:58-66, bin/eval_3d.py:60-62) through InferStep (BatchNorm
          folded into the convs, one captured CUDA graph), batch 1 latency and batch 32 throughput
--config r50x5_384 measures BASELINE.json configs[4] (ResNet-50, 5 stages, 384x384, 16 images per GPU).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

JOINTS = 17
CONFIGS = {   # conv FLOPs per image fwd+bwd: SURVEY.md section 6 (FlopCounterMode on the reference)
    'r34x4_256': dict(fe='resnet34', n_stages=4, res=256, batch=32, flops=161.35e9, cpu_batch=8,
                      workload='configs[1]: 4-stage ResNet-34 MargiPose, 256x256, 17 joints',
                      metric='images/sec fwd+bwd (4-stage ResNet-34 MargiPose, 256x256, 17 joints)'),
    'r50x5_384': dict(fe='resnet50', n_stages=5, res=384, batch=16, flops=449.75e9, cpu_batch=2,
                      workload='configs[4]: ResNet-50 backbone, 5 stages, 384x384, 17 joints',
                      metric='images/sec fwd+bwd (5-stage ResNet-50 MargiPose, 384x384, 17 joints)'),
}
CFG = CONFIGS['r34x4_256']
DESC = RES = FLOPS_PER_IMAGE = METRIC = None


def select_config(name):
    global CFG, DESC, RES, FLOPS_PER_IMAGE, METRIC
    CFG = CONFIGS[name]
    DESC = {'type': 'margipose', 'version': '6.0.1',
            'settings': {'n_stages': CFG['n_stages'], 'axis_permutation': True,
                         'feature_extractor': CFG['fe'], 'pixelwise_loss': 'jsd'}}
    RES, FLOPS_PER_IMAGE, METRIC = CFG['res'], CFG['flops'], CFG['metric']


select_config('r34x4_256')


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get('bf16_tflops_sustained', 1415.1), p.get('hbm_gbs', 6449.1), 'measured'
    return 1400.0, 6650.0, 'fallback'


class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx = max(mx, float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def synthetic(batch, n_sets, seed, device=None, pinned=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    sets = []
    for _ in range(n_sets):
        x = torch.randn(batch, 3, RES, RES, generator=g)
        t = torch.rand(batch, JOINTS, 3, generator=g) * 1.6 - 0.8
        m = torch.ones(batch, JOINTS)
        if pinned:
            x, t, m = x.pin_memory(), t.pin_memory(), m.pin_memory()
        elif device is not None:
            x, t, m = x.to(device), t.to(device), m.to(device)
        sets.append((x, t, m))
    return sets


# ------------------------------------------------------------------------------ CPU reference
def cpu_reference(batch, steps, warmup, budget_s):
    """fwd + loss + bwd + SGD of the reference algorithm on the host cores (fp32 oracle port)."""
    import torch
    from oracle import model_oracle as M
    from oracle import dsnt_oracle as D
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    om = M.create_oracle(DESC).train()
    opt = torch.optim.SGD(om.parameters(), lr=1e-3, momentum=0.9)
    data = synthetic(batch, 2, seed=1)
    times = []
    t_begin = time.perf_counter()
    done = 0
    for i in range(warmup + steps):
        x, t, m = data[i % len(data)]
        t0 = time.perf_counter()
        opt.zero_grad()
        out = om(x)
        loss = D.average_loss(om.forward_3d_losses(out, t), m)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            done += 1
        if i >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    total = sum(times)
    return {'value': batch * done / total, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d step(s) of batch %d (fwd+loss+bwd+SGD) after %d warm-up, fp32 oracle port of the '
                      'reference on CPU' % (done, batch, warmup),
            'steps': done, 'ms_per_step': 1e3 * total / max(done, 1)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    warm = max(1, min(args.warmup, 10))      # the warm-up the driver asks for (a CPU step takes ~1 s)
    base = cpu_reference(batch=CFG['cpu_batch'], steps=max(1, args.steps), warmup=warm, budget_s=150.0)
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'images/s', 'n_gpus': args.gpus,
            'steps': base['steps'], 'warmup': warm, 'ms_per_step': base['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': '%s, training step; CPU sample of batch %d per step'
                                   % (CFG['workload'], CFG['cpu_batch'])},
            'cpu_baseline': {k: base[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': base['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------- stock PyTorch on the same GPU
def gpu_library_baseline(batch, steps=5, warmup=3):
    """SURVEY.md section 8(d) / BASELINE.md 4.6: the reference module's training step (bin/train_3d.py:159-186)
    executed by stock PyTorch on the B200 -- cuDNN convolutions / BatchNorm, ATen softmax and elementwise ops,
    autograd, torch.optim.SGD -- in the three precisions a user could pick.  channels_last + cudnn.benchmark
    (utils.py:23-24).  None of this repository's kernels are on this path; it is the bar to beat on the box."""
    import torch
    from oracle import model_oracle as M
    from oracle import dsnt_oracle as D
    dev = torch.device('cuda', torch.cuda.current_device())
    torch.backends.cudnn.benchmark = True
    x, t, m = synthetic(batch, 1, seed=7, device=dev)[0]
    x = x.contiguous(memory_format=torch.channels_last)
    out = {'module': 'oracle restatement of the reference nn.Module (same op sites), stock PyTorch %s / cuDNN %s'
                     % (torch.__version__, torch.backends.cudnn.version()),
           'batch': batch, 'steps': steps, 'warmup': warmup, 'unit': 'images/s'}
    for name in ('fp32', 'tf32', 'bf16_autocast'):
        torch.backends.cudnn.allow_tf32 = name != 'fp32'
        torch.backends.cuda.matmul.allow_tf32 = name != 'fp32'
        torch.manual_seed(0)
        om = M.create_oracle(DESC).to(dev).to(memory_format=torch.channels_last).train()
        opt = torch.optim.SGD(om.parameters(), lr=1e-3, momentum=0.9)

        def one():
            opt.zero_grad(set_to_none=True)
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=name == 'bf16_autocast'):
                y = om(x)
                hm = [[h.float() for h in hs] for hs in (om.xy_heatmaps, om.zy_heatmaps, om.xz_heatmaps)]
            om.xy_heatmaps, om.zy_heatmaps, om.xz_heatmaps = hm
            loss = D.average_loss(om.forward_3d_losses(y.float(), t), m)
            loss.backward()
            opt.step()
            return loss
        try:
            for _ in range(warmup):
                one()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = one()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {'value': batch / (ms * 1e-3), 'ms_per_step': ms, 'last_loss': loss.item()}
        except Exception as exc:       # comparator only: never fail the bench line because of it
            out[name] = {'error': '%s: %s' % (type(exc).__name__, str(exc)[:200])}
        del om, opt
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    return out


def run_torch_gpu(args):
    import torch
    if int(os.environ.get('RANK', '0')) != 0:
        return
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    base = gpu_library_baseline(args.batch or CFG['batch'], steps=max(1, min(args.steps, 20)), warmup=3)
    best = max((v['value'] for v in base.values() if isinstance(v, dict) and 'value' in v), default=None)
    print(json.dumps({'impl': 'torch-gpu', 'metric': METRIC, 'value': best, 'unit': 'images/s', 'n_gpus': 1,
                      'higher_is_better': True, 'data': 'synthetic', 'config': {'workload': CFG['workload']},
                      'gpu_library_baseline': base}))


def tail_sweep(peak_bw, sizes=(32, 64, 128), batch=128):
    """BASELINE.json configs[3]: the fused flat_softmax + dsnt (+ xyz combination) kernel, and the full fused tail
    (+ Gaussian targets, JS divergences, Euclidean loss), forward and backward, three planes, 17 joints, batch 128:
    achieved HBM GB/s = algorithmic bytes (8 per heatmap element forward: logit in, probability out; 12 backward:
    probability + upstream gradient in, d logit out) / CUDA-event time, inputs rotated through > 126 MB (L2)."""
    import torch
    from margipose_b200 import dsntnn as K
    B, J = batch, JOINTS
    out = []
    for S in sizes:
        n_sets = max(2, int(300e6 // (3 * B * J * S * S * 4)) + 1)
        sets = [[torch.randn(B, J, S, S, device='cuda') for _ in range(3)] for _ in range(n_sets)]
        probs = [[torch.empty_like(t) for t in s] for s in sets]
        gup = [[torch.randn(B, J, S, S, device='cuda') * 1e-3 for _ in range(3)] for _ in range(n_sets)]
        dz = [[torch.empty_like(t) for t in s] for s in sets]
        target = torch.rand(B, J, 3, device='cuda') * 1.6 - 0.8
        coords = torch.empty(B, J, 3, device='cuda')
        loss = torch.empty(B, J, device='cuda')
        w = torch.full((B, J), 1.0 / (B * J), device='cuda')
        modes = {
            'softmax_dsnt_fwd': (lambda i: K._tail_fwd(sets[i], True, prob=probs[i], coords=coords), 8),
            'full_tail_fwd': (lambda i: K._tail_fwd(sets[i], True, prob=probs[i], target=target, coords=coords,
                                                    loss=loss), 8),
            'full_tail_bwd': (lambda i: K._tail_bwd(probs[i], gup[i], dz[i], target=target, coords=coords, w=w,
                                                    project=True), 12),
        }
        row = {'heatmap': S}
        for name, (fn, bpe) in modes.items():
            for i in range(n_sets):
                fn(i)
            torch.cuda.synchronize()
            reps = 5 * n_sets
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for r in range(reps):
                    fn(r % n_sets)
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            gbs = 3 * B * J * S * S * bpe / us / 1e3
            row[name] = {'us': us, 'GB/s': gbs, 'frac': gbs / peak_bw}
        out.append(row)
        del sets, probs, gup, dz
        torch.cuda.empty_cache()
    return out


def precise_mode_bench(dev, batch, steps=8):
    """The same training step in the bf16x3 arithmetic (bf16 pairs, three tensor-core passes per convolution):
    the mode whose results match the fp32 reference to ~2e-4 (PARITY.md).  Device-resident inputs, CUDA events."""
    import torch
    from margipose_b200.models import create_model
    from margipose_b200.optim import FlatSGD
    from margipose_b200.train import TrainStep
    torch.manual_seed(0)
    model = create_model(DESC).set_precision('bf16x3').to(dev).train()
    opt = FlatSGD(model, lr=1e-3, momentum=0.9)
    step = TrainStep(model, opt, batch=batch, height=RES, width=RES)
    data = synthetic(batch, 2, seed=300, device=dev)
    for i in range(step.warmup + 3):
        step.load(*data[i % 2])
        step.run()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step.load(*data[i % 2])
        step.run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    out = {'precision': 'bf16x3', 'value': batch / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms, 'steps': steps,
           'batch': batch, 'last_loss': step.loss.item(),
           'note': 'matches the fp32 reference golden to ~2e-4 in the coordinates (PARITY.md section 2a)'}
    del step, opt, model
    torch.cuda.empty_cache()
    return out


def inference_bench(model, dev, batch, reps=50):
    """Eval-mode forward through InferStep (folded BatchNorm, captured graph): batch 1 latency (the reference's
    `margipose infer` / eval_3d.py batch size) and batch `batch` throughput; fp32 NCHW input copied from pinned
    host memory and the (B, 17, 3) result read back to the host inside the timed region, plus the device-only
    time of the captured forward.  The model keeps the running statistics the training steps above produced."""
    import torch
    from margipose_b200.infer import InferStep
    out = {'api': 'margipose_b200.infer.InferStep (model.eval(), BatchNorm folded into the conv epilogues, '
                  'one CUDA graph)', 'unit': 'images/s'}
    model.eval()
    for b in sorted({1, batch}):
        inf = InferStep(model, b, RES, RES)
        x = torch.randn(b, 3, RES, RES).pin_memory()
        for _ in range(5):
            inf(x).cpu()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            xyz = inf(x).cpu()
        wall = (time.perf_counter() - t0) / reps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            inf.run()
        e1.record()
        torch.cuda.synchronize(dev)
        gpu_ms = e0.elapsed_time(e1) / reps
        out['batch_%d' % b] = {'e2e_ms': 1e3 * wall, 'e2e_images_per_s': b / wall, 'gpu_ms': gpu_ms,
                               'gpu_images_per_s': b / (gpu_ms * 1e-3), 'launches': inf.launches(),
                               'finite': bool(torch.isfinite(xyz).all())}
        del inf
    model.train()
    return out


def cpu_forward_baseline():
    """BASELINE.json configs[0] / BASELINE.md section 4.4: batch-1 eval forward of the reference algorithm on CPU."""
    import torch
    from oracle import model_oracle as M
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    om = M.create_oracle(DESC).eval()
    x = torch.randn(1, 3, RES, RES)
    with torch.no_grad():
        om(x)
        t0 = time.perf_counter()
        for _ in range(3):
            om(x)
        ms = 1e3 * (time.perf_counter() - t0) / 3
    return {'batch_1_forward_ms': ms, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '3 eval forwards of batch 1 after 1 warm-up, fp32 oracle port on CPU'}


def cpu_tail_baseline(sizes=(32, 64, 128), batch=128):
    """the reference's tail (flat_softmax + heatmaps_to_coords + 3 x js_reg_losses +
    euclidean_losses, and its autograd backward) on the host cores at the sweep's shapes, as effective GB/s with the
    GPU sweep's accounting (8 B per heatmap element forward, 8 + 12 = 20 B forward + backward)."""
    import torch
    from oracle import dsnt_oracle as D
    torch.set_num_threads(os.cpu_count() or 1)
    out = {'cores': torch.get_num_threads(), 'kind': 'port', 'unit': 'GB/s',
           'note': 'fp32 oracle port of dsntnn.py on CPU, batch %d x %d joints x 3 planes, best of two calls after one '
                   'warm-up' % (batch, JOINTS)}
    g = torch.Generator().manual_seed(0)
    for S in sizes:
        z = [torch.randn(batch, JOINTS, S, S, generator=g).requires_grad_() for _ in range(3)]
        target = torch.rand(batch, JOINTS, 3, generator=g) * 1.6 - 0.8
        mask = torch.ones(batch, JOINTS)

        def fwd():
            p = [D.flat_softmax(t) for t in z]
            return D.average_loss(D.losses_3d([p[0]], [p[1]], [p[2]], target, 'jsd'), mask)
        fwd()
        tf = 1e9
        for _ in range(2):       # (with the autograd graph being recorded, as in the reference's training loop)
            t0 = time.perf_counter()
            fwd()
            tf = min(tf, time.perf_counter() - t0)
        row = {'fwd_ms': 1e3 * tf, 'fwd_GB/s': 3 * batch * JOINTS * S * S * 8 / tf / 1e9}
        if S <= 64:     # the autograd graph of the reference tail holds ~10 full-size temporaries per plane
            fwd().backward()
            tb = 1e9
            for _ in range(2):
                t0 = time.perf_counter()
                fwd().backward()
                tb = min(tb, time.perf_counter() - t0)
            row.update({'fwd_bwd_ms': 1e3 * tb, 'fwd_bwd_GB/s': 3 * batch * JOINTS * S * S * 20 / tb / 1e9})
        out['heatmap_%d' % S] = row
        del z
    return out


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from margipose_b200.models import create_model
    from margipose_b200.optim import FlatSGD
    from margipose_b200.train import TrainStep
    from margipose_b200 import parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, K, W = args.batch or CFG['batch'], args.steps, max(args.warmup, 3)

    if args.deterministic:
        from margipose_b200 import utils
        utils.init_algorithms(deterministic=True)
    torch.manual_seed(0)
    model = create_model(DESC).set_precision(args.precision).to(dev).train()
    opt = FlatSGD(model, lr=1e-3, momentum=0.9)
    if world > 1:
        parallel.sync_model(model)
    step = TrainStep(model, opt, batch=B, height=RES, width=RES, use_graph=not args.no_graph)
    dev_sets = synthetic(B, 4, seed=100 + rank, device=dev)
    host_sets = synthetic(B, 2, seed=200 + rank, pinned=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # warm-up: >= W steps, and enough for the CUDA graph to be captured and replayed once
    for i in range(max(W, step.warmup + 2)):
        step.load(*dev_sets[i % len(dev_sets)])
        step.run()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step.load(*dev_sets[i % len(dev_sets)])
        step.run()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    # the clock / throttle sampler covers the device-timed region; it is stopped here because a polling nvidia-smi
    # takes driver locks that delay launches, which the per-step synchronisation of the end-to-end loop below exposes
    clocks = sampler.stop() if rank == 0 else None

    # end to end: pinned host inputs every step, every step's loss read back.  As a prefetching loader would, the copy of
    # batch i + 1 is started (TrainStep.submit(..., prefetch=next batch): copy stream, staging slot) once step i has been
    # queued; as an asynchronous logger would, step i's loss (4 bytes into pinned host memory, queued behind the step) is
    # waited for after step i + 1 has been queued, so the GPU does not idle while the host turns around.  Every step's
    # inputs cross PCIe and every step's loss is read on the host inside the timed region.
    # (warm-up through the same path: the first prefetch creates the copy stream and the staging slots, the first submit
    # the pinned loss ring; the last warm-up step stages the first timed batch)
    step.prefetch(*host_sets[1])
    for i in range(3):
        step(*host_sets[(i + 1) % len(host_sets)], prefetch=host_sets[i % len(host_sets)])
    barrier()
    t0 = time.perf_counter()
    last = pending = None
    for i in range(K):
        # (the first timed batch was staged by the warm-up, so the last step stages one more: K PCIe copies in the region)
        nxt = host_sets[(i + 1) % len(host_sets)]
        queued = step.submit(*host_sets[i % len(host_sets)], prefetch=nxt)
        if pending is not None:
            last = pending.item()
        pending = queued
    last = pending.item()
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = e2e_s.item()
    h2d = sum(t.numel() * t.element_size() for t in host_sets[0])
    d2h = 4

    # roofline of the dominant kernel: every mp_conv_igemm launch of one step (fprop + dgrad), each
    # alone on the GPU, in program order, replayed from a CUDA graph and timed with CUDA events
    roof = None
    if rank == 0:
        eng = model.engine_for(B, RES, RES, True)
        peak_tf, peak_bw, src = peaks()
        classes = {}
        for name in ('mp_conv_igemm', 'mp_conv_wgrad', 'mp_bn_fwd', 'mp_bn_bwd_reduce', 'mp_bn_bwd_apply'):
            classes[name] = eng.time_kernel_class(name)
        n, t_ms, fl = classes['mp_conv_igemm']
        achieved = fl / (t_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'igemm_traffic.json')
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get('dram_bytes_per_launch')
        wn, wt, wf = classes['mp_conv_wgrad']
        roof = {'bound': 'tensor', 'kernel': 'igemm_kernel (mp_conv_igemm: conv fprop + dgrad)', 'achieved': achieved,
                'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'traffic': traffic,
                'traffic_note': 'DRAM bytes (read + write) of the dominant grouped launch, 3 x (128->128 3x3 @32x32) forward + BN '
                                'statistics, ncu --set full (profiles/r01_igemm_ncu.md launch 0); algorithmic bytes 51.2 MB, the '
                                'bf16 outputs stay in the 126 MB L2',
                'peak_source': src + ' bf16 sustained (MEASURED_PEAKS.json)', 'launches_per_step': n,
                'avg_launch_us': 1e3 * t_ms / n, 'flops_per_launch': fl / n,
                'serial_ms_per_step': {k: v[1] for k, v in classes.items()},
                'wgrad_tflops': wf / (wt * 1e-3) / 1e12 if wn else None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tail = None
    if world == 1 and not args.skip_tail:
        tail = {'bound': 'hbm', 'unit': 'GB/s', 'peak': peak_bw, 'peak_source': src + ' copy bandwidth (MEASURED_PEAKS.json)',
                'workload': 'configs[3]: heatmap 32/64/128, 17 joints, batch 128, three planes',
                'sweep': tail_sweep(peak_bw)}
        if not args.skip_cpu:
            tail['cpu_baseline'] = cpu_tail_baseline()
    infer = None
    if world == 1 and not args.skip_infer:
        infer = inference_bench(model, dev, B)
        if not args.skip_cpu:
            infer['cpu_baseline'] = cpu_forward_baseline()
    precise = None
    if world == 1 and not args.skip_precise and args.precision == 'bf16':
        precise = precise_mode_bench(dev, B)
    cpu = None
    if world == 1 and not args.skip_cpu:
        cpu = cpu_reference(batch=CFG['cpu_batch'], steps=3, warmup=1, budget_s=60.0)
        cpu = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    n_launches = step.launches_per_step() * K * 2
    gpu_lib = None
    if world == 1 and not args.skip_gpu_lib:
        del step, eng
        model.drop_engines()
        torch.cuda.empty_cache()
        gpu_lib = gpu_library_baseline(B)
    value = B * world * K / (ms * 1e-3)
    line = {'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.precision, 'data': 'synthetic',
            'config': {'workload': '%s, batch %d per GPU, fwd + 3D loss + bwd + SGD-momentum step'
                                   % (CFG['workload'], B),
                       'global_batch': B * world, 'parallelism': 'dp%d' % world,
                       'cuda_graph': not args.no_graph, 'deterministic': bool(args.deterministic),
                       'l2': 'per-step working set (~10 GB of activations) exceeds the 126 MB L2; 4 input sets rotate'},
            'conv_flops_per_image': FLOPS_PER_IMAGE,
            'conv_tflops_whole_step': value / world * FLOPS_PER_IMAGE / 1e12,
            'roofline': roof, 'tail_roofline': tail, 'cpu_baseline': cpu, 'gpu_library_baseline': gpu_lib,
            'inference': infer, 'precise_mode': precise,
            'e2e': {'value': B * world * K / e2e_s, 'unit': 'images/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'last_loss': last,
                    'note': 'TrainStep.submit per step (the call behind TrainStep.__call__), loss of step i read after step i + 1 is queued; the next batch is announced with TrainStep.prefetch (pinned '
                            'host -> staging slot on a copy stream), so its PCIe copy overlaps the running step'},
            'gpu_launches': n_launches, 'clocks': clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'torch-gpu'])
    ap.add_argument('--config', default='r34x4_256', choices=sorted(CONFIGS))
    ap.add_argument('--batch', type=int, default=0, help='images per GPU per step (default: the config\'s)')
    ap.add_argument('--skip-gpu-lib', action='store_true', help='skip the stock-PyTorch-on-GPU comparator')
    ap.add_argument('--skip-infer', action='store_true', help='skip the eval-mode inference measurement')
    ap.add_argument('--skip-precise', action='store_true', help='skip the bf16x3 (fp32-grade) mode measurement')
    ap.add_argument('--deterministic', action='store_true',
                    help='bit-wise reproducible mode (margipose_b200.utils.init_algorithms(deterministic=True))')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'bf16x3'],
                    help='engine arithmetic of the measured step (bf16x3: bf16 pairs, fp32-grade results)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-tail', action='store_true', help='skip the soft-argmax fusion HBM sweep (configs[3])')
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == 'reference':
        run_reference(args)
    elif args.impl == 'torch-gpu':
        run_torch_gpu(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
