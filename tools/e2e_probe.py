"""Per-step wall time of TrainStep.__call__ fed from pinned host memory (what bench.py's e2e measures), with and
without prefetch, to see where host-side time goes:  python tools/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from margipose_b200.models import create_model
from margipose_b200.optim import FlatSGD
from margipose_b200.train import TrainStep

bench.select_config(os.environ.get('CONFIG', 'r34x4_256'))
B, RES = bench.CFG['batch'], bench.RES
torch.manual_seed(0)
model = create_model(bench.DESC).cuda().train()
opt = FlatSGD(model, lr=1e-3, momentum=0.9)
step = TrainStep(model, opt, batch=B, height=RES, width=RES)
host = bench.synthetic(B, 2, seed=1, pinned=True)
dev_sets = bench.synthetic(B, 2, seed=2, device=torch.device('cuda'))
for i in range(8):
    step(*host[i % 2])
torch.cuda.synchronize()
for mode in ('plain', 'prefetch-before', 'prefetch-after'):
    ts = []
    if mode != 'plain':
        step.prefetch(*host[0])
    for i in range(30):
        t0 = time.perf_counter()
        if mode == 'prefetch-before':
            step.prefetch(*host[(i + 1) % 2])
            step(*host[i % 2])
        elif mode == 'prefetch-after':
            step(*host[i % 2], prefetch=host[(i + 1) % 2])
        else:
            step(*host[i % 2])
        ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    print('%-16s median %.2f  min %.2f  p90 %.2f  max %.2f ms' % (mode, ts[15], ts[0], ts[27], ts[-1]))
# device-resident steps and the submit pipeline (what bench.py times)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(30):
    step.load(*dev_sets[i % 2])
    step.run()
torch.cuda.synchronize()
print('device-resident  %.2f ms per step' % (1e3 * (time.perf_counter() - t0) / 30))
for label, pre in (('submit+prefetch ', True), ('submit, no prefetch', False)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pending = None
    if pre:
        step.prefetch(*host[0])
    for i in range(30):
        q = step.submit(*host[i % 2], prefetch=host[(i + 1) % 2] if (pre and i + 1 < 30) else None)
        if pending is not None:
            pending.item()
        pending = q
    pending.item()
    torch.cuda.synchronize()
    print('%s %.2f ms per step' % (label, 1e3 * (time.perf_counter() - t0) / 30))
# host-only cost of the calls (GPU idle is not counted: sync first, time the enqueue part)
torch.cuda.synchronize()
t0 = time.perf_counter(); step.prefetch(*host[0]); t1 = time.perf_counter()
step._take_staged(host[0][0]); t2 = time.perf_counter(); step.run(); t3 = time.perf_counter()
torch.cuda.synchronize()
print('host enqueue: prefetch %.3f ms, take_staged %.3f ms, run (graph launches) %.3f ms' % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
