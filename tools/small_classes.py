"""Serial (each class alone, CUDA-graph replay) time per step of the non-conv, non-BatchNorm kernel classes:
    python tools/small_classes.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200.models import create_model
import bench

torch.manual_seed(0)
model = create_model(bench.DESC).cuda().train()
x = torch.randn(32, 3, 256, 256, device='cuda')
model(x)
eng = model.engine_for(32, 256, 256, True)
torch.cuda.synchronize()
names = set()
for segs in (eng.fwd, eng.bwd):
    for kind, body in segs:
        for lane in ([body] if kind == 'serial' else body):
            names |= {getattr(op, 'name', 'tail') for op in lane}
tot = 0.0
for n in sorted(names):
    k, ms, _ = eng.time_kernel_class(n, reps=5)
    tot += ms
    print('%-22s %4d launches %8.1f us  (%.1f us each)' % (n, k, 1e3 * ms, 1e3 * ms / max(k, 1)))
print('sum %.2f ms' % tot)
