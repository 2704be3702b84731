"""Workload for the ncu capture of the fused tail kernels (BASELINE.json configs[3]: heatmap 32/64/128, 17 joints,
batch 128, three planes): per size two warm-up launches and ONE launch each of
  softmax + dsnt + xyz (forward, no loss), the full fused forward (+ Gaussian, JS x3, Euclid), the full fused backward.

    ncu --set full --clock-control none --import-source on -k regex:tail_ -o gpurun_out/r02_tail python tools/ncu_tail.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from margipose_b200 import dsntnn as K  # noqa: E402

B, J = 128, 17
for S in (32, 64, 128):
    z = [torch.randn(B, J, S, S, device='cuda') for _ in range(3)]
    p = [torch.empty_like(t) for t in z]
    gup = [torch.randn(B, J, S, S, device='cuda') * 1e-3 for _ in range(3)]
    dz = [torch.empty_like(t) for t in z]
    target = torch.rand(B, J, 3, device='cuda') * 1.6 - 0.8
    coords = torch.empty(B, J, 3, device='cuda')
    loss = torch.empty(B, J, device='cuda')
    w = torch.full((B, J), 1.0 / (B * J), device='cuda')
    flush = torch.empty(64 << 20, device='cuda')      # 256 MB > L2 between launches
    for fn in (lambda: K._tail_fwd(z, True, prob=p, coords=coords),
               lambda: K._tail_fwd(z, True, prob=p, target=target, coords=coords, loss=loss),
               lambda: K._tail_bwd(p, gup, dz, target=target, coords=coords, w=w, project=True)):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
print('done')
