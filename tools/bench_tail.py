"""BASELINE.json configs[3]: flat_softmax + dsnt (+ xyz combination + JS + Euclid) fusion sweep --
heatmap side 32/64/128, 17 joints, batch 128, three planes -- achieved HBM GB/s vs the measured peak.
Algorithmic bytes per heatmap element per plane: forward 8 (read logit, write probability);
backward 12 (read probability, read upstream gradient, write d logit)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200 import dsntnn as K

peak = 6449.1
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peak = json.load(open(p)).get('hbm_gbs', peak)
B, J = 128, 17
rows = []
for S in (32, 64, 128):
    n_sets = max(2, int(300e6 // (3 * B * J * S * S * 4)) + 1)      # rotate > 126 MB of inputs: defeat L2
    sets = [[torch.randn(B, J, S, S, device='cuda') for _ in range(3)] for _ in range(n_sets)]
    probs = [[torch.empty_like(t) for t in s] for s in sets]
    gup = [[torch.randn(B, J, S, S, device='cuda') * 1e-3 for _ in range(3)] for _ in range(n_sets)]
    dz = [[torch.empty_like(t) for t in s] for s in sets]
    target = torch.rand(B, J, 3, device='cuda') * 1.6 - 0.8
    coords = torch.empty(B, J, 3, device='cuda'); loss = torch.empty(B, J, device='cuda'); w = torch.full((B, J), 1.0 / (B * J), device='cuda')
    def fwd(i):
        K._tail_fwd(sets[i], True, prob=probs[i], target=target, coords=coords, loss=loss)
    def bwd(i):
        K._tail_bwd(probs[i], gup[i], dz[i], target=target, coords=coords, w=w, project=True)
    out = {}
    for name, fn, bytes_per_elem in (('fwd', fwd, 8), ('bwd', bwd, 12)):
        for i in range(n_sets): fn(i)
        torch.cuda.synchronize()
        reps = 5 * n_sets
        graph = torch.cuda.CUDAGraph()          # graph replay: GPU time, not Python call overhead
        with torch.cuda.graph(graph):
            for r in range(reps): fn(r % n_sets)
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        gbs = 3 * B * J * S * S * bytes_per_elem / us / 1e3
        out[name] = (us, gbs)
    rows.append((S, out))
    print('S=%3d  fwd %8.1f us %7.0f GB/s (%.2f of measured peak)   bwd %8.1f us %7.0f GB/s (%.2f)   [%d rotating input sets]'
          % (S, out['fwd'][0], out['fwd'][1], out['fwd'][1] / peak, out['bwd'][0], out['bwd'][1], out['bwd'][1] / peak, n_sets))
