"""Per-shape table of the conv launches of one training step: every mp_conv_igemm / mp_conv_wgrad op of the engine's
programs is replayed alone from a CUDA graph (warm, back to back) and grouped by geometry.  Shows where the time of the
kernel class goes relative to the tensor peak:

    python tools/igemm_table.py [mp_conv_igemm|mp_conv_wgrad] [tunable=value ...]
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from margipose_b200.models import create_model  # noqa: E402
from margipose_b200._lib import lib  # noqa: E402
import bench  # noqa: E402

name = 'mp_conv_igemm'
for kv in sys.argv[1:]:
    if '=' not in kv:
        name = kv
        continue
    k, v = kv.split('=')
    assert lib().mp_set_tunable(k.encode(), int(v)) == 0, kv
PEAK = bench.peaks()[0] if hasattr(bench, 'peaks') else 1374.4
torch.manual_seed(0)
model = create_model(bench.DESC).cuda().train()
x = torch.randn(32, 3, 256, 256, device='cuda')
model(x)
eng = model.engine_for(32, 256, 256, True)
torch.cuda.synchronize()

ops = []
for segs in (eng.fwd, eng.bwd):
    for kind, body in segs:
        for lane in ([body] if kind == 'serial' else body):
            ops += [op for op in lane if getattr(op, 'name', None) == name]


def time_op(op, reps=8, rounds=3):
    g = torch.cuda.CUDAGraph()
    op()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps):
            op()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * rounds)


def signature(op):
    parts = getattr(op, 'parts', None)
    a = parts[0].args if parts else op.args
    n = len(parts) if parts else 1
    if name == 'mp_conv_igemm':
        srcs = len({a.taps[i].src for i in range(a.n_taps)})
        return '%dx [%d img %3dx%-3d] N=%-3d taps=%-2d cblk=%d src=%d%s%s' % (
            n, a.n_img, a.out_h, a.out_w, a.w_rows, a.n_taps, a.cblocks, srcs,
            ' +res' if a.res else '', ' +stats' if a.stat_sum else '')
    return '%dx %s' % (n, ' '.join('%s=%s' % (f, getattr(a, f)) for f, _t in a._fields_
                                  if f in ('n_img', 'grid_h', 'grid_w', 'm_real', 'n_real', 'n_cols', 'n_off', 'n_taps')))


rows = collections.OrderedDict()
for op in ops:
    us = time_op(op)
    r = rows.setdefault(signature(op), [0, 0.0, 0.0])
    r[0] += 1
    r[1] += us
    r[2] += op.flops
tot_us = sum(r[1] for r in rows.values())
tot_fl = sum(r[2] for r in rows.values())
print('%s: %d launches, %.0f us warm and alone, %.2f TFLOP -> %.0f TFLOP/s (%.2f of %.0f)' % (
    name, len(ops), tot_us, tot_fl / 1e12, tot_fl / tot_us / 1e6, tot_fl / tot_us / 1e6 / PEAK, PEAK))
print('%-72s %4s %8s %7s %8s %6s %8s' % ('geometry', 'n', 'us', 'share', 'us/launch', 'frac', 'lost us'))
for sig, (n, us, fl) in sorted(rows.items(), key=lambda kv: -(kv[1][1] - kv[1][2] / PEAK / 1e6)):
    ideal = fl / PEAK / 1e6
    print('%-72s %4d %8.1f %6.1f%% %8.1f %6.2f %8.1f' % (sig, n, us, 100 * us / tot_us, us / n, ideal / us, us - ideal))
