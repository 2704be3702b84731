"""Per-shape timing of the tcgen05 conv kernels (CUDA events, L2-warm like inside a training step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200 import convops as C
from margipose_b200._lib import lib

SHAPES = [  # cin, cout, k, stride, transposed, n, h, w
    (128, 128, 3, 1, False, 32, 32, 32),
    (192, 192, 3, 1, False, 32, 16, 16),
    (128, 128, 1, 1, False, 32, 32, 32),
    (192, 192, 1, 1, False, 32, 16, 16),
    (128, 192, 3, 2, False, 32, 32, 32),
    (192, 128, 3, 2, True, 32, 16, 16),
    (64, 64, 3, 1, False, 32, 64, 64),
    (128, 17, 3, 1, False, 32, 32, 32),
]
# tunable sets: "a=1 b=2 / a=0" runs the sweep once per set (later sets keep earlier values unless overridden)
SETS = [x.split() for x in ' '.join(sys.argv[1:]).split('/')] if len(sys.argv) > 1 else [[]]
which = os.environ.get('WHICH', 'fwd,dgrad,wgrad').split(',')
iters = int(os.environ.get('ITERS', '20'))
only = os.environ.get('ONLY')

def timeit(fn):
    """GPU time per call: `iters` back-to-back calls captured in a CUDA graph (no host overhead)."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    if os.environ.get('NOGRAPH'):
        return 1.0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for tset in SETS:
    for kv in tset:
        k_, v_ = kv.split('=')
        assert lib().mp_set_tunable(k_.encode(), int(v_)) == 0, kv
    if len(SETS) > 1:
        print('== ' + ' '.join(tset))
    for idx, (cin, cout, k, s, tr, n, h, w) in enumerate(SHAPES):
        if only is not None and idx != int(only): continue
        g = C.ConvGeom(cin, cout, k, s, tr)
        ho, wo = g.out_hw(h, w)
        x = torch.randn(n, h, w, g.cin_p, device='cuda').to(torch.bfloat16)
        dy = torch.randn(n, ho, wo, g.cout_p, device='cuda').to(torch.bfloat16)
        master = torch.randn(g.master_shape, device='cuda') * 0.05
        wf, wb = C.pack_fwd(g, master), C.pack_bwd(g, master)
        y = torch.zeros(n, ho, wo, g.cout_p, device='cuda', dtype=torch.bfloat16)
        dx = torch.zeros_like(x)
        dw = torch.zeros(g.master_shape, device='cuda')
        stats = torch.zeros(2, g.cout_p, device='cuda')
        flops = 2.0 * n * (h * w if tr else ho * wo) * g.taps * cin * cout
        out = '%-28s' % str((cin, cout, k, s, 'T' if tr else '', n, h, w))
        if 'fwd' in which:
            t = timeit(lambda: C.conv_forward(g, x, wf, y, stats=(stats[0], stats[1])))
            out += ' fwd %7.1f us %6.0f TF' % (t, flops / t / 1e6)
            t = timeit(lambda: C.conv_forward(g, x, wf, y))
            out += ' (nostat %6.1f us %5.0f TF)' % (t, flops / t / 1e6)
        if 'dgrad' in which:
            t = timeit(lambda: C.conv_dgrad(g, dy, wb, dx))
            out += ' dgrad %7.1f us %6.0f TF' % (t, flops / t / 1e6)
        if 'wgrad' in which:
            t = timeit(lambda: C.conv_wgrad(g, x, dy, dw))
            out += ' wgrad %7.1f us %6.0f TF' % (t, flops / t / 1e6)
        print(out)
