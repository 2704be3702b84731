"""Forward / backward program time of the bench workload under a set of kernel tunables:
    python tools/step_time.py igemm_pair=1 wgrad_ctas=64 ...
(CUDA-graph replay of the engine's forward and backward launch programs, 3 column lanes + aux streams.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200.models import create_model
from margipose_b200._lib import lib
import bench

for kv in sys.argv[1:]:
    k, v = kv.split('=')
    assert lib().mp_set_tunable(k.encode(), int(v)) == 0, kv
B = int(os.environ.get('BATCH', '32'))
torch.manual_seed(0)
model = create_model(bench.DESC)
if os.environ.get('PRECISION'):
    model.set_precision(os.environ['PRECISION'])
model = model.cuda().train()
x = torch.randn(B, 3, 256, 256, device='cuda')
model(x)
eng = model.engine_for(B, 256, 256, True)
torch.cuda.synchronize()


def time_graph(segs, reps=5):
    eng._run(segs); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng._run(segs)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


f, b = time_graph(eng.fwd), time_graph(eng.bwd)
cls = {n: eng.time_kernel_class(n)[1] for n in ('mp_conv_igemm', 'mp_conv_wgrad', 'mp_bn_fwd', 'mp_bn_bwd_reduce', 'mp_bn_bwd_apply')}
print('%-40s fwd %6.2f  bwd %6.2f  sum %6.2f ms | serial igemm %.2f wgrad %.2f bn %.2f / %.2f / %.2f' % (
    ' '.join(sys.argv[1:]) or '(defaults)', f, b, f + b, cls['mp_conv_igemm'], cls['mp_conv_wgrad'], cls['mp_bn_fwd'],
    cls['mp_bn_bwd_reduce'], cls['mp_bn_bwd_apply']))
