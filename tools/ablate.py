"""Where does the step time go?  Replays CUDA graphs of the forward / backward programs with
selected kernel classes removed (numerics become garbage; timing stays meaningful)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200.models import create_model
from margipose_b200._lib import lib
import bench

for kv in sys.argv[1:]:
    k, v = kv.split('=')
    lib().mp_set_tunable(k.encode(), int(v))
B = int(os.environ.get('BATCH', '32'))
torch.manual_seed(0)
model = create_model(bench.DESC).cuda().train()
x = torch.randn(B, 3, 256, 256, device='cuda')
model(x)
eng = model.engine_for(B, 256, 256, True)
torch.cuda.synchronize()

def filt(segs, drop):
    out = []
    for kind, body in segs:
        if kind == 'serial':
            out.append((kind, [op for op in body if getattr(op, 'name', 'tail') not in drop]))
        else:
            out.append((kind, [[op for op in lane if getattr(op, 'name', 'tail') not in drop] for lane in body]))
    return out

def serial(segs):
    ops = []
    for kind, body in segs:
        for lane in ([body] if kind == 'serial' else body):
            ops += lane
    return [('serial', ops)]

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

print('launches fwd %d bwd %d' % (eng.launches(eng.fwd), eng.launches(eng.bwd)))
print('fwd                      %.2f ms   (serialised lanes: %.2f)' % (time_graph(eng.fwd), time_graph(serial(eng.fwd))))
print('bwd                      %.2f ms   (serialised lanes: %.2f)' % (time_graph(eng.bwd), time_graph(serial(eng.bwd))))
for drop in (['mp_bn_fwd'], ['mp_conv_igemm'],):
    print('fwd without %-22s %.2f ms' % (drop, time_graph(filt(eng.fwd, drop))))
for drop in (['mp_conv_wgrad'], ['mp_bn_bwd_reduce', 'mp_bn_bwd_apply'], ['mp_conv_igemm'], ['mp_combiner_bwd'],
             ['mp_conv_wgrad', 'mp_conv_igemm'], ['mp_conv_wgrad', 'mp_bn_bwd_reduce', 'mp_bn_bwd_apply', 'mp_combiner_bwd']):
    print('bwd without %-60s %.2f ms' % (drop, time_graph(filt(eng.bwd, drop))))
only = lambda segs, keep: filt(segs, set(n for n in ['mp_conv_igemm', 'mp_conv_wgrad', 'mp_bn_fwd', 'mp_bn_bwd_reduce', 'mp_bn_bwd_apply', 'mp_combiner_bwd', 'mp_combiner_fwd', 'mp_stem_im2col', 'mp_maxpool_fwd', 'mp_maxpool_bwd', 'mp_axis_permute', 'mp_add_bf16', 'tail']) - set(keep))
for keep in (['mp_conv_igemm'], ['mp_bn_fwd']):
    print('fwd only %-25s %.2f ms  serial %.2f' % (keep, time_graph(only(eng.fwd, keep)), time_graph(serial(only(eng.fwd, keep)))))
for keep in (['mp_conv_igemm'], ['mp_conv_wgrad'], ['mp_bn_bwd_reduce'], ['mp_bn_bwd_apply']):
    print('bwd only %-25s %.2f ms  serial %.2f' % (keep, time_graph(only(eng.bwd, keep)), time_graph(serial(only(eng.bwd, keep)))))
