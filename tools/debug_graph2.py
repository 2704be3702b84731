import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200.models import create_model
from margipose_b200.optim import FlatSGD
from margipose_b200 import dsntnn as K
desc = {'type': 'margipose', 'version': '6.0.1', 'settings': dict(n_stages=2, feature_extractor='resnet18')}
torch.manual_seed(0)
model = create_model(desc).cuda().train()
opt = FlatSGD(model, lr=1e-3, momentum=0.9)
x = torch.randn(2, 3, 256, 256, device='cuda'); t = torch.rand(2, 17, 3, device='cuda'); m = torch.ones(2, 17, device='cuda')
for i in range(3):
    opt.zero_grad()
    out = model(x)
    l = K.average_loss(model.forward_3d_losses(out, t), m)
    l.backward(); opt.step()
torch.cuda.synchronize()
eng = model.engine_for(2, 256, 256, True)
def try_capture(name, fn):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        return True
    except Exception as e:
        print(name, 'FAILED', str(e).split('\n')[0][:160]); torch.cuda.synchronize()
        return False
print('whole bwd program (main thread):', try_capture('bwd', lambda: eng._run(eng.bwd)))
for si, (kind, body) in enumerate(eng.bwd):
    ok = try_capture('seg%d' % si, lambda: eng._run([(kind, body)]))
    print('segment', si, kind, 'ok' if ok else 'FAILED')
    if not ok:
        lanes = [body] if kind == 'serial' else body
        for li, lane in enumerate(lanes):
            for oi, op in enumerate(lane):
                if not try_capture('op', op):
                    print('   first failing op: lane', li, 'index', oi, getattr(op, 'name', 'tail'))
                    a = getattr(op, 'args', None)
                    if a is not None and hasattr(a, 'n_taps'):
                        print('   n_taps', a.n_taps, [getattr(a, f, None) for f in ('m_real', 'n_real', 'n_cols', 'n_img', 'grid_h', 'grid_w', 'out_h', 'out_w', 'w_rows', 'w_k')])
                    break
            else:
                continue
            break
        break
