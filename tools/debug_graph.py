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
def fwd():
    return model(x)
def fwd_loss():
    out = model(x)
    return K.average_loss(model.forward_3d_losses(out, t), m)
def full():
    opt.zero_grad()
    l = fwd_loss()
    l.backward()
    return l
for i in range(3):
    full(); opt.step()
torch.cuda.synchronize()
for name, fn, mode in [('engine-forward-only', lambda: model.engine_for(2, 256, 256, True).forward(x), 'global'),
                       ('fwd', fwd, 'global'), ('fwd_loss', fwd_loss, 'global'), ('full', full, 'global'),
                       ('full-threadlocal', full, 'thread_local'), ('sgd', opt.step, 'global')]:
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            r = fn()
        g.replay(); torch.cuda.synchronize()
        print(name, 'OK')
    except Exception as e:
        print(name, 'FAILED', type(e).__name__, str(e).split('\n')[0][:200])
        torch.cuda.synchronize()
