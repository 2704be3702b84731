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
def full():
    opt.zero_grad()
    out = model(x)
    l = K.average_loss(model.forward_3d_losses(out, t), m)
    l.backward()
for i in range(3):
    full(); opt.step()
torch.cuda.synchronize()
eng = model.engine_for(2, 256, 256, True)
def try_capture(name, fn):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, 'OK')
    except Exception as e:
        print(name, 'FAILED', str(e).split('\n')[0][:160]); torch.cuda.synchronize()
leaf = [[p.detach().clone().requires_grad_() for p in row] for row in eng.probs]
def tail_only():
    losses = 0
    for row in leaf:
        l, _ = K.fused_tail_losses(row[0], row[1], row[2], t)
        losses = losses + l
    K.average_loss(losses, m).backward()
tail_only(); torch.cuda.synchronize()
try_capture('tail-only backward', tail_only)
def plain_torch():
    a = torch.ones(4, device='cuda', requires_grad=True)
    (a * 2).sum().backward()
plain_torch()
try_capture('plain torch backward', plain_torch)
real_bwd = eng.backward
calls = []
eng.backward = lambda grads: calls.append(1)
try_capture('full with no-op engine backward', full)
def only_copy(grads):
    for tt, row in enumerate(eng.gin):
        for k, g in enumerate(row):
            if grads[tt][k] is None: g.zero_()
            else: g.copy_(grads[tt][k])
eng.backward = only_copy
try_capture('full with gin copies only', full)
def serial_bwd(grads):
    only_copy(grads)
    for kind, body in eng.bwd:
        for lane in ([body] if kind == 'serial' else body):
            for op in lane: op()
eng.backward = serial_bwd
try_capture('full with serial engine backward (no side streams)', full)
eng.backward = real_bwd
try_capture('full', full)
