import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden.make_golden import model_inputs
from margipose_b200.models import create_model
from margipose_b200.optim import FlatSGD
from margipose_b200 import dsntnn as K

desc = {'type': 'margipose', 'version': '6.0.1', 'settings': dict(n_stages=1, feature_extractor='resnet18')}
torch.manual_seed(5)
ms = [create_model(desc) for _ in range(3)]
for m in ms[1:]:
    m.load_state_dict(ms[0].state_dict())
for m in ms:
    m.cuda().train()
x, target, mask = model_inputs(6, 2)
x, target, mask = x.cuda(), target.cuda(), mask.cuda()
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
opts = [FlatSGD(ms[0], lr=lr, momentum=0.9, weight_decay=1e-4), FlatSGD(ms[1], lr=lr, momentum=0.9, weight_decay=1e-4)]
ms[2](x)
opts.append(torch.optim.SGD(ms[2].parameters(), lr=lr, momentum=0.9, weight_decay=1e-4))
def rel(a, b): return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
for step in range(4):
    ls = []
    for m, o in zip(ms, opts):
        o.zero_grad()
        out = m(x)
        l = K.average_loss(m.forward_3d_losses(out, target), mask)
        l.backward()
        ls.append(l.item())
    g = [torch.cat([p.grad.flatten() for p in m.parameters()]) for m in ms]
    print('step', step, 'loss', ls, 'grad rel flat-vs-flat', rel(g[0], g[1]), 'flat-vs-torch', rel(g[0], g[2]), 'gnorm', g[0].norm().item())
    for o in opts: o.step()
    w = [torch.cat([p.detach().flatten() for p in m.parameters()]) for m in ms]
    print('        param rel flat-vs-flat', rel(w[0], w[1]), 'flat-vs-torch', rel(w[0], w[2]))
