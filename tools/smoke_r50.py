"""BASELINE.json configs[4] shape check: ResNet-50 stem, 5 stages, 384x384 (48x48 heatmaps, 24^3 mid volume)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200.models import create_model
from margipose_b200.optim import FlatSGD
from margipose_b200.train import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
desc = {'type': 'margipose', 'version': '6.0.1', 'settings': dict(n_stages=5, feature_extractor='resnet50', axis_permutation=True, pixelwise_loss='jsd')}
torch.manual_seed(0)
model = create_model(desc).cuda().train()
opt = FlatSGD(model, lr=1e-3, momentum=0.9)
step = TrainStep(model, opt, batch=B, height=384, width=384)
x = torch.randn(B, 3, 384, 384, device='cuda'); t = torch.rand(B, 17, 3, device='cuda') * 1.6 - 0.8
losses = [step(x, t) for _ in range(6)]
print('losses', losses, 'heatmap', tuple(model.xy_heatmaps[-1].shape))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step.load(x, t); step.run()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
eng = model.engine_for(B, 384, 384, True)
print('R50x5 @384 batch %d: %.2f ms/step, %.1f img/s, activations %.1f GB' % (B, dt * 1e3, B / dt, eng.activation_bytes() / 1e9))
