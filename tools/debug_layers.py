"""Layer-wise parity trace: CUDA engine vs the bf16-emulating oracle (run on the GPU box)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model_oracle as M
from oracle import dsnt_oracle as D
from tests.golden.make_golden import model_inputs
from margipose_b200.models import create_model

fe = sys.argv[1] if len(sys.argv) > 1 else 'resnet18'
stages = int(sys.argv[2]) if len(sys.argv) > 2 else 1
batch = 2
desc = {'type': 'margipose', 'version': '6.0.1',
        'settings': dict(n_stages=stages, feature_extractor=fe, axis_permutation=True, pixelwise_loss='jsd')}
torch.manual_seed(31)
om = M.create_oracle(desc, emulate_bf16=True).train()
model = create_model(desc)
model.load_state_dict(om.state_dict())
model.cuda().train()
x, target, mask = model_inputs(32, batch)
om.nm.trace = []
out_o = om(x)
out = model(x.cuda())
eng = model.engine_for(batch, 256, 256, True)
assert len(eng.trace) == len(om.nm.trace), (len(eng.trace), len(om.nm.trace))
for (name, buf, c), want in zip(eng.trace, om.nm.trace):
    if buf.dtype == torch.bfloat16:
        got = buf[..., :c].float().cpu().permute(0, 3, 1, 2)
    else:
        got = buf.cpu()
    err = ((got - want).norm() / want.norm()).item()
    print('%-28s rel L2 %.3e  max|d| %.3e  |want| %.3e' % (name, err, (got - want).abs().max().item(), want.abs().max().item()))
