"""Serial time of the BatchNorm kernel classes of one training step under a set of tunables:
    python tools/bn_time.py bn_fwd_minb=3 ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200.models import create_model
from margipose_b200._lib import lib
import bench

for kv in sys.argv[1:]:
    k, v = kv.split('=')
    assert lib().mp_set_tunable(k.encode(), int(v)) == 0, kv
torch.manual_seed(0)
model = create_model(bench.DESC).cuda().train()
x = torch.randn(32, 3, 256, 256, device='cuda')
model(x)
eng = model.engine_for(32, 256, 256, True)
torch.cuda.synchronize()
out = {n: eng.time_kernel_class(n, reps=5)[1] for n in ('mp_bn_fwd', 'mp_bn_bwd_reduce', 'mp_bn_bwd_apply')}
print('%-50s' % (' '.join(sys.argv[1:]) or '(defaults)'), ' '.join('%s %.3f' % (k[3:], v) for k, v in out.items()),
      'sum %.3f ms' % sum(out.values()))
