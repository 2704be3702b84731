"""Timeline of one persistent CTA of mp_conv_igemm (tunable igemm_trace: CTA (0,0,0) stamps %globaltimer).
    python tools/trace_igemm.py [batch] [tunable=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from margipose_b200 import convops as C
from margipose_b200._lib import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
for kv in sys.argv[1:]:
    if '=' in kv:
        k, v = kv.split('=')
        assert lib().mp_set_tunable(k.encode(), int(v)) == 0
for (cin, cout, k, h) in ((128, 128, 3, 32), (192, 192, 3, 16)):
    g = C.ConvGeom(cin, cout, k, 1, False)
    x = torch.randn(n, h, h, g.cin_p, device='cuda').to(torch.bfloat16)
    master = torch.randn(g.master_shape, device='cuda') * 0.05
    wf = C.pack_fwd(g, master)
    y = torch.zeros(n, h, h, g.cout_p, device='cuda', dtype=torch.bfloat16)
    stats = torch.zeros(2, g.cout_p, device='cuda')
    trace = torch.zeros(64, dtype=torch.int64, device='cuda')
    for with_stats in (False, True):
        for _ in range(3):
            C.conv_forward(g, x, wf, y, stats=(stats[0], stats[1]) if with_stats else None)
        torch.cuda.synchronize()
        lib().mp_set_tunable(b'igemm_trace', trace.data_ptr())
        trace.zero_()
        C.conv_forward(g, x, wf, y, stats=(stats[0], stats[1]) if with_stats else None)
        torch.cuda.synchronize()
        lib().mp_set_tunable(b'igemm_trace', 0)
        t = trace.cpu().tolist()
        t0 = t[0]
        rel = lambda v: (v - t0) / 1e3 if v else float('nan')
        print('%d->%d %dx%d batch %d stats=%s: prologue done %.2f us, exit %.2f us' % (cin, cout, h, h, n, with_stats, rel(t[1]), rel(t[2])))
        for it in range(15):
            if t[4 + it * 4] == 0:
                break
            print('   tile %2d: operands landed %6.2f  MMAs issued %6.2f  accumulator done %6.2f  drained %6.2f' % (
                it, rel(t[4 + it * 4]), rel(t[5 + it * 4]), rel(t[6 + it * 4]), rel(t[7 + it * 4])))
