"""Time of the BatchNorm kernels against the number of pixels (C = 128): separates the fixed per-launch cost from the
streaming rate.  Each launch is replayed from a CUDA graph, back to back (PDL edges as in the training step), over
buffers that rotate through more than the L2.

    python tools/bn_scale.py [tunable=value ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes  # noqa: E402

import torch  # noqa: E402
from margipose_b200 import ops  # noqa: E402
from margipose_b200._lib import lib, stream_ptr  # noqa: E402

FROM_SUMS = False
for kv in sys.argv[1:]:
    if kv == 'from_sums':       # mp_bn_fwd derives its coefficients from the sums itself (no scale / shift from the conv)
        FROM_SUMS = True
        continue
    k, v = kv.split('=')
    assert lib().mp_set_tunable(k.encode(), int(v)) == 0, kv
dev = torch.device('cuda')
C = 128


def branch(M):
    y = torch.randn(M, C, device=dev).to(torch.bfloat16).reshape(1, M, 1, C)
    yf = y.float().reshape(M, C)
    ones = torch.ones(C, device=dev)
    return ops.BnBranchT(y, ones.clone(), torch.zeros(C, device=dev), running_mean=torch.zeros(C, device=dev),
                         running_var=ones.clone(), sum=yf.sum(0).contiguous(), sq=(yf * yf).sum(0).contiguous(),
                         save_mean=torch.zeros(C, device=dev), save_invstd=torch.zeros(C, device=dev),
                         dy=torch.zeros(1, M, 1, C, dtype=torch.bfloat16, device=dev),
                         dgamma=torch.zeros(C, device=dev), dbeta=torch.zeros(C, device=dev))


def timed(fns, reps=4):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))


print('%-10s %8s | %-28s | %-28s | %-28s' % ('pixels', 'MB/tensor', 'fwd 1 input (2 passes)', 'bwd reduce f1 (2 passes)',
                                               'bwd apply f1 (3 passes)'))
for M in (4096, 16384, 32768, 65536, 98304, 196608, 393216):
    mb = M * C * 2 / 1e6
    nset = max(2, int(400 / (3 * mb)) + 1)          # rotate through > 400 MB
    nset = min(nset, 24)
    fw, rd, ap = [], [], []
    for _ in range(nset):
        a = branch(M)
        out = torch.zeros(1, M, 1, C, dtype=torch.bfloat16, device=dev)
        args = ops.bn_args(a, None, relu_a=True, out=out, C=C, hw=M)
        coef = torch.ones(2, C, device=dev)
        if not FROM_SUMS:
            args.a.scale, args.a.shift = coef[0].data_ptr(), coef[1].data_ptr()
        fw.append(lambda args=args: ops.bn_fwd(args, dev))
        dout = torch.randn(M, C, device=dev).to(torch.bfloat16).reshape(1, M, 1, C)
        sums = torch.zeros(4, C, device=dev)
        bargs = ops.bn_args(a, None, relu_a=True, out=out, dout=dout, sums=sums, C=C, hw=M)
        if not FROM_SUMS:
            bargs.a.scale, bargs.a.shift = coef[0].data_ptr(), coef[1].data_ptr()
        bargs._keep = (a, out, dout, sums, args, coef)
        rd.append(lambda b=bargs: lib().mp_bn_bwd_reduce(ctypes.byref(b), stream_ptr(dev)))
        ap.append(lambda b=bargs: lib().mp_bn_bwd_apply(ctypes.byref(b), stream_ptr(dev)))
    tf, tr, ta = timed(fw), timed(rd), timed(ap)
    print('%-10d %8.1f | %6.1f us %6.0f GB/s        | %6.1f us %6.0f GB/s        | %6.1f us %6.0f GB/s' % (
        M, mb, tf, 2 * mb / tf * 1e3, tr, 2 * mb / tr * 1e3, ta, 3 * mb / ta * 1e3))
