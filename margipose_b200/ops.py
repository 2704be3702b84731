"""Tensor-level wrappers of the elementwise / BatchNorm entry points of the C ABI
(include/margipose_b200.h).  The engine records these launches with pre-built argument structs;
the wrappers here are the eager form (used by tests, tools and one-off calls).  CUDA only."""
import ctypes

import torch

from ._lib import (BnArgs, PackEntry, lib, check, stream_ptr, planes, require_cuda)


def _p(t):
    return t.data_ptr() if t is not None else None


class BnBranchT:
    """Tensors of one BatchNorm branch: conv output y (M, Cp) bf16 + its nn.BatchNorm2d state."""

    def __init__(self, y, gamma, beta, running_mean=None, running_var=None, sum=None, sq=None,
                 save_mean=None, save_invstd=None, conv_bias=None, dy=None, dgamma=None, dbeta=None):
        self.__dict__.update(locals())

    def fill(self, br):
        br.y, br.sum, br.sq = _p(self.y), _p(self.sum), _p(self.sq)
        br.gamma, br.beta = _p(self.gamma), _p(self.beta)
        br.running_mean, br.running_var = _p(self.running_mean), _p(self.running_var)
        br.save_mean, br.save_invstd = _p(self.save_mean), _p(self.save_invstd)
        br.conv_bias, br.dy, br.dgamma, br.dbeta = _p(self.conv_bias), _p(self.dy), _p(self.dgamma), _p(self.dbeta)


def bn_args(a, b=None, res=None, relu_a=False, relu_out=False, out=None, out_nchw=None, dout=None,
            dout_nchw=None, dres=None, sums=None, C=None, hw=0, training=True, momentum=0.1, eps=1e-5,
            stat_replicas=1, stat_stride=0):
    args = BnArgs()
    a.fill(args.a)
    if b is not None:
        b.fill(args.b)
    args.res, args.out, args.out_nchw = _p(res), _p(out), _p(out_nchw)
    args.dout, args.dout_nchw, args.dres, args.sums = _p(dout), _p(dout_nchw), _p(dres), _p(sums)
    args.relu_a, args.relu_out = int(relu_a), int(relu_out)
    args.Cp = a.y.shape[-1]
    args.M = a.y.numel() // args.Cp
    args.C = C if C is not None else a.gamma.numel()
    args.HW = hw
    args.training, args.momentum, args.eps = int(training), momentum, eps
    args.stat_replicas, args.stat_stride = stat_replicas, stat_stride
    return args


def bn_fwd(args, device):
    check(lib().mp_bn_fwd(ctypes.byref(args), stream_ptr(device)), 'mp_bn_fwd')


def bn_bwd(args, device):
    check(lib().mp_bn_bwd_reduce(ctypes.byref(args), stream_ptr(device)), 'mp_bn_bwd_reduce')
    check(lib().mp_bn_bwd_apply(ctypes.byref(args), stream_ptr(device)), 'mp_bn_bwd_apply')


def maxpool_fwd(x):
    require_cuda(x)
    n, h, w, c = x.shape
    y = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=x.device)
    idx = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=x.device)
    check(lib().mp_maxpool_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr(), n, h, w, c, 0, stream_ptr(x.device)),
          'mp_maxpool_fwd')
    return y, idx


def maxpool_bwd(dy, idx):
    n, ho, wo, c = dy.shape
    dx = torch.empty(n, 2 * ho, 2 * wo, c, dtype=torch.bfloat16, device=dy.device)
    check(lib().mp_maxpool_bwd(dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), n, 2 * ho, 2 * wo, c,
                               stream_ptr(dy.device)), 'mp_maxpool_bwd')
    return dx


def axis_permute(x, mode, channels):
    n, s, s2, cp = x.shape
    assert s == s2
    out = torch.empty_like(x)
    check(lib().mp_axis_permute(x.data_ptr(), out.data_ptr(), mode, n, s, channels, cp, stream_ptr(x.device)),
          'mp_axis_permute')
    return out


def combiner_fwd(probs, w, inp):
    n, j, h, wd = probs[0].shape
    out = torch.empty_like(inp)
    check(lib().mp_combiner_fwd(planes(probs), w.data_ptr(), inp.data_ptr(), out.data_ptr(), n, j, h * wd,
                                inp.shape[-1], 0, stream_ptr(inp.device)), 'mp_combiner_fwd')
    return out


def combiner_bwd(dout, probs, w, dprobs, dw, accumulate):
    n, j, h, wd = probs[0].shape
    check(lib().mp_combiner_bwd(dout.data_ptr(), planes(probs), w.data_ptr(), planes(dprobs), dw.data_ptr(),
                                int(accumulate), n, j, h * wd, dout.shape[-1], 0, stream_ptr(dout.device)),
          'mp_combiner_bwd')


def stem_im2col(x):
    require_cuda(x)
    n, c, h, w = x.shape
    assert c == 3 and x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty(n, h // 2, w // 2, 192, dtype=torch.bfloat16, device=x.device)
    check(lib().mp_stem_im2col(x.data_ptr(), out.data_ptr(), n, h, w, 0, stream_ptr(x.device)), 'mp_stem_im2col')
    return out


def add_bf16(tensors):
    out = torch.empty_like(tensors[0])
    arr = (ctypes.c_void_p * 4)(*([t.data_ptr() for t in tensors] + [None] * (4 - len(tensors))))
    check(lib().mp_add_bf16(ctypes.byref(arr), len(tensors), out.data_ptr(), out.numel(), 0,
                            stream_ptr(out.device)), 'mp_add_bf16')
    return out


def sgd_step(param, grad, buf, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False,
             first_step=False, grad_scale=1.0, hyper=None):
    """hyper: optional device tensor (lr, momentum, dampening, weight_decay, grad_scale) read by the kernel
    at run time instead of the launch-time scalars (CUDA-graph-safe schedules)."""
    check(lib().mp_sgd_step_hp(param.data_ptr(), grad.data_ptr(), _p(buf), param.numel(), lr, momentum, dampening,
                               weight_decay, int(nesterov), int(first_step), grad_scale, _p(hyper),
                               stream_ptr(param.device)), 'mp_sgd_step')
