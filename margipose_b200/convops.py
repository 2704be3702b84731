"""Host-side geometry of the convolution layers on the MargiPose hot path.

Turns one nn.Conv2d / nn.ConvTranspose2d op site of the reference
(/root/reference/src/margipose/models/margipose_model.py:33,67-68,73-74,79-82 and the torchvision
ResNet blocks behind :130-135) into launches of the tcgen05 implicit-GEMM kernels behind the C
ABI (include/margipose_b200.h: mp_conv_igemm, mp_conv_wgrad): which 5-D view of the NHWC
activation each filter tap reads, with which shift / parity, and which K-slice of the packed
weight matrix it multiplies.

Conventions
  activations : bf16 (N, H, W, Cp) contiguous, Cp = channels padded to a multiple of 64
  weights     : fp32 master in channels-last memory -- Conv2d (Co, Ci, kh, kw) stored as
                [Co][kh*kw][Ci], ConvTranspose2d (Ci, Co, kh, kw) stored as [Ci][kh*kw][Co] --
                and two bf16 packs per layer (forward and data-gradient orientation), each
                [rows padded to 64][taps][K padded to 64]
Supported: kernel 1 or 3 (padding k // 2), stride 1 or 2; transposed only with stride 2 and
output_padding 1 -- everything the reference's columns and ResNet blocks use.

Split ("bf16x3") precision mode: while a `SplitPairs` object is installed in `SPLIT`, every activation and
every weight pack is a PAIR of bf16 tensors (hi = bf16(v), lo = bf16(v - hi), ~16 significant bits) and each
convolution / weight gradient becomes three tensor-core passes -- hi*hi, lo*hi, hi*lo -- over one accumulator
(DESIGN.md "Precision modes").  The geometry code below is unchanged by it.
"""
import ctypes

import torch

from ._lib import (IgemmArgs, WgradArgs, View5, MP_MAX_TAPS, lib, check, stream_ptr)


def pad64(c):
    return (c + 63) // 64 * 64


class SplitPairs:
    """hi tensor -> lo tensor registry of the split precision mode (keyed by device pointer)."""

    def __init__(self):
        self._lo = {}

    def register(self, hi, lo):
        assert hi.shape == lo.shape and hi.dtype == lo.dtype == torch.bfloat16
        self._lo[hi.data_ptr()] = lo
        return hi

    def lo(self, hi):
        return self._lo[hi.data_ptr()]

    def delta(self, hi):
        """Element offset from hi to lo (the C ABI's lo_delta)."""
        lo = self._lo[hi.data_ptr()]
        d = lo.data_ptr() - hi.data_ptr()
        assert d > 0 and d % 16 == 0
        return d // 2

    def value(self, hi):
        return hi.float() + self._lo[hi.data_ptr()].float()


def split_bf16(v):
    """fp32 tensor -> (hi, lo) bf16 pair."""
    hi = v.to(torch.bfloat16)
    return hi, (v - hi.float()).to(torch.bfloat16)


SPLIT = None    # a SplitPairs while launches are to be issued in split precision mode


def _view(t, parity=False):
    """5-D TMA view (C, W, P, H, N) of an (N, H, W, Cp) bf16 tensor; parity=True exposes the
    even/odd rows and columns as separate coordinates (stride-2 taps)."""
    n, h, w, c = t.shape
    assert t.dtype == torch.bfloat16 and t.is_contiguous() and c % 64 == 0
    v = View5()
    v.ptr = t.data_ptr()
    if not parity:
        dims, strides = (c, w, 1, h, n), (1, c, w * c, w * c, h * w * c)
    else:
        assert h % 2 == 0 and w % 2 == 0
        dims, strides = (2 * c, w // 2, 2, h // 2, n), (1, 2 * c, w * c, 2 * w * c, h * w * c)
    for i in range(5):
        v.dim[i] = dims[i]
        v.stride[i] = strides[i]
    return v


def _axis_taps_s1(k):
    """(kernel index, shift) pairs along one axis of a stride-1 conv: in = out + r - k//2."""
    return [(r, r - k // 2) for r in range(k)]


def _axis_taps_s2(k):
    """(kernel index, parity, shift in the decimated axis) for in = 2*out + r - k//2."""
    out = []
    for r in range(k):
        off = r - k // 2
        out.append((r, off % 2, (off - off % 2) // 2))
    return out


def _axis_taps_up(k, parity):
    """Taps that reach output parity `parity` of a stride-2 transposed relation
    out = 2*in + r - k//2 : (kernel index, shift of `in` relative to out // 2)."""
    out = []
    for r in range(k):
        off = r - k // 2            # out = 2*in + off  ->  in = (out - off) / 2
        if (parity - off) % 2 == 0:
            out.append((r, (parity - off) // 2))
    return out


class ConvGeom:
    """Geometry of one conv layer (shared by forward, dgrad and wgrad launches)."""

    def __init__(self, cin, cout, k, stride=1, transposed=False):
        assert k in (1, 3) and stride in (1, 2)
        assert not transposed or stride == 2
        self.cin, self.cout, self.k, self.stride, self.transposed = cin, cout, k, stride, transposed
        self.cin_p, self.cout_p = pad64(cin), pad64(cout)
        self.taps = k * k

    # master weight memory: [rows][taps][cols] fp32 (channels-last of the torch parameter)
    @property
    def master_shape(self):
        return (self.cin, self.taps, self.cout) if self.transposed else (self.cout, self.taps, self.cin)

    @property
    def torch_weight_shape(self):
        return (self.cin, self.cout, self.k, self.k) if self.transposed else \
            (self.cout, self.cin, self.k, self.k)

    def out_hw(self, h, w):
        if self.transposed:
            return 2 * h, 2 * w
        return (h // 2, w // 2) if self.stride == 2 else (h, w)

    # ---- packs: fwd multiplies x (K = cin) giving cout rows; bwd multiplies dy (K = cout)
    def fwd_pack_shape(self):
        return (self.cout_p, self.taps * self.cin_p)

    def bwd_pack_shape(self):
        return (self.cin_p, self.taps * self.cout_p)


def _fill_taps(args, taps):
    assert 1 <= len(taps) <= MP_MAX_TAPS
    args.n_taps = len(taps)
    for i, (src, c0, dw, p, dh, koff) in enumerate(taps):
        t = args.taps[i]
        t.src, t.c0, t.dw, t.p, t.dh, t.koff = src, c0, dw, p, dh, koff


def _igemm(srcs, wmat, taps, cblocks, n_img, out_h, out_w, out, out_strides, out_c, res=None,
           stats=None, out_offset=0, bn=None, defer=None, ep=None):
    """One convolution accumulation.  Split precision mode: x_hi * W_hi + x_lo * W_hi + x_hi * W_lo, where `wmat` is
    [rows][K_hi | K_lo] (the W_lo taps read K columns further).  With one source and a tripled tap list that fits
    MP_MAX_TAPS it is ONE launch over the sources (x_hi, x_lo); otherwise (the fused data gradient of a residual
    block has two sources already) three passes chained through `acc_in`, with residual, statistics, BatchNorm
    finalize and epilogue affine on the last one."""
    S = SPLIT
    if S is None:
        return _igemm_one(srcs, wmat, taps, cblocks, n_img, out_h, out_w, out, out_strides, out_c, res=res,
                          stats=stats, out_offset=out_offset, bn=bn, defer=defer, ep=ep)
    assert wmat.shape[1] % 2 == 0
    k_half = wmat.shape[1] // 2
    common = (cblocks, n_img, out_h, out_w, out, out_strides, out_c)
    out_lo = S.lo(out)
    res_lo = S.lo(res) if res is not None else None
    lo_taps = [(src, c0, dw, p, dh, koff + k_half) for (src, c0, dw, p, dh, koff) in taps]
    if len(srcs) == 1 and 3 * len(taps) <= MP_MAX_TAPS:
        (x, par), = srcs
        fused = list(taps) + [(1, c0, dw, p, dh, koff) for (_s, c0, dw, p, dh, koff) in taps] + lo_taps
        flop_channels = _FLOP_CHANNELS[0]
        _FLOP_CHANNELS[0] = flop_channels / 3.0        # algorithmic FLOPs: the convolution once
        _igemm_one([(x, par), (S.lo(x), par)], wmat, fused, *common, res=res, res_lo=res_lo, stats=stats,
                   out_offset=out_offset, bn=bn, defer=defer, ep=ep, out_lo=out_lo)
        _FLOP_CHANNELS[0] = flop_channels
        return
    # a residual that aliases the output (in-place accumulation) must be read before the first pass overwrites it
    first = dict(res=res, res_lo=res_lo) if (res is not None and res.data_ptr() == out.data_ptr()) else {}
    assert not (first and ep is not None), 'an in-place residual cannot be combined with an epilogue affine'
    last = {} if first else dict(res=res, res_lo=res_lo)
    flop_channels = _FLOP_CHANNELS[0]
    _FLOP_CHANNELS[0] = 0          # algorithmic FLOPs are counted once, on the last pass
    _igemm_one(srcs, wmat, taps, *common, out_offset=out_offset, defer=defer, out_lo=out_lo, **first)
    _igemm_one([(S.lo(t), par) for t, par in srcs], wmat, taps, *common, out_offset=out_offset, defer=defer,
               out_lo=out_lo, acc_in=out)
    _FLOP_CHANNELS[0] = flop_channels
    _igemm_one(srcs, wmat, lo_taps, *common, stats=stats, out_offset=out_offset, bn=bn, defer=defer, ep=ep,
               out_lo=out_lo, acc_in=out, **last)


def _igemm_one(srcs, wmat, taps, cblocks, n_img, out_h, out_w, out, out_strides, out_c, res=None,
               stats=None, out_offset=0, bn=None, defer=None, ep=None, out_lo=None, acc_in=None, res_lo=None):
    """srcs: list of (tensor, parity) pairs; see mp_conv_igemm in include/margipose_b200.h.
    bn = dict(branch=BnBranch, counter=ptr, channels=C, count=M, momentum=, eps=): fuse the BatchNorm
    finalize into the launch(es); defer = list collecting the argument structs of a multi-launch
    conv so the caller can set the shared arrival total before launching.
    ep = (scale ptr, shift ptr, relu mode): per-channel affine + ReLU in the epilogue (eval-mode BatchNorm).
    out_lo (split mode): low half of the output pair; acc_in: output pair (hi tensor) of the earlier passes;
    res_lo: low half of the residual pair."""
    a = IgemmArgs()
    for i, (t, parity) in enumerate(srcs):
        a.src[i] = _view(t, parity)
    a.wmat = wmat.data_ptr()
    a.w_rows, a.w_k = wmat.shape
    _fill_taps(a, taps)
    a.cblocks = cblocks
    a.n_img, a.out_h, a.out_w = n_img, out_h, out_w
    a.out = out.data_ptr() + 2 * out_offset
    a.res = (res.data_ptr() + 2 * out_offset) if res is not None else None
    a.out_sn, a.out_sh, a.out_sw = out_strides
    a.out_c = out_c
    if stats is not None:
        a.stat_sum, a.stat_sq = stats[0].data_ptr(), stats[1].data_ptr()
        a.stat_replicas, a.stat_stride = (stats[2], stats[3]) if len(stats) > 2 else (1, 0)
    a._flops = 2.0 * n_img * out_h * out_w * len(taps) * _FLOP_CHANNELS[0]
    if bn is not None:
        a._keep = bn['branch']
        a.bn = ctypes.addressof(bn['branch'])
        a.bn_counter, a.bn_channels, a.bn_count = bn['counter'], bn['channels'], bn['count']
        a.bn_momentum, a.bn_eps = bn['momentum'], bn['eps']
        a.bn_launches = 1
    if ep is not None:
        a.ep_scale, a.ep_shift, a.ep_relu = ep
    if out_lo is not None:
        a.lo_delta = (out_lo.data_ptr() - out.data_ptr()) // 2
        assert a.lo_delta > 0 and (res is None or res_lo.data_ptr() - res.data_ptr() == 2 * a.lo_delta), \
            'the halves of every activation pair must be the same distance apart'
        if acc_in is not None:
            a.acc_in = acc_in.data_ptr() + 2 * out_offset
    if defer is not None:
        defer.append((a, out.device))
    else:
        _igemm_launch(a, out.device)


# real (unpadded) cin*cout of the layer being launched; set by conv_forward / conv_dgrad so the
# per-launch algorithmic FLOP count ignores channel padding
_FLOP_CHANNELS = [0]


def _igemm_launch(a, device):
    """Enqueues one mp_conv_igemm; the engine swaps this hook to record launches instead."""
    check(lib().mp_conv_igemm(ctypes.byref(a), stream_ptr(device)), 'mp_conv_igemm')


def _wgrad_launch(a, device):
    check(lib().mp_conv_wgrad(ctypes.byref(a), stream_ptr(device)), 'mp_conv_wgrad')


def _wgrad(a_t, b_t, b_parity, taps, m_real, n_real, n_cols, n_slots, n_img, grid_h, grid_w, dw):
    """One weight gradient.  Split precision mode: a_hi*b_hi + a_lo*b_hi + a_hi*b_lo per pixel chunk, accumulated
    by ONE launch into the same accumulator (mp_wgrad_args.a_lo / b_lo)."""
    S = SPLIT
    rest = (b_parity, taps, m_real, n_real, n_cols, n_slots, n_img, grid_h, grid_w, dw)
    if S is None:
        return _wgrad_one(a_t, b_t, *rest)
    _wgrad_one(a_t, b_t, *rest, a_lo=S.lo(a_t), b_lo=S.lo(b_t))


def _wgrad_one(a_t, b_t, b_parity, taps, m_real, n_real, n_cols, n_slots, n_img, grid_h, grid_w, dw,
               a_lo=None, b_lo=None):
    """See mp_conv_wgrad in include/margipose_b200.h; wide b operands go in 256-channel slices."""
    for n_off in range(0, n_cols, 256):
        a = WgradArgs()
        a.a = _view(a_t)
        a.b = _view(b_t, b_parity)
        if a_lo is not None:
            a.a_lo = _view(a_lo)
            a.b_lo = _view(b_lo, b_parity)
        _fill_taps(a, taps)
        a.m_real, a.n_real, a.n_slots = m_real, n_real, n_slots
        a.n_cols, a.n_off = min(256, n_cols - n_off), n_off
        a.n_img, a.grid_h, a.grid_w = n_img, grid_h, grid_w
        a.dw = dw.data_ptr()
        a._flops = 2.0 * n_img * grid_h * grid_w * len(taps) * m_real * min(a.n_cols, n_real - n_off)
        if n_off < n_real:
            _wgrad_launch(a, dw.device)


def _down_taps(k, c_in_p, k_stride, src=0, koff0=0):
    """Taps of a stride-2 gather (conv forward s2 / transposed-conv dgrad) over a parity view."""
    taps = []
    for r, ph, dh in _axis_taps_s2(k):
        for s, pw, dw in _axis_taps_s2(k):
            taps.append((src, pw * c_in_p, dw, ph, dh, koff0 + (r * k + s) * k_stride))
    return taps


def _s1_taps(k, k_stride, sign, src=0, koff0=0):
    taps = []
    for r, dh in _axis_taps_s1(k):
        for s, dw in _axis_taps_s1(k):
            taps.append((src, 0, sign * dw, 0, sign * dh, koff0 + (r * k + s) * k_stride))
    return taps


def conv_forward(g, x, wpack, out, stats=None, res=None, bn=None, ep=None):
    """y = conv(x) (or conv_transpose(x)); x (N,H,W,cin_p), out (N,Ho,Wo,cout_p), both bf16 NHWC.
    stats = (sum, sumsq[, replicas, stride]) fp32 (cout_p) accumulators for the BatchNorm batch
    statistics (optionally `replicas` copies `stride` floats apart to spread the atomics)."""
    n, h, w, _ = x.shape
    ho, wo = g.out_hw(h, w)
    assert tuple(out.shape) == (n, ho, wo, g.cout_p)
    assert tuple(wpack.shape) == (g.cout_p, g.taps * g.cin_p * (2 if SPLIT is not None else 1))
    cb = g.cin_p // 64
    _FLOP_CHANNELS[0] = g.cin * g.cout
    if not g.transposed:
        if g.stride == 1:
            taps, src = _s1_taps(g.k, g.cin_p, +1), (x, False)
        else:
            taps, src = _down_taps(g.k, g.cin_p, g.cin_p), (x, True)
        _igemm([src], wpack, taps, cb, n, ho, wo, out, (ho * wo * g.cout_p, wo * g.cout_p, g.cout_p),
               g.cout_p, res=res, stats=stats, bn=bn, ep=ep)
    else:
        if ep is not None and g.k == 1:
            raise ValueError('an epilogue affine cannot be fused into a 1x1 transposed conv: three of its four '
                             'output parity classes are not written by any launch')
        _scatter_up(g.k, (x, False), wpack, g.cin_p, cb, n, h, w, out, g.cout_p, stats, res, bn=bn, ep=ep)


def _scatter_up(k, src, wpack, k_stride, cb, n, h, w, out, c_out_p, stats, res, extra=None, bn=None, ep=None):
    """The stride-2 'up' relation out[2a+ph, 2b+pw] = sum of taps over in[a+dh, b+dw]: one launch
    per output parity class, scattered into the (N, 2h, 2w, C) output.  Classes no tap reaches
    (1x1 kernels) stay zero in `out` (the buffer is zero-initialised once by its owner).
    extra = (k2, src2, k_stride2, koff0): a second source fused into the same accumulation."""
    wo = 2 * w
    strides = (4 * h * w * c_out_p, 2 * wo * c_out_p, 2 * c_out_p)
    pending = [] if bn is not None else None
    for ph in (0, 1):
        for pw in (0, 1):
            taps = [(0, 0, dw, 0, dh, (r * k + s) * k_stride)
                    for r, dh in _axis_taps_up(k, ph) for s, dw in _axis_taps_up(k, pw)]
            srcs = [src]
            if extra is not None:
                k2, src2, ks2, koff0 = extra
                srcs.append(src2)
                taps += [(1, 0, dw, 0, dh, koff0 + (r * k2 + s) * ks2)
                         for r, dh in _axis_taps_up(k2, ph) for s, dw in _axis_taps_up(k2, pw)]
            if not taps:
                continue
            _igemm(srcs, wpack, taps, cb, n, h, w, out, strides, c_out_p, res=res, stats=stats,
                   out_offset=(ph * wo + pw) * c_out_p, bn=bn, defer=pending, ep=ep)
    if pending:   # the BatchNorm statistics are complete when the CTAs of ALL parity classes have arrived
        n_bn = sum(1 for a, _dev in pending if a.bn)
        for a, dev in pending:
            if a.bn:
                a.bn_launches = n_bn
            _igemm_launch(a, dev)


def conv_dgrad(g, dy, wpack_bwd, dx, res=None, second=None):
    """dx = d/dx of conv_forward (optionally + res).  `second` = (g2, dy2, koff0) fuses the data
    gradient of another conv that read the same x with the same stride (the residual block's
    shortcut); wpack_bwd then holds both layers' bwd packs side by side along K."""
    n, ho, wo, _ = dy.shape
    cb = g.cout_p // 64
    _FLOP_CHANNELS[0] = g.cin * g.cout
    if second is not None:
        g2, dy2, koff0 = second
        assert g2.cout_p == g.cout_p and g2.cin_p == g.cin_p and g2.stride == g.stride and \
            g2.transposed == g.transposed
    if not g.transposed:
        if g.stride == 1:
            taps, srcs = _s1_taps(g.k, g.cout_p, -1), [(dy, False)]
            if second is not None:
                taps += _s1_taps(g2.k, g2.cout_p, -1, src=1, koff0=koff0)
                srcs.append((dy2, False))
            _igemm(srcs, wpack_bwd, taps, cb, n, ho, wo, dx, (ho * wo * g.cin_p, wo * g.cin_p, g.cin_p),
                   g.cin_p, res=res)
        else:
            extra = None
            if second is not None:
                extra = (g2.k, (dy2, False), g2.cout_p, koff0)
            _scatter_up(g.k, (dy, False), wpack_bwd, g.cout_p, cb, n, ho, wo, dx, g.cin_p, None, res,
                        extra=extra)
    else:
        h, w = ho // 2, wo // 2
        taps, srcs = _down_taps(g.k, g.cout_p, g.cout_p), [(dy, True)]
        if second is not None:
            taps += _down_taps(g2.k, g2.cout_p, g2.cout_p, src=1, koff0=koff0)
            srcs.append((dy2, True))
        _igemm(srcs, wpack_bwd, taps, cb, n, h, w, dx, (h * w * g.cin_p, w * g.cin_p, g.cin_p),
               g.cin_p, res=res)


def conv_wgrad(g, x, dy, dw):
    """dw += d/dW of conv_forward; dw is the fp32 master-layout gradient [rows][taps][cols]."""
    assert dw.dtype == torch.float32 and dw.is_contiguous() and tuple(dw.shape) == g.master_shape
    if not g.transposed:
        n, ho, wo, _ = dy.shape
        if g.stride == 1:
            taps, parity = _s1_taps(g.k, 1, +1), False
        else:
            taps, parity = _down_taps(g.k, g.cin_p, 1), True
        _wgrad(dy, x, parity, taps, g.cout, g.cin, g.cin_p, g.taps, n, ho, wo, dw)
    else:
        n, h, w, _ = x.shape
        _wgrad(x, dy, True, _down_taps(g.k, g.cout_p, 1), g.cin, g.cout, g.cout_p, g.taps, n, h, w, dw)


# ---- reference packing in torch (used by tests and by model materialisation until the CUDA pack
# kernel takes over); master is the fp32 [rows][taps][cols] tensor.
def pack_fwd(g, master):
    """bf16 [cout_p][taps*cin_p]: rows = output channels, K = (tap, input channel)."""
    m = master.permute(2, 1, 0) if g.transposed else master          # -> [cout][taps][cin]
    out = torch.zeros(g.cout_p, g.taps, g.cin_p, dtype=torch.bfloat16, device=master.device)
    out[:g.cout, :, :g.cin] = m.to(torch.bfloat16)
    return out.reshape(g.cout_p, g.taps * g.cin_p)


def pack_bwd(g, master):
    """bf16 [cin_p][taps*cout_p]: rows = input channels, K = (tap, output channel)."""
    m = master if g.transposed else master.permute(2, 1, 0)          # -> [cin][taps][cout]
    out = torch.zeros(g.cin_p, g.taps, g.cout_p, dtype=torch.bfloat16, device=master.device)
    out[:g.cin, :, :g.cout] = m.to(torch.bfloat16)
    return out.reshape(g.cin_p, g.taps * g.cout_p)


def master_from_torch(g, weight):
    """torch parameter (Co,Ci,kh,kw) / (Ci,Co,kh,kw) -> master [rows][taps][cols] fp32."""
    a, b, kh, kw = weight.shape
    return weight.permute(0, 2, 3, 1).reshape(a, kh * kw, b).contiguous().float()
