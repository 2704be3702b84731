"""Shared checks: parity of the tcgen05 implicit-GEMM convolution kernels (through the C ABI) against
torch's CPU convolution -- the backend the reference's nn.Conv2d / nn.ConvTranspose2d call.

Operands are rounded to bf16 on both sides (DESIGN.md "Precision"), accumulation is fp32, the
kernel rounds its output to bf16 once: tolerance rtol 1e-2 / atol 2e-2 on O(1) outputs
(bf16 has 8 mantissa bits -> 2^-9 relative rounding, plus accumulation-order differences).
Weight gradients stay fp32: rtol 2e-3 / atol 2e-3 * sqrt(pixels)."""
import pytest
import torch
import torch.nn.functional as F

DEV = 'cuda'


def _sync():
    if DEV == 'cuda':
        torch.cuda.synchronize()


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def to_nhwc(x, cp):
    """fp32 NCHW (cpu) -> bf16 NHWC padded (cuda)."""
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, cp, dtype=torch.bfloat16, device=DEV)
    out[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16).to(DEV)
    return out


def from_nhwc(t, c):
    return t[..., :c].float().cpu().permute(0, 3, 1, 2).contiguous()


def ref_conv(g, x, w):
    if g.transposed:
        return F.conv_transpose2d(x, w, None, 2, g.k // 2, 1)
    return F.conv2d(x, w, None, g.stride, g.k // 2)


CASES = [  # cin, cout, k, stride, transposed, n, h, w
    (128, 128, 3, 1, False, 2, 32, 32),
    (192, 192, 3, 1, False, 3, 16, 16),
    (64, 64, 3, 1, False, 1, 64, 64),
    (128, 128, 1, 1, False, 2, 32, 32),
    (128, 17, 3, 1, False, 2, 32, 32),
    (17, 17, 3, 1, False, 2, 32, 32),
    (128, 17, 1, 1, False, 1, 32, 32),
    (128, 192, 3, 2, False, 2, 32, 32),
    (128, 192, 1, 2, False, 2, 32, 32),
    (64, 128, 3, 2, False, 1, 64, 64),
    (192, 128, 3, 2, True, 2, 16, 16),
    (192, 128, 1, 2, True, 2, 16, 16),
    (192, 192, 3, 1, False, 2, 24, 24),
    (128, 128, 3, 1, False, 1, 48, 48),
    (128, 192, 3, 2, False, 1, 48, 48),
    (192, 128, 3, 2, True, 1, 24, 24),
    (256, 512, 1, 1, False, 1, 16, 16),
    (512, 128, 1, 1, False, 1, 16, 16),
    (64, 64, 3, 1, False, 1, 96, 96),
]


def _mk(case, seed=0):
    from margipose_b200 import convops as C
    cin, cout, k, stride, tr, n, h, w = case
    g = C.ConvGeom(cin, cout, k, stride, tr)
    gen = torch.Generator().manual_seed(seed)
    x = _bf(torch.randn(n, cin, h, w, generator=gen))
    wt = _bf(torch.randn(*g.torch_weight_shape, generator=gen) / (cin * k * k) ** 0.5)
    return C, g, x, wt


def check_conv_forward_and_stats(case):
    C, g, x, wt = _mk(case)
    n, _, h, w = x.shape
    ho, wo = g.out_hw(h, w)
    master = C.master_from_torch(g, wt).to(DEV)
    xg = to_nhwc(x, g.cin_p)
    out = torch.zeros(n, ho, wo, g.cout_p, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(2, g.cout_p, device=DEV)
    C.conv_forward(g, xg, C.pack_fwd(g, master), out, stats=(stats[0], stats[1]))
    _sync()
    want = ref_conv(g, x, wt)
    got = from_nhwc(out, g.cout)
    torch.testing.assert_close(got, want, rtol=1e-2, atol=2e-2)
    assert out[..., g.cout:].abs().max().item() == 0 if g.cout_p > g.cout else True
    of = out.float().reshape(-1, g.cout_p)
    torch.testing.assert_close(stats[0], of.sum(0), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(stats[1], (of * of).sum(0), rtol=1e-4, atol=1e-2)


def check_conv_dgrad_and_wgrad(case):
    C, g, x, wt = _mk(case, seed=1)
    n, _, h, w = x.shape
    ho, wo = g.out_hw(h, w)
    gen = torch.Generator().manual_seed(5)
    dy = _bf(torch.randn(n, g.cout, ho, wo, generator=gen))
    res = _bf(torch.randn(n, g.cin, h, w, generator=gen))
    xr, wr = x.clone().requires_grad_(), wt.clone().requires_grad_()
    ref_conv(g, xr, wr).backward(dy)
    master = C.master_from_torch(g, wt).to(DEV)
    dyg, xg = to_nhwc(dy, g.cout_p), to_nhwc(x, g.cin_p)
    dx = torch.zeros(n, h, w, g.cin_p, dtype=torch.bfloat16, device=DEV)
    C.conv_dgrad(g, dyg, C.pack_bwd(g, master), dx)
    _sync()
    if not (g.k == 1 and g.stride == 2 and not g.transposed):   # 1x1 s2: odd pixels are never written
        torch.testing.assert_close(from_nhwc(dx, g.cin), xr.grad, rtol=1e-2, atol=2e-2)
    else:
        torch.testing.assert_close(from_nhwc(dx, g.cin)[..., ::2, ::2], xr.grad[..., ::2, ::2],
                                   rtol=1e-2, atol=2e-2)
    if not (g.k == 1 and g.stride == 2):
        dx2 = torch.zeros_like(dx)
        C.conv_dgrad(g, dyg, C.pack_bwd(g, master), dx2, res=to_nhwc(res, g.cin_p))
        torch.testing.assert_close(from_nhwc(dx2, g.cin), xr.grad + res, rtol=1e-2, atol=3e-2)
    dw = torch.zeros(g.master_shape, device=DEV)
    C.conv_wgrad(g, xg, dyg, dw)
    C.conv_wgrad(g, xg, dyg, dw)   # accumulates
    _sync()
    want = 2 * C.master_from_torch(g, wr.grad)
    scale = (n * ho * wo) ** 0.5
    torch.testing.assert_close(dw.cpu(), want, rtol=2e-3, atol=2e-3 * scale)


def check_fused_block_dgrad(stride, tr):
    """dx of a residual block's 3x3 main conv and 1x1 shortcut conv in ONE accumulation."""
    from margipose_b200 import convops as C
    cin, cout = (192, 128) if tr else (128, 192 if stride == 2 else 128)
    n, h, w = 2, (16 if tr else 32), (16 if tr else 32)
    g1, g2 = C.ConvGeom(cin, cout, 3, stride, tr), C.ConvGeom(cin, cout, 1, stride, tr)
    gen = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(n, cin, h, w, generator=gen)).requires_grad_()
    w1 = _bf(torch.randn(*g1.torch_weight_shape, generator=gen) / (cin * 9) ** 0.5)
    w2 = _bf(torch.randn(*g2.torch_weight_shape, generator=gen) / cin ** 0.5)
    ho, wo = g1.out_hw(h, w)
    dy1 = _bf(torch.randn(n, cout, ho, wo, generator=gen))
    dy2 = _bf(torch.randn(n, cout, ho, wo, generator=gen))
    (ref_conv(g1, x, w1) * dy1).sum().backward()
    (ref_conv(g2, x, w2) * dy2).sum().backward()
    p1 = C.pack_bwd(g1, C.master_from_torch(g1, w1).to(DEV))
    p2 = C.pack_bwd(g2, C.master_from_torch(g2, w2).to(DEV))
    both = torch.cat([p1, p2], 1).contiguous()
    dx = torch.zeros(n, h, w, g1.cin_p, dtype=torch.bfloat16, device=DEV)
    C.conv_dgrad(g1, to_nhwc(dy1, g1.cout_p), both, dx,
                 second=(g2, to_nhwc(dy2, g2.cout_p), p1.shape[1]))
    _sync()
    torch.testing.assert_close(from_nhwc(dx, cin), x.grad, rtol=1e-2, atol=3e-2)


# ---------------------------------------------------------------------------- split ("bf16x3") precision mode
def _pair(x, cp):
    """fp32 NCHW (cpu) -> registered (hi, lo) bf16 NHWC padded pair, halves a fixed distance apart."""
    from margipose_b200 import convops as C
    n, c, h, w = x.shape
    buf = torch.zeros(2, n, h, w, cp, dtype=torch.bfloat16, device=DEV)
    v = x.permute(0, 2, 3, 1).contiguous()
    hi, lo = C.split_bf16(v)
    buf[0, ..., :c] = hi.to(DEV)
    buf[1, ..., :c] = lo.to(DEV)
    return buf


def _pack_pair(packed_fp32):
    """Split-mode weight operand: [rows][K_hi | K_lo] in ONE matrix (what mp_pack_weights writes with lo_off = K)."""
    from margipose_b200 import convops as C
    hi, lo = C.split_bf16(packed_fp32)
    return torch.cat([hi, lo], 1).contiguous()


def _pack_fp32(g, master, bwd):
    """Same layout as convops.pack_fwd / pack_bwd, kept in fp32 (the split mode packs W_hi and W_lo from it)."""
    if not bwd:
        m = master.permute(2, 1, 0) if g.transposed else master
        out = torch.zeros(g.cout_p, g.taps, g.cin_p, device=master.device)
        out[:g.cout, :, :g.cin] = m
        return out.reshape(g.cout_p, g.taps * g.cin_p)
    m = master if g.transposed else master.permute(2, 1, 0)
    out = torch.zeros(g.cin_p, g.taps, g.cout_p, device=master.device)
    out[:g.cin, :, :g.cout] = m
    return out.reshape(g.cin_p, g.taps * g.cout_p)


def check_split_conv(case):
    """Forward (+ BatchNorm sums), data gradient (+ residual) and weight gradient with every operand a bf16 pair
    and three tensor-core passes per accumulation, against the fp64 convolution of the SAME fp32 inputs:
    agreement ~1e-5 relative (pair rounding 2^-17 per operand), i.e. fp32-grade results from bf16 MMAs."""
    from margipose_b200 import convops as C
    cin, cout, k, stride, tr, n, h, w = case
    g = C.ConvGeom(cin, cout, k, stride, tr)
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(n, cin, h, w, generator=gen)
    wt = torch.randn(*g.torch_weight_shape, generator=gen) / (cin * k * k) ** 0.5
    ho, wo = g.out_hw(h, w)
    dy = torch.randn(n, cout, ho, wo, generator=gen)
    res = torch.randn(n, cin, h, w, generator=gen)
    xr, wr = x.double().requires_grad_(), wt.double().requires_grad_()
    y_ref = ref_conv(g, xr, wr)
    y_ref.backward(dy.double())
    master = C.master_from_torch(g, wt).to(DEV)
    pairs = C.SplitPairs()

    def reg(buf):
        return pairs.register(buf[0], buf[1])
    xg, dyg, resg = reg(_pair(x, g.cin_p)), reg(_pair(dy, g.cout_p)), reg(_pair(res, g.cin_p))
    wf, wb = _pack_pair(_pack_fp32(g, master, False)), _pack_pair(_pack_fp32(g, master, True))
    out = reg(torch.zeros(2, n, ho, wo, g.cout_p, dtype=torch.bfloat16, device=DEV))
    dx = reg(torch.zeros(2, n, h, w, g.cin_p, dtype=torch.bfloat16, device=DEV))
    dx2 = reg(torch.zeros(2, n, h, w, g.cin_p, dtype=torch.bfloat16, device=DEV))
    stats = torch.zeros(2, g.cout_p, device=DEV)
    dw = torch.zeros(g.master_shape, device=DEV)
    old, C.SPLIT = C.SPLIT, pairs
    try:
        C.conv_forward(g, xg, wf, out, stats=(stats[0], stats[1]))
        C.conv_dgrad(g, dyg, wb, dx)
        if not (g.k == 1 and g.stride == 2):
            C.conv_dgrad(g, dyg, wb, dx2, res=resg)
        C.conv_wgrad(g, xg, dyg, dw)
    finally:
        C.SPLIT = old
    _sync()

    def val(t, c):
        return from_nhwc(pairs.value(t), c)
    got = val(out, g.cout)
    err_y = ((got - y_ref.detach().float()).abs().max() / y_ref.abs().max()).item()
    torch.testing.assert_close(got, y_ref.detach().float(), rtol=2e-5, atol=2e-5 * y_ref.abs().max().item())
    of = pairs.value(out).reshape(-1, g.cout_p)
    torch.testing.assert_close(stats[0], of.sum(0), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(stats[1], (of * of).sum(0), rtol=1e-4, atol=1e-2)
    gx = xr.grad.float()
    scale = gx.abs().max().item()
    if not (g.k == 1 and g.stride == 2 and not g.transposed):
        torch.testing.assert_close(val(dx, g.cin), gx, rtol=2e-5, atol=2e-5 * scale)
    else:
        torch.testing.assert_close(val(dx, g.cin)[..., ::2, ::2], gx[..., ::2, ::2], rtol=2e-5, atol=2e-5 * scale)
    if not (g.k == 1 and g.stride == 2):
        torch.testing.assert_close(val(dx2, g.cin), gx + res, rtol=2e-5, atol=2e-5 * (scale + 4))
    want = C.master_from_torch(g, wr.grad.float())
    torch.testing.assert_close(dw.cpu(), want, rtol=5e-5, atol=5e-5 * want.abs().max().item())
    return err_y
