"""GPU parity of the BatchNorm / pooling / permutation / combiner / im2col / SGD kernels (through
the C ABI) against plain fp32 PyTorch on CPU -- the ops the reference calls at those sites.

bf16 tensors are exact inputs on both sides; outputs rounded to bf16 are compared with
rtol 1e-2 / atol 1e-2 (one bf16 rounding), fp32 outputs with rtol 1e-4 / atol 1e-5."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = dict(rtol=1e-2, atol=1e-2)
F32 = dict(rtol=1e-4, atol=1e-5)


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _nhwc(x, cp):   # fp32 NCHW cpu -> bf16 NHWC padded cuda
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, cp, dtype=torch.bfloat16, device='cuda')
    out[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
    return out


def _nchw(t, c):
    return t[..., :c].float().cpu().permute(0, 3, 1, 2).contiguous()


@pytest.fixture(params=[0, 7], ids=['register-staged', 'tma-staged'])
def bn_tma(request):
    """Both implementations of the BatchNorm kernels: register-staged and TMA-staged (cp.async.bulk tile rings;
    tunable "bn_tma" bit mask: forward / backward reduce / backward apply)."""
    from margipose_b200._lib import lib
    assert lib().mp_set_tunable(b'bn_tma', request.param) == 0
    yield request.param
    lib().mp_set_tunable(b'bn_tma', 1)


@pytest.mark.parametrize('C,Cp,mode', [(128, 128, 'rb'), (17, 64, 'rb_logits'), (192, 192, 'act'),
                                       (64, 64, 'basic_id'), (128, 128, 'basic_down')])
def test_bn_forward_backward(C, Cp, mode, bn_tma):
    from margipose_b200 import ops
    gen = torch.Generator().manual_seed(C)
    n, h, w = 3, 24, 24      # 1728 pixels: several tiles per block, a ragged last tile
    ya = _bf(torch.randn(n, C, h, w, generator=gen) * 2 + 0.5)
    yb = _bf(torch.randn(n, C, h, w, generator=gen))
    res = _bf(torch.randn(n, C, h, w, generator=gen))
    dout = _bf(torch.randn(n, C, h, w, generator=gen))
    bna, bnb = torch.nn.BatchNorm2d(C), torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        for bn in (bna, bnb):
            bn.weight.copy_(torch.rand(C, generator=gen) + 0.5)
            bn.bias.copy_(torch.randn(C, generator=gen) * 0.1)
    ya_r, yb_r, res_r = ya.clone().requires_grad_(), yb.clone().requires_grad_(), res.clone().requires_grad_()
    if mode in ('rb', 'rb_logits'):
        want = F.relu(bna(ya_r)) + bnb(yb_r)
    elif mode == 'act':
        want = F.relu(bna(ya_r))
    elif mode == 'basic_id':
        want = F.relu(bna(ya_r) + res_r)
    else:
        want = F.relu(bna(ya_r) + bnb(yb_r))
    want.backward(dout)

    dev = torch.device('cuda')
    M = n * h * w

    def branch(y, bn):
        yg = _nhwc(y, Cp)
        yf = yg.float().reshape(M, Cp)
        return ops.BnBranchT(yg, bn.weight.detach().cuda(), bn.bias.detach().cuda(),
                             running_mean=torch.zeros(C, device=dev), running_var=torch.ones(C, device=dev),
                             sum=yf.sum(0).contiguous(), sq=(yf * yf).sum(0).contiguous(),
                             save_mean=torch.zeros(Cp, device=dev), save_invstd=torch.zeros(Cp, device=dev),
                             dy=torch.zeros(n, h, w, Cp, dtype=torch.bfloat16, device=dev),
                             dgamma=torch.zeros(C, device=dev), dbeta=torch.zeros(C, device=dev))
    a = branch(ya, bna)
    b = branch(yb, bnb) if mode in ('rb', 'rb_logits', 'basic_down') else None
    resg = _nhwc(res, Cp) if mode == 'basic_id' else None
    out = torch.zeros(n, h, w, Cp, dtype=torch.bfloat16, device=dev) if mode != 'rb_logits' else None
    out_nchw = torch.zeros(n, C, h, w, device=dev) if mode == 'rb_logits' else None
    relu_a = mode in ('rb', 'rb_logits', 'act')
    relu_out = mode in ('basic_id', 'basic_down')
    args = ops.bn_args(a, b, res=resg, relu_a=relu_a, relu_out=relu_out, out=out, out_nchw=out_nchw, C=C, hw=h * w)
    ops.bn_fwd(args, dev)
    torch.cuda.synchronize()
    if out is not None:
        torch.testing.assert_close(_nchw(out, C), want.detach(), **BF)
        if Cp > C:
            assert out[..., C:].abs().max().item() == 0
    else:
        torch.testing.assert_close(out_nchw.cpu(), want.detach(), **F32)
    torch.testing.assert_close(a.running_mean.cpu(), bna.running_mean, **F32)
    torch.testing.assert_close(a.running_var.cpu(), bna.running_var, **F32)

    # backward
    dres = torch.zeros(n, h, w, Cp, dtype=torch.bfloat16, device=dev) if mode == 'basic_id' else None
    sums = torch.zeros(4, Cp, device=dev)
    if mode == 'rb_logits':
        bargs = ops.bn_args(a, b, relu_a=relu_a, relu_out=relu_out, out=out, dout_nchw=dout.cuda(), sums=sums,
                            C=C, hw=h * w)
    else:
        bargs = ops.bn_args(a, b, res=resg, relu_a=relu_a, relu_out=relu_out, out=out, dout=_nhwc(dout, Cp),
                            dres=dres, sums=sums, C=C, hw=h * w)
    ops.bn_bwd(bargs, dev)
    torch.cuda.synchronize()
    torch.testing.assert_close(_nchw(a.dy, C), ya_r.grad, **BF)
    torch.testing.assert_close(a.dgamma.cpu(), bna.weight.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(a.dbeta.cpu(), bna.bias.grad, rtol=1e-3, atol=1e-3)
    if b is not None:
        torch.testing.assert_close(_nchw(b.dy, C), yb_r.grad, **BF)
        torch.testing.assert_close(b.dgamma.cpu(), bnb.weight.grad, rtol=1e-3, atol=1e-3)
    if dres is not None:
        torch.testing.assert_close(_nchw(dres, C), res_r.grad, **BF)


def test_bn_eval_mode_uses_running_stats():
    from margipose_b200 import ops
    gen = torch.Generator().manual_seed(1)
    n, C, h, w = 2, 64, 4, 4
    y = _bf(torch.randn(n, C, h, w, generator=gen))
    bn = torch.nn.BatchNorm2d(C).eval()
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(C, generator=gen))
        bn.running_var.copy_(torch.rand(C, generator=gen) + 0.5)
    a = ops.BnBranchT(_nhwc(y, C), bn.weight.detach().cuda(), bn.bias.detach().cuda(),
                      running_mean=bn.running_mean.cuda(), running_var=bn.running_var.cuda())
    out = torch.zeros(n, h, w, C, dtype=torch.bfloat16, device='cuda')
    ops.bn_fwd(ops.bn_args(a, relu_a=True, out=out, training=False), torch.device('cuda'))
    torch.testing.assert_close(_nchw(out, C), F.relu(bn(y)).detach(), **BF)


def test_maxpool_forward_backward_with_ties():
    from margipose_b200 import ops
    gen = torch.Generator().manual_seed(2)
    x = F.relu(_bf(torch.randn(2, 64, 16, 16, generator=gen)))   # ReLU output: many exact ties at 0
    xr = x.clone().requires_grad_()
    want = F.max_pool2d(xr, 3, 2, 1)
    dy = _bf(torch.randn(want.shape, generator=gen))
    want.backward(dy)
    y, idx = ops.maxpool_fwd(_nhwc(x, 64))
    dx = ops.maxpool_bwd(_nhwc(dy, 64), idx)
    torch.testing.assert_close(_nchw(y, 64), want.detach(), rtol=0, atol=0)
    torch.testing.assert_close(_nchw(dx, 64), xr.grad, **BF)


@pytest.mark.parametrize('mode', [1, 2])
def test_axis_permute_matches_reference_expression(mode):
    from margipose_b200 import ops
    gen = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(2, 192, 16, 16, generator=gen))
    perm = (0, 3, 2, 1) if mode == 1 else (0, 2, 1, 3)      # margipose_model.py:95,97
    want = torch.cat([t.permute(*perm) for t in x.split(16, -3)], -3)
    got = ops.axis_permute(_nhwc(x, 192), mode, 192)
    torch.testing.assert_close(_nchw(got, 192), want, rtol=0, atol=0)
    back = ops.axis_permute(got, mode, 192)
    torch.testing.assert_close(_nchw(back, 192), x, rtol=0, atol=0)


def test_combiner_forward_backward():
    from margipose_b200 import ops
    gen = torch.Generator().manual_seed(4)
    n, J, h, w, C = 2, 17, 32, 32, 128
    probs = [torch.softmax(torch.randn(n, J, h * w, generator=gen), -1).view(n, J, h, w) for _ in range(3)]
    wt = (torch.randn(C, 3 * J, 1, 1, generator=gen) * 0.2).requires_grad_()
    inp = _bf(torch.randn(n, C, h, w, generator=gen))
    pr = [p.clone().requires_grad_() for p in probs]
    want = inp + F.conv2d(torch.cat(pr, -3), wt)
    dout = _bf(torch.randn(n, C, h, w, generator=gen) * 1e-2)
    want.backward(dout)
    pg = [p.cuda() for p in probs]
    wg = wt.detach().reshape(C, 3 * J).contiguous().cuda()
    out = ops.combiner_fwd(pg, wg, _nhwc(inp, C))
    torch.testing.assert_close(_nchw(out, C), want.detach(), **BF)
    dps = [torch.full((n, J, h, w), 0.5, device='cuda') for _ in range(3)]
    dw = torch.zeros(C, 3 * J, device='cuda')
    ops.combiner_bwd(_nhwc(dout, C), pg, wg, dps, dw, accumulate=True)
    for k in range(3):
        torch.testing.assert_close(dps[k].cpu() - 0.5, pr[k].grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(dw.cpu(), wt.grad.reshape(C, 3 * J), rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize('h,w', [(64, 64), (48, 80), (24, 16)])
def test_stem_im2col_conv_equals_7x7_stride2(h, w):
    """The stem gather (a block per 32 output pixels of a row; ragged last tile, images narrower than a tile) followed by
    the 1x1 GEMM is the 7x7 stride-2 convolution (margipose_model.py:130); the uint8 NHWC variant with /255 and the
    ImageNet normalisation fused gives the same patches as the fp32 path fed the normalised image."""
    from margipose_b200 import ops, convops as C
    from margipose_b200._lib import lib, check, stream_ptr
    import ctypes
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, h, w, generator=gen)
    wt = _bf(torch.randn(64, 3, 7, 7, generator=gen) / 147 ** 0.5)
    want = F.conv2d(_bf(x), wt, None, 2, 3)
    patches = ops.stem_im2col(x.cuda())
    assert patches[..., 147:].abs().max().item() == 0
    g = C.ConvGeom(147, 64, 1)
    master = wt.permute(0, 2, 3, 1).reshape(64, 1, 147).contiguous().cuda()
    out = torch.zeros(2, h // 2, w // 2, 64, dtype=torch.bfloat16, device='cuda')
    C.conv_forward(g, patches, C.pack_fwd(g, master), out)
    torch.testing.assert_close(_nchw(out, 64), want, rtol=1e-2, atol=2e-2)
    # uint8 NHWC input
    img = torch.randint(0, 256, (2, h, w, 3), generator=gen, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    xn = ((img.float() / 255 - torch.tensor(mean)) / torch.tensor(std)).permute(0, 3, 1, 2).contiguous()
    p8 = torch.empty_like(patches)
    dev = torch.device('cuda')
    imgc = img.cuda()
    check(lib().mp_stem_im2col_u8(imgc.data_ptr(), p8.data_ptr(), (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std),
                                  2, h, w, 0, stream_ptr(dev)), 'mp_stem_im2col_u8')
    torch.cuda.synchronize()
    torch.testing.assert_close(p8.float(), ops.stem_im2col(xn.cuda()).float(), rtol=1e-2, atol=1e-2)


def test_add_and_sgd():
    from margipose_b200 import ops
    gen = torch.Generator().manual_seed(6)
    ts = [_bf(torch.randn(4096, generator=gen)) for _ in range(4)]
    got = ops.add_bf16([t.to(torch.bfloat16).cuda() for t in ts])
    torch.testing.assert_close(got.float().cpu(), _bf(sum(ts)), **BF)
    p = torch.randn(1000, generator=gen)
    pr = p.clone().requires_grad_()
    opt = torch.optim.SGD([pr], lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=True)
    pg, buf = p.cuda(), torch.zeros(1000, device='cuda')
    for step in range(3):
        g = torch.randn(1000, generator=gen)
        pr.grad = g.clone()
        opt.step()
        ops.sgd_step(pg, g.cuda(), buf, 0.1, momentum=0.9, weight_decay=1e-4, nesterov=True,
                     first_step=(step == 0))
    torch.testing.assert_close(pg.cpu(), pr.detach(), rtol=1e-5, atol=1e-6)
