"""CUDA-backed mirror of the reference's `margipose.dsntnn` functional API.

Same names, argument meaning and error behaviour as /root/reference/src/margipose/dsntnn.py
(flat_softmax :124, dsnt :84, js_reg_losses :220, euclidean_losses :133, average_loss :99,
make_gauss :154), for (B, J, H, W) fp32 CUDA heatmaps.  Every function is one launch of the
fused sm_100a tail kernels through the C ABI (include/margipose_b200.h); gradients are the
hand-derived backward kernel, not an autograd graph.  There is no CPU
path: CPU tensors raise.

`fused_tail_losses` / `fused_tail_from_logits` expose the whole-stage fusion the model uses.
"""
import torch

from . import _lib
from ._lib import lib, check, ptr, planes, stream_ptr, require_cuda


def _prep(t, name, dims=None):
    require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError('%s must be float32 (got %s)' % (name, t.dtype))
    if dims is not None and t.dim() != dims:
        raise ValueError('%s must be %d-dimensional (got %d)' % (name, dims, t.dim()))
    return t.contiguous()


def _tail_fwd(ins, from_logits, prob=None, ab=None, js=None, mu=None, target=None,
              valid_depth=None, coords=None, loss=None, accumulate=False, pixelwise=True,
              sigma=1.0):
    ref = next(t for t in ins if t is not None)
    B, J, H, W = ref.shape
    check(lib().mp_tail_fwd(planes(ins), int(from_logits), planes(prob), planes(ab), planes(js),
                            planes(mu), ptr(target), ptr(valid_depth), ptr(coords), ptr(loss),
                            int(accumulate), int(pixelwise), float(sigma), B, J, H, W,
                            stream_ptr(ref.device)), 'mp_tail_fwd')


def _tail_bwd(prob, gup, out, target=None, coords=None, w=None, valid_depth=None, mu=None,
              coef=None, project=False, pixelwise=True, sigma=1.0):
    ref = next(t for t in prob if t is not None)
    B, J, H, W = ref.shape
    check(lib().mp_tail_bwd(planes(prob), planes(gup), planes(out), ptr(target), ptr(coords),
                            ptr(w), ptr(valid_depth), planes(mu), planes(coef), int(project),
                            int(pixelwise), float(sigma), B, J, H, W, stream_ptr(ref.device)),
          'mp_tail_bwd')


class _FlatSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        p = torch.empty_like(z)
        _tail_fwd([z, None, None], True, prob=[p, None, None])
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        dz = torch.empty_like(p)
        _tail_bwd([p, None, None], [g.contiguous(), None, None], [dz, None, None], project=True)
        return dz


def flat_softmax(inp):
    """Softmax over the flattened spatial dims of a (B, J, H, W) tensor (dsntnn.py:124-130)."""
    require_cuda(inp)
    shape = inp.shape
    if inp.dim() < 3:
        raise ValueError('flat_softmax expects at least 3 dims')
    z = _prep(inp.reshape(shape[0], shape[1], 1, -1) if inp.dim() != 4 else inp, 'inp', 4)
    return _FlatSoftmax.apply(z).reshape(shape)


class _Dsnt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p):
        B, J = p.shape[:2]
        ab = torch.empty(B, J, 2, device=p.device, dtype=torch.float32)
        _tail_fwd([p, None, None], False, ab=[ab, None, None])
        ctx.save_for_backward(p)
        return ab

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        coef = torch.cat([torch.zeros_like(g[..., :1]), g], -1).contiguous()
        dp = torch.empty_like(p)
        _tail_bwd([p, None, None], None, [dp, None, None], coef=[coef, None, None], project=False)
        return dp


def dsnt(heatmaps):
    """Soft-argmax: (B, J, H, W) probabilities -> (B, J, 2) = (x from last dim, y from rows)
    (dsntnn.py:84-96)."""
    return _Dsnt.apply(_prep(heatmaps, 'heatmaps', 4))


class _JsReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, mu, sigma):
        B, J = p.shape[:2]
        js = torch.empty(B, J, device=p.device, dtype=torch.float32)
        _tail_fwd([p, None, None], False, js=[js, None, None], mu=[mu, None, None], sigma=sigma)
        ctx.save_for_backward(p, mu)
        ctx.sigma = sigma
        return js

    @staticmethod
    def backward(ctx, g):
        p, mu = ctx.saved_tensors
        coef = torch.stack([g, torch.zeros_like(g), torch.zeros_like(g)], -1).contiguous()
        dp = torch.empty_like(p)
        _tail_bwd([p, None, None], None, [dp, None, None], mu=[mu, None, None],
                  coef=[coef, None, None], project=False, sigma=ctx.sigma)
        return dp, None, None


def js_reg_losses(heatmaps, mu_t, sigma_t):
    """Per-location Jensen-Shannon divergence to a Gaussian target (dsntnn.py:220-232)."""
    p = _prep(heatmaps, 'heatmaps', 4)
    mu = _prep(mu_t, 'mu_t')
    ndims = mu.size(-1)
    assert p.dim() == ndims + 2, 'expected heatmaps to be a {}D tensor'.format(ndims + 2)
    assert p.size()[:-ndims] == mu.size()[:-1]
    if mu.requires_grad:
        raise NotImplementedError('gradients w.r.t. the Gaussian means are not on the hot path')
    return _JsReg.apply(p, mu, float(sigma_t))


class _Euclid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, t):
        n, d = a.numel() // a.shape[-1], a.shape[-1]
        out = torch.empty(a.shape[:-1], device=a.device, dtype=torch.float32)
        check(lib().mp_euclid_fwd(ptr(a), ptr(t), n, d, ptr(out), stream_ptr(a.device)),
              'mp_euclid_fwd')
        ctx.save_for_backward(a, t, out)
        return out

    @staticmethod
    def backward(ctx, g):
        a, t, dist = ctx.saved_tensors
        n, d = a.numel() // a.shape[-1], a.shape[-1]
        ga = torch.empty_like(a)
        check(lib().mp_euclid_bwd(ptr(g.contiguous()), ptr(a), ptr(t), ptr(dist), n, d, ptr(ga),
                                  stream_ptr(a.device)), 'mp_euclid_bwd')
        return ga, (-ga if ctx.needs_input_grad[1] else None)


def euclidean_losses(actual, target):
    """Euclidean distance over the last dim (dsntnn.py:133-151)."""
    assert actual.size() == target.size(), 'input tensors must have the same size'
    return _Euclid.apply(_prep(actual, 'actual'), _prep(target, 'target'))


class _MaskedMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, losses, mask):
        out2 = torch.empty(2, device=losses.device, dtype=torch.float32)
        check(lib().mp_masked_mean_fwd(ptr(losses), ptr(mask), losses.numel(), ptr(out2),
                                       stream_ptr(losses.device)), 'mp_masked_mean_fwd')
        ctx.save_for_backward(mask, out2)
        ctx.shape = losses.shape
        return out2[0].clone()

    @staticmethod
    def backward(ctx, g):
        mask, out2 = ctx.saved_tensors
        gl = torch.empty(ctx.shape, device=out2.device, dtype=torch.float32)
        g = g.reshape(1).contiguous()
        check(lib().mp_masked_mean_bwd(ptr(g), ptr(mask), ptr(out2), gl.numel(), ptr(gl),
                                       stream_ptr(out2.device)), 'mp_masked_mean_bwd')
        return gl, None


def average_loss(losses, mask=None):
    """Masked mean with the denominator clamped to >= 1 (dsntnn.py:99-121)."""
    losses = _prep(losses, 'losses')
    if mask is not None:
        assert mask.size() == losses.size(), 'mask must be the same size as losses'
        mask = _prep(mask, 'mask')
    return _MaskedMean.apply(losses, mask)


def make_gauss(means, size, sigma, normalize=True):
    """Renders 2D Gaussians; means (..., 2) = (x, y) normalised, size (H, W), sigma in pixels
    (dsntnn.py:154-195).  Forward only on this path (targets do not require grad)."""
    if len(size) != 2 or means.size(-1) != 2:
        raise NotImplementedError('only 2D heatmaps are on the MargiPose hot path')
    mu = _prep(means, 'means')
    H, W = int(size[0]), int(size[1])
    out = torch.empty(*mu.shape[:-1], H, W, device=mu.device, dtype=torch.float32)
    check(lib().mp_make_gauss(ptr(mu), ptr(out), int(bool(normalize)), float(sigma),
                              mu.numel() // 2, H, W, stream_ptr(mu.device)), 'mp_make_gauss')
    return out


class _FusedTailLosses(torch.autograd.Function):
    """probabilities (3 planes) + xyz targets -> per-joint stage loss and xyz coordinates."""

    @staticmethod
    def forward(ctx, xy, zy, xz, target, valid_depth, pixelwise, sigma):
        B, J = xy.shape[:2]
        coords = torch.empty(B, J, 3, device=xy.device, dtype=torch.float32)
        loss = torch.empty(B, J, device=xy.device, dtype=torch.float32)
        _tail_fwd([xy, zy, xz], False, target=target, valid_depth=valid_depth, coords=coords,
                  loss=loss, pixelwise=pixelwise, sigma=sigma)
        ctx.save_for_backward(xy, zy, xz, target, coords, valid_depth)
        ctx.cfg = (pixelwise, sigma)
        ctx.mark_non_differentiable(coords)
        return loss, coords

    @staticmethod
    def backward(ctx, g_loss, _g_coords):
        xy, zy, xz, target, coords, valid_depth = ctx.saved_tensors
        pixelwise, sigma = ctx.cfg
        outs = [torch.empty_like(xy), torch.empty_like(zy), torch.empty_like(xz)]
        _tail_bwd([xy, zy, xz], None, outs, target=target, coords=coords,
                  w=g_loss.contiguous(), valid_depth=valid_depth, project=False,
                  pixelwise=pixelwise, sigma=sigma)
        return outs[0], outs[1], outs[2], None, None, None, None


def fused_tail_losses(xy_hm, zy_hm, xz_hm, target_xyz, valid_depth=None, pixelwise=True,
                      sigma=1.0):
    """One stage of models/margipose_model.py:236-252 (or :223-234 where valid_depth == 0) in a
    single launch.  Returns (losses (B, J), coords (B, J, 3))."""
    xy, zy, xz = (_prep(t, 'heatmap', 4) for t in (xy_hm, zy_hm, xz_hm))
    target = _prep(target_xyz, 'target')
    if valid_depth is not None:
        require_cuda(valid_depth)
        valid_depth = valid_depth.to(torch.int32).contiguous()
    return _FusedTailLosses.apply(xy, zy, xz, target, valid_depth, bool(pixelwise), float(sigma))


class _FusedTailFromLogits(torch.autograd.Function):
    """logits (3 planes) + targets -> probabilities, coordinates, per-joint loss: the K4/K5
    fusion of SURVEY.md section 2b (softmax + expectations + xyz + Gaussian + JS x3 + Euclid)."""

    @staticmethod
    def forward(ctx, zxy, zzy, zxz, target, valid_depth, pixelwise, sigma):
        B, J = zxy.shape[:2]
        probs = [torch.empty_like(zxy), torch.empty_like(zzy), torch.empty_like(zxz)]
        coords = torch.empty(B, J, 3, device=zxy.device, dtype=torch.float32)
        loss = torch.empty(B, J, device=zxy.device, dtype=torch.float32)
        _tail_fwd([zxy, zzy, zxz], True, prob=probs, target=target, valid_depth=valid_depth,
                  coords=coords, loss=loss, pixelwise=pixelwise, sigma=sigma)
        ctx.save_for_backward(probs[0], probs[1], probs[2], target, coords, valid_depth)
        ctx.cfg = (pixelwise, sigma)
        ctx.mark_non_differentiable(coords)
        return probs[0], probs[1], probs[2], coords, loss

    @staticmethod
    def backward(ctx, g_xy, g_zy, g_xz, _g_coords, g_loss):
        xy, zy, xz, target, coords, valid_depth = ctx.saved_tensors
        pixelwise, sigma = ctx.cfg
        outs = [torch.empty_like(xy), torch.empty_like(zy), torch.empty_like(xz)]
        gup = [g.contiguous() if g is not None else None for g in (g_xy, g_zy, g_xz)]
        w = g_loss.contiguous() if g_loss is not None else torch.zeros_like(coords[..., 0])
        _tail_bwd([xy, zy, xz], gup, outs, target=target, coords=coords, w=w,
                  valid_depth=valid_depth, project=True, pixelwise=pixelwise, sigma=sigma)
        return outs[0], outs[1], outs[2], None, None, None, None


def fused_tail_from_logits(logits_xy, logits_zy, logits_xz, target_xyz, valid_depth=None,
                           pixelwise=True, sigma=1.0):
    """Returns (p_xy, p_zy, p_xz, coords (B,J,3), losses (B,J)) from raw column outputs."""
    zs = [_prep(t, 'logits', 4) for t in (logits_xy, logits_zy, logits_xz)]
    target = _prep(target_xyz, 'target')
    if valid_depth is not None:
        require_cuda(valid_depth)
        valid_depth = valid_depth.to(torch.int32).contiguous()
    return _FusedTailFromLogits.apply(zs[0], zs[1], zs[2], target, valid_depth, bool(pixelwise),
                                      float(sigma))


def heatmaps_to_coords(xy_hm, zy_hm, xz_hm):
    """models/margipose_model.py:254-261 in one launch (differentiable)."""
    return _HeatmapsToCoords.apply(*(_prep(t, 'heatmap', 4) for t in (xy_hm, zy_hm, xz_hm)))


class _HeatmapsToCoords(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xy, zy, xz):
        B, J = xy.shape[:2]
        coords = torch.empty(B, J, 3, device=xy.device, dtype=torch.float32)
        _tail_fwd([xy, zy, xz], False, coords=coords)
        ctx.save_for_backward(xy, zy, xz)
        return coords

    @staticmethod
    def backward(ctx, g):
        xy, zy, xz = ctx.saved_tensors
        z = torch.zeros_like(g[..., 0])
        # x = a_xy, y = b_xy, z = (a_zy + b_xz) / 2
        coef = [torch.stack([z, g[..., 0], g[..., 1]], -1).contiguous(),
                torch.stack([z, 0.5 * g[..., 2], z], -1).contiguous(),
                torch.stack([z, z, 0.5 * g[..., 2]], -1).contiguous()]
        outs = [torch.empty_like(xy), torch.empty_like(zy), torch.empty_like(xz)]
        _tail_bwd([xy, zy, xz], None, outs, coef=coef, project=False)
        return outs[0], outs[1], outs[2]
