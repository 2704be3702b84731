"""Oracle for the differentiable tail: spatial softmax, soft-argmax, marginal
combination, Gaussian rendering, Jensen-Shannon regulariser, Euclidean loss.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Written from the maths in
SURVEY.md Appendix B; each function names the reference lines it restates
(paths relative to /root/reference/src/margipose/).  Everything is built from
differentiable torch ops so torch autograd supplies the oracle gradients.
"""
import torch

EPS = 1e-24  # dsntnn.py:199 (KL epsilon) and dsntnn.py:194 (Gaussian normaliser)


def pixel_centres(n, dtype=torch.float32, device=None):
    """c_i = (2 i + 1) / n - 1, the cell centres of n cells spanning [-1, 1].

    Restates dsntnn.py:12-36 (`_normalized_linspace`), keeping its evaluation
    order `i * (2/n) + (-(n-1)/n)` so fp32 values agree bit-for-bit.
    """
    first = -(n - 1.0) / n
    return torch.arange(n, dtype=dtype, device=device) * (2.0 / n) + first


def flat_softmax(logits):
    """Softmax over all trailing (spatial) dims of a (B, J, ...) tensor.

    Restates dsntnn.py:124-130.
    """
    b, j = logits.shape[:2]
    return torch.softmax(logits.reshape(b * j, -1), dim=-1).reshape(logits.shape)


def dsnt(heatmaps):
    """Soft-argmax of (B, J, H, W) probabilities -> (B, J, 2) = (col coord, row coord).

    Restates dsntnn.py:84-96 / :39-62: marginalise over the other axis, then
    take the expectation of the pixel-centre coordinate.
    """
    h, w = heatmaps.shape[-2:]
    cw = pixel_centres(w, heatmaps.dtype, heatmaps.device)
    ch = pixel_centres(h, heatmaps.dtype, heatmaps.device)
    col = (heatmaps.sum(-2) * cw).sum(-1)
    row = (heatmaps.sum(-1) * ch).sum(-1)
    return torch.stack([col, row], -1)


def heatmaps_to_coords(xy_hm, zy_hm, xz_hm):
    """(x, y) from the xy plane; z = mean of zy's column coord and xz's row coord.

    Restates models/margipose_model.py:254-261.
    """
    xy = dsnt(xy_hm)
    zy = dsnt(zy_hm)
    xz = dsnt(xz_hm)
    z = 0.5 * (zy[..., 0] + xz[..., 1])
    return torch.stack([xy[..., 0], xy[..., 1], z], -1)


def make_gauss(means, size, sigma, normalize=True):
    """Separable Gaussians centred at `means` (B, J, 2)=(col, row) on an (H, W) grid.

    Restates dsntnn.py:154-195: exponent k * (c - mu)^2 with
    k = -0.5 * (n / (2 sigma))^2 per axis, product of the two axis factors,
    divided by (sum + 1e-24) when normalising.
    """
    h, w = size
    cw = pixel_centres(w, means.dtype, means.device)
    ch = pixel_centres(h, means.dtype, means.device)
    kw = -0.5 * (1.0 / (2.0 * sigma / w)) ** 2
    kh = -0.5 * (1.0 / (2.0 * sigma / h)) ** 2
    ew = (((cw - means[..., 0:1]) ** 2) * kw).exp()          # (B, J, W)
    eh = (((ch - means[..., 1:2]) ** 2) * kh).exp()          # (B, J, H)
    g = eh.unsqueeze(-1) * ew.unsqueeze(-2)                  # (B, J, H, W)
    if not normalize:
        return g
    total = g.sum(-1, keepdim=True).sum(-2, keepdim=True) + EPS
    return g / total


def _kl(p, q):
    # dsntnn.py:198-202
    return (p * ((p + EPS).log() - (q + EPS).log())).sum(-1).sum(-1)


def js_reg_losses(heatmaps, mu_t, sigma_t):
    """Jensen-Shannon divergence between each heatmap and its target Gaussian.

    Restates dsntnn.py:205-232.
    """
    g = make_gauss(mu_t, heatmaps.shape[-2:], sigma_t)
    m = 0.5 * (heatmaps + g)
    return 0.5 * _kl(heatmaps, m) + 0.5 * _kl(g, m)


def euclidean_losses(actual, target):
    """L2 distance over the last dim. Restates dsntnn.py:133-151."""
    return ((actual - target) ** 2).sum(-1).sqrt()


def average_loss(losses, mask=None):
    """Masked mean with the denominator clamped to >= 1. Restates dsntnn.py:99-121."""
    if mask is None:
        return losses.sum() / max(losses.numel(), 1)
    return (losses * mask).sum() / mask.sum().clamp(1)


def losses_3d(xy_hms, zy_hms, xz_hms, target, pixelwise_loss='jsd', sigma=1.0):
    """Per-joint 3D loss summed over stages. Restates models/margipose_model.py:236-252."""
    t = target[..., :3]
    t_xy = t[..., [0, 1]]
    t_zy = t[..., [2, 1]]
    t_xz = t[..., [0, 2]]
    total = 0
    for xy, zy, xz in zip(xy_hms, zy_hms, xz_hms):
        if pixelwise_loss == 'jsd':
            total = total + js_reg_losses(xy, t_xy, sigma)
            total = total + js_reg_losses(zy, t_zy, sigma)
            total = total + js_reg_losses(xz, t_xz, sigma)
        elif pixelwise_loss is not None:
            raise Exception('unrecognised pixelwise loss: {}'.format(pixelwise_loss))
        total = total + euclidean_losses(heatmaps_to_coords(xy, zy, xz), t)
    return total


def losses_2d(xy_hms, zy_hms, xz_hms, target, pixelwise_loss='jsd', sigma=1.0):
    """Per-joint 2D loss summed over stages. Restates models/margipose_model.py:223-234."""
    t_xy = target[..., :2]
    total = 0
    for xy, zy, xz in zip(xy_hms, zy_hms, xz_hms):
        if pixelwise_loss == 'jsd':
            total = total + js_reg_losses(xy, t_xy, sigma)
        elif pixelwise_loss is not None:
            raise Exception('unrecognised pixelwise loss: {}'.format(pixelwise_loss))
        total = total + euclidean_losses(heatmaps_to_coords(xy, zy, xz)[..., :2], t_xy)
    return total


def forward_loss(xy_hms, zy_hms, xz_hms, target, mask, valid_depth, pixelwise_loss='jsd'):
    """The caller contract of bin/train_3d.py:126-142 (3D / 2D / mixed batches)."""
    target = target[..., :3]
    vd = [int(v) for v in valid_depth]
    if 0 not in vd:
        losses = losses_3d(xy_hms, zy_hms, xz_hms, target, pixelwise_loss)
    elif 1 not in vd:
        losses = losses_2d(xy_hms, zy_hms, xz_hms, target, pixelwise_loss)
    else:
        l3 = losses_3d(xy_hms, zy_hms, xz_hms, target, pixelwise_loss)
        l2 = losses_2d(xy_hms, zy_hms, xz_hms, target, pixelwise_loss)
        sel = torch.tensor(vd, dtype=torch.bool).unsqueeze(-1)
        losses = torch.where(sel, l3, l2)
    return average_loss(losses, mask)
