"""CPU emulation of the mp_conv_igemm / mp_conv_wgrad launch contracts (include/margipose_b200.h)
-- TEST INFRASTRUCTURE.  It executes the SAME tap tables / 5-D views the host code in
margipose_b200/convops.py hands to the CUDA kernels, in plain fp32 torch on CPU, so the host-side
geometry (shifts, parities, K offsets, scatter strides) is checked without a GPU."""
import torch


def _view5(t, parity):
    """(N, H', P, W', C') fp32 tensor following the view convention of convops._view."""
    n, h, w, c = t.shape
    t = t.float()
    if not parity:
        return t.reshape(n, h, 1, w, c)
    return t.reshape(n, h // 2, 2, w // 2, 2 * c)


def _gather(v, c0, nc, dw, p, dh, out_h, out_w):
    """v[n, h+dh, p, w+dw, c0:c0+nc] over an (out_h, out_w) grid with zero fill out of bounds."""
    n, hh, _, ww, cc = v.shape
    out = torch.zeros(n, out_h, out_w, nc)
    h_lo, h_hi = max(0, -dh), min(out_h, hh - dh)
    w_lo, w_hi = max(0, -dw), min(out_w, ww - dw)
    c_hi = min(c0 + nc, cc)
    if h_lo < h_hi and w_lo < w_hi and c0 < c_hi:
        out[:, h_lo:h_hi, w_lo:w_hi, :c_hi - c0] = \
            v[:, h_lo + dh:h_hi + dh, p, w_lo + dw:w_hi + dw, c0:c_hi]
    return out


def _strided(t, n_img, out_h, out_w, out_c, strides, off):
    # (as_strided takes an ABSOLUTE storage offset: tensors here may be views into a larger hi / lo buffer)
    return torch.as_strided(t, (n_img, out_h, out_w, out_c), tuple(strides) + (1,), t.storage_offset() + off)


def igemm(srcs, wmat, taps, cblocks, n_img, out_h, out_w, out, out_strides, out_c, res=None,
          stats=None, out_offset=0, bn=None, defer=None, ep=None, out_lo=None, acc_in=None, res_lo=None):
    views = [_view5(t, parity) for t, parity in srcs]
    wm = wmat.float()
    acc = torch.zeros(n_img, out_h, out_w, wm.shape[0])
    nc = cblocks * 64
    for src, c0, dw, p, dh, koff in taps:
        a = _gather(views[src], c0, nc, dw, p, dh, out_h, out_w)
        acc += a @ wm[:, koff:koff + nc].t()
    acc = acc[..., :out_c]
    sn, sh, sw = out_strides
    shape = (n_img, out_h, out_w, out_c, out_strides, out_offset)
    dst = _strided(out, *shape)
    shape = (n_img, out_h, out_w, out_c, out_strides, out_offset)
    if acc_in is not None:     # split mode: the pair written by the earlier passes (acc_in is `out` itself)
        assert acc_in.data_ptr() == out.data_ptr() and out_lo is not None
        acc = acc + _strided(out, *shape).float() + _strided(out_lo, *shape).float()
    if ep is not None:     # (scale, shift, relu mode) as fp32 tensors here (device pointers on the GPU)
        scale, shift, relu = ep
        acc = acc * scale[:out_c] + shift[:out_c]
        if relu == 1:
            acc = acc.clamp_min(0)
    if res is not None:
        acc = acc + _strided(res, *shape).float()
        if res_lo is not None:
            acc = acc + _strided(res_lo, *shape).float()
    if ep is not None and ep[2] == 2:
        acc = acc.clamp_min(0)
    rounded = acc.to(torch.bfloat16)
    dst.copy_(rounded)
    stored = rounded.float()
    if out_lo is not None:
        lo = (acc - stored).to(torch.bfloat16)
        _strided(out_lo, *shape).copy_(lo)
        stored = stored + lo.float()
    if stats is not None:
        r = stored.reshape(-1, out_c)
        stats[0][:out_c] += r.sum(0)
        stats[1][:out_c] += (r * r).sum(0)


def wgrad(a_t, b_t, b_parity, taps, m_real, n_real, n_cols, n_slots, n_img, grid_h, grid_w, dw, a_lo=None, b_lo=None):
    # split mode: a * b + a_lo * b + a * b_lo (the lo * lo term is dropped, as on the GPU)
    for at, bt in [(a_t, b_t)] + ([(a_lo, b_t), (a_t, b_lo)] if a_lo is not None else []):
        va = _view5(at, False)
        vb = _view5(bt, b_parity)
        a = _gather(va, 0, va.shape[-1], 0, 0, 0, grid_h, grid_w)[..., :m_real].reshape(-1, m_real)
        for _src, c0, ddw, p, dh, slot in taps:
            b = _gather(vb, c0, n_cols, ddw, p, dh, grid_h, grid_w)[..., :n_real].reshape(-1, n_real)
            dw[:, slot, :] += a.t() @ b


def install(monkeypatch):
    from margipose_b200 import convops
    monkeypatch.setattr(convops, '_igemm_one', igemm)
    monkeypatch.setattr(convops, '_wgrad_one', wgrad)
