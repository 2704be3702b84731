"""ctypes binding of the C ABI declared in include/margipose_b200.h.

The product path has NO fallback: if the shared library is missing or a call fails, we raise.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libmargipose_b200.so')

c_float_p = ctypes.c_void_p   # device pointers travel as integers
c_int = ctypes.c_int
c_double = ctypes.c_double
c_void_p = ctypes.c_void_p
c_size_t = ctypes.c_size_t
PlaneTable = ctypes.c_void_p * 3

MP_MAX_TAPS = 30


class View5(ctypes.Structure):
    _fields_ = [('ptr', c_void_p), ('dim', ctypes.c_int64 * 5), ('stride', ctypes.c_int64 * 5)]


class Tap(ctypes.Structure):
    _fields_ = [('src', ctypes.c_int32), ('c0', ctypes.c_int32), ('dw', ctypes.c_int32),
                ('p', ctypes.c_int32), ('dh', ctypes.c_int32), ('koff', ctypes.c_int32)]


class IgemmArgs(ctypes.Structure):
    _fields_ = [('src', View5 * 2), ('wmat', c_void_p), ('w_rows', ctypes.c_int64),
                ('w_k', ctypes.c_int64), ('n_taps', ctypes.c_int32), ('taps', Tap * MP_MAX_TAPS),
                ('cblocks', ctypes.c_int32), ('n_img', ctypes.c_int32), ('out_h', ctypes.c_int32),
                ('out_w', ctypes.c_int32), ('out', c_void_p), ('res', c_void_p),
                ('out_sn', ctypes.c_int64), ('out_sh', ctypes.c_int64), ('out_sw', ctypes.c_int64),
                ('out_c', ctypes.c_int32), ('stat_sum', c_void_p), ('stat_sq', c_void_p),
                ('stat_replicas', ctypes.c_int32), ('stat_stride', ctypes.c_int64),
                ('bn', c_void_p), ('bn_counter', c_void_p), ('bn_launches', ctypes.c_int32),
                ('bn_channels', ctypes.c_int32), ('bn_count', ctypes.c_int64),
                ('bn_momentum', ctypes.c_float), ('bn_eps', ctypes.c_float),
                ('ep_scale', c_void_p), ('ep_shift', c_void_p), ('ep_relu', ctypes.c_int32),
                ('acc_in', c_void_p), ('lo_delta', ctypes.c_int64)]


class WgradArgs(ctypes.Structure):
    _fields_ = [('a', View5), ('b', View5), ('n_taps', ctypes.c_int32), ('taps', Tap * MP_MAX_TAPS),
                ('m_real', ctypes.c_int32), ('n_real', ctypes.c_int32), ('n_cols', ctypes.c_int32),
                ('n_slots', ctypes.c_int32), ('n_img', ctypes.c_int32), ('grid_h', ctypes.c_int32),
                ('grid_w', ctypes.c_int32), ('n_off', ctypes.c_int32), ('dw', c_void_p),
                ('a_lo', View5), ('b_lo', View5)]


class BnBranch(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ('y', 'sum', 'sq', 'gamma', 'beta', 'running_mean',
                                        'running_var', 'save_mean', 'save_invstd', 'conv_bias',
                                        'scale', 'shift', 'coef', 'dy', 'dgamma', 'dbeta')]


class BnArgs(ctypes.Structure):
    _fields_ = [('a', BnBranch), ('b', BnBranch), ('res', c_void_p), ('relu_a', ctypes.c_int32),
                ('relu_out', ctypes.c_int32), ('out', c_void_p), ('out_nchw', c_void_p),
                ('dout', c_void_p), ('dout_nchw', c_void_p), ('dres', c_void_p), ('sums', c_void_p), ('bwd_counter', c_void_p),
                ('stat_replicas', ctypes.c_int32), ('stat_stride', ctypes.c_int64),
                ('M', ctypes.c_int64), ('C', ctypes.c_int32), ('Cp', ctypes.c_int32),
                ('HW', ctypes.c_int32), ('training', ctypes.c_int32), ('momentum', ctypes.c_float),
                ('eps', ctypes.c_float), ('lo_delta', ctypes.c_int64)]


class BnFoldEntry(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ('gamma', 'beta', 'running_mean', 'running_var', 'conv_bias',
                                        'scale', 'shift')] + \
               [('C', ctypes.c_int32), ('Cp', ctypes.c_int32), ('eps', ctypes.c_float),
                ('reserved', ctypes.c_int32)]


class PackEntry(ctypes.Structure):
    _fields_ = [('src_off', ctypes.c_int64), ('dst_off', ctypes.c_int64),
                ('dst_row_stride', ctypes.c_int64), ('work_off', ctypes.c_int64),
                ('work_end', ctypes.c_int64), ('A', ctypes.c_int32), ('B', ctypes.c_int32),
                ('taps', ctypes.c_int32), ('transpose', ctypes.c_int32), ('rows_p', ctypes.c_int32),
                ('cols_p', ctypes.c_int32), ('lo_off', ctypes.c_int64)]


_lib = None


class MargiposeB200Error(RuntimeError):
    pass


def _signatures():
    P, I, D, PT = c_void_p, c_int, c_double, ctypes.POINTER(PlaneTable)
    return {
        'mp_abi_version': (I, []),
        'mp_last_error': (ctypes.c_char_p, []),
        'mp_tail_fwd': (I, [PT, I, PT, PT, PT, PT, P, P, P, P, I, I, D, I, I, I, I, P]),
        'mp_tail_bwd': (I, [PT, PT, PT, P, P, P, P, PT, PT, I, I, D, I, I, I, I, P]),
        'mp_masked_mean_fwd': (I, [P, P, I, P, P]),
        'mp_masked_mean_bwd': (I, [P, P, P, I, P, P]),
        'mp_euclid_fwd': (I, [P, P, I, I, P, P]),
        'mp_euclid_bwd': (I, [P, P, P, P, I, I, P, P]),
        'mp_make_gauss': (I, [P, P, I, D, I, I, I, P]),
        'mp_conv_igemm': (I, [ctypes.POINTER(IgemmArgs), P]),
        'mp_conv_igemm_grouped': (I, [ctypes.POINTER(IgemmArgs), I, P]),
        'mp_conv_wgrad': (I, [ctypes.POINTER(WgradArgs), P]),
        'mp_conv_wgrad_grouped': (I, [ctypes.POINTER(WgradArgs), I, P]),
        'mp_set_tunable': (I, [ctypes.c_char_p, ctypes.c_int64]),
        'mp_bn_fold_eval': (I, [P, I, P]),
        'mp_bn_stats': (I, [ctypes.POINTER(BnArgs), P]),
        'mp_bn_stats_grouped': (I, [ctypes.POINTER(BnArgs), I, P]),
        'mp_bn_fwd': (I, [ctypes.POINTER(BnArgs), P]),
        'mp_bn_bwd_reduce': (I, [ctypes.POINTER(BnArgs), P]),
        'mp_bn_bwd_apply': (I, [ctypes.POINTER(BnArgs), P]),
        'mp_bn_fwd_grouped': (I, [ctypes.POINTER(BnArgs), I, P]),
        'mp_bn_bwd_reduce_grouped': (I, [ctypes.POINTER(BnArgs), I, P]),
        'mp_bn_bwd_apply_grouped': (I, [ctypes.POINTER(BnArgs), I, P]),
        'mp_maxpool_fwd': (I, [P, P, P, I, I, I, I, ctypes.c_int64, P]),
        'mp_maxpool_bwd': (I, [P, P, P, I, I, I, I, P]),
        'mp_axis_permute': (I, [P, P, I, I, I, I, I, P]),
        'mp_combiner_fwd': (I, [PT, P, P, P, I, I, I, I, ctypes.c_int64, P]),
        'mp_combiner_bwd': (I, [P, PT, P, PT, P, I, I, I, I, I, ctypes.c_int64, P]),
        'mp_stem_im2col': (I, [P, P, I, I, I, ctypes.c_int64, P]),
        'mp_stem_im2col_u8': (I, [P, P, ctypes.POINTER(ctypes.c_float * 3), ctypes.POINTER(ctypes.c_float * 3),
                                  I, I, I, ctypes.c_int64, P]),
        'mp_add_bf16': (I, [ctypes.POINTER(c_void_p * 4), I, P, ctypes.c_int64, ctypes.c_int64, P]),
        'mp_pack_weights': (I, [P, P, P, I, ctypes.c_int64, P]),
        'mp_sgd_step': (I, [P, P, P, ctypes.c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                            ctypes.c_float, I, I, ctypes.c_float, P]),
        'mp_sgd_step_hp': (I, [P, P, P, ctypes.c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                               ctypes.c_float, I, I, ctypes.c_float, P, P]),
    }


def lib():
    """Loads (once) and returns the C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MargiposeB200Error(
                'margipose_b200: %s not found. Build it with `python -m margipose_b200.build` '
                '(or __graft_entry__.build()); there is no CPU / PyTorch fallback.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _signatures().items():
            fn = getattr(handle, name)     # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_signatures().keys())


def check(rc, what=''):
    if rc != 0:
        msg = lib().mp_last_error().decode('utf-8', 'replace')
        raise MargiposeB200Error('%s failed (code %d): %s' % (what or 'margipose_b200 call', rc, msg))


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def planes(ts):
    """A 3-entry pointer table (host array of device pointers) for the plane-wise entry points."""
    if ts is None:
        return None
    return ctypes.byref(PlaneTable(*[(t.data_ptr() if t is not None else None) for t in ts]))


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MargiposeB200Error(
                'margipose_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path'
                % t.device)
