"""Error contract of the C ABI (include/margipose_b200.h "Conventions"): every compute entry point validates its
arguments BEFORE touching CUDA, returns MP_ERR_ARG (-1) instead of throwing or crashing, and leaves a message that
names the call in mp_last_error().  No kernel is launched here (runs on a CPU-only box); the parity tests proper
are the `-m gpu` ones."""
import ctypes

import pytest

from margipose_b200 import _lib as L

MP_ERR_ARG = -1
FAKE = 0x1000        # stands for a device pointer; argument checks never dereference device pointers on the host


def _tables():
    null3 = ctypes.byref(L.PlaneTable(None, None, None))
    some3 = ctypes.byref(L.PlaneTable(FAKE, None, None))
    return null3, some3


def _cases(h):
    null3, some3 = _tables()
    N = None
    return [
        ('mp_tail_fwd', 'no input plane', lambda: h.mp_tail_fwd(null3, 1, null3, N, N, N, N, N, N, N, 0, 1, 1.0, 2, 17, 32, 32, N)),
        ('mp_tail_fwd', 'empty batch', lambda: h.mp_tail_fwd(some3, 1, null3, N, N, N, N, N, N, N, 0, 1, 1.0, 0, 17, 32, 32, N)),
        ('mp_tail_fwd', 'plane larger than 16384 elements', lambda: h.mp_tail_fwd(some3, 1, null3, N, N, N, N, N, N, N, 0, 1, 1.0, 2, 17, 256, 256, N)),
        ('mp_tail_fwd', 'loss without targets', lambda: h.mp_tail_fwd(some3, 1, null3, N, N, N, N, N, N, FAKE, 0, 1, 1.0, 2, 17, 32, 32, N)),
        ('mp_tail_bwd', 'null plane tables', lambda: h.mp_tail_bwd(N, N, N, N, N, N, N, N, N, 1, 1, 1.0, 2, 17, 32, 32, N)),
        ('mp_tail_bwd', 'output plane without probabilities', lambda: h.mp_tail_bwd(null3, N, some3, N, N, N, N, N, N, 1, 1, 1.0, 2, 17, 32, 32, N)),
        ('mp_tail_bwd', 'fused mode without coords / w', lambda: h.mp_tail_bwd(some3, N, some3, FAKE, N, N, N, N, N, 1, 1, 1.0, 2, 17, 32, 32, N)),
        ('mp_masked_mean_fwd', 'null', lambda: h.mp_masked_mean_fwd(N, N, 4, N, N)),
        ('mp_masked_mean_bwd', 'null', lambda: h.mp_masked_mean_bwd(N, N, N, 4, N, N)),
        ('mp_euclid_fwd', 'null', lambda: h.mp_euclid_fwd(N, N, 4, 3, N, N)),
        ('mp_euclid_fwd', 'zero dimensions', lambda: h.mp_euclid_fwd(FAKE, FAKE, 4, 0, FAKE, N)),
        ('mp_euclid_bwd', 'null', lambda: h.mp_euclid_bwd(N, N, N, N, 4, 3, N, N)),
        ('mp_make_gauss', 'null', lambda: h.mp_make_gauss(N, N, 1, 1.0, 4, 32, 32, N)),
        ('mp_conv_igemm', 'null', lambda: h.mp_conv_igemm(N, N)),
        ('mp_conv_igemm_grouped', 'no problems', lambda: h.mp_conv_igemm_grouped(N, 0, N)),
        ('mp_conv_wgrad', 'null', lambda: h.mp_conv_wgrad(N, N)),
        ('mp_conv_wgrad_grouped', 'no problems', lambda: h.mp_conv_wgrad_grouped(N, 0, N)),
        ('mp_set_tunable', 'unknown name', lambda: h.mp_set_tunable(b'no_such_tunable', 1)),
        ('mp_set_tunable', 'null name', lambda: h.mp_set_tunable(N, 1)),
        ('mp_bn_fold_eval', 'null', lambda: h.mp_bn_fold_eval(N, 3, N)),
        ('mp_bn_stats', 'null', lambda: h.mp_bn_stats(N, N)),
        ('mp_bn_stats_grouped', 'no problems', lambda: h.mp_bn_stats_grouped(N, 0, N)),
        ('mp_bn_fwd', 'null', lambda: h.mp_bn_fwd(N, N)),
        ('mp_bn_fwd_grouped', 'no problems', lambda: h.mp_bn_fwd_grouped(N, 0, N)),
        ('mp_bn_bwd_reduce', 'null', lambda: h.mp_bn_bwd_reduce(N, N)),
        ('mp_bn_bwd_reduce_grouped', 'no problems', lambda: h.mp_bn_bwd_reduce_grouped(N, 0, N)),
        ('mp_bn_bwd_apply', 'null', lambda: h.mp_bn_bwd_apply(N, N)),
        ('mp_bn_bwd_apply_grouped', 'no problems', lambda: h.mp_bn_bwd_apply_grouped(N, 0, N)),
        ('mp_maxpool_fwd', 'null', lambda: h.mp_maxpool_fwd(N, N, N, 2, 64, 32, 32, 0, N)),
        ('mp_maxpool_bwd', 'null', lambda: h.mp_maxpool_bwd(N, N, N, 2, 64, 32, 32, N)),
        ('mp_axis_permute', 'null', lambda: h.mp_axis_permute(N, N, 2, 64, 32, 32, 0, N)),
        ('mp_combiner_fwd', 'null', lambda: h.mp_combiner_fwd(N, N, N, N, 2, 17, 32, 32, 0, N)),
        ('mp_combiner_bwd', 'null', lambda: h.mp_combiner_bwd(N, N, N, N, N, 2, 17, 32, 32, 0, 0, N)),
        ('mp_stem_im2col', 'null', lambda: h.mp_stem_im2col(N, N, 2, 256, 256, 0, N)),
        ('mp_stem_im2col_u8', 'null', lambda: h.mp_stem_im2col_u8(N, N, N, N, 2, 256, 256, 0, N)),
        ('mp_add_bf16', 'null', lambda: h.mp_add_bf16(N, 2, N, 0, 0, N)),
        ('mp_pack_weights', 'null', lambda: h.mp_pack_weights(N, N, N, 1, 1, N)),
        ('mp_sgd_step', 'null', lambda: h.mp_sgd_step(N, N, N, 8, 0.1, 0.9, 0.0, 0.0, 0, 0, 1.0, N)),
        ('mp_sgd_step_hp', 'null', lambda: h.mp_sgd_step_hp(N, N, N, 8, 0.1, 0.9, 0.0, 0.0, 0, 0, 1.0, N, N)),
    ]


def test_every_entry_point_rejects_bad_arguments_without_a_gpu():
    h = L.lib()
    seen = set()
    for name, what, call in _cases(h):
        rc = call()
        msg = h.mp_last_error().decode('utf-8', 'replace')
        assert rc == MP_ERR_ARG, '%s (%s): expected MP_ERR_ARG, got %d' % (name, what, rc)
        # grouped / _hp variants report under their family's name
        family = name.replace('_grouped', '').replace('_hp', '')
        assert family in msg, '%s (%s): message does not name the call: %r' % (name, what, msg)
        seen.add(name)
    missing = set(L.exported_symbols()) - seen - {'mp_abi_version', 'mp_last_error'}
    assert not missing, 'entry points without an argument-check case: %s' % sorted(missing)


def test_failed_calls_surface_as_python_exceptions():
    """The Python mirror never ignores a return code: `check` raises with the library's message (the reference's
    error behaviour is Python exceptions)."""
    h = L.lib()
    rc = h.mp_set_tunable(b'no_such_tunable', 1)
    with pytest.raises(L.MargiposeB200Error, match='no_such_tunable'):
        L.check(rc, 'mp_set_tunable')
    L.check(0)      # success is silent


def test_last_error_is_per_thread():
    """mp_last_error() is thread-local: a failure on one thread does not overwrite another thread's message."""
    import threading
    h = L.lib()
    assert h.mp_set_tunable(b'first_thread_name', 1) == MP_ERR_ARG
    seen = {}

    def other():
        assert h.mp_set_tunable(b'second_thread_name', 1) == MP_ERR_ARG
        seen['other'] = h.mp_last_error().decode()

    t = threading.Thread(target=other)
    t.start()
    t.join()
    assert 'second_thread_name' in seen['other']
    assert 'first_thread_name' in h.mp_last_error().decode()
