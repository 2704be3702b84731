"""GPU parity of the tcgen05 implicit-GEMM convolution kernels through the C ABI (see
tests/conv_cases.py for the checks and tolerances)."""
import pytest

from tests import conv_cases as K

pytestmark = pytest.mark.gpu
IDS = lambda c: 'x'.join(map(str, c))   # noqa: E731


@pytest.mark.parametrize('case', K.CASES, ids=IDS)
def test_conv_forward_and_stats(case):
    K.DEV = 'cuda'
    K.check_conv_forward_and_stats(case)


@pytest.mark.parametrize('case', K.CASES, ids=IDS)
def test_conv_dgrad_and_wgrad(case):
    K.DEV = 'cuda'
    K.check_conv_dgrad_and_wgrad(case)


@pytest.mark.parametrize('stride,tr', [(1, False), (2, False), (2, True)])
def test_fused_block_dgrad(stride, tr):
    K.DEV = 'cuda'
    K.check_fused_block_dgrad(stride, tr)


DEFAULTS = {'igemm_halo': 1, 'wgrad_halo': 0, 'igemm_pair': 1, 'igemm_resident': 1, 'wgrad_kp': 128, 'igemm_mt': 2,
            'igemm_mt_ctas': 0, 'wgrad_slice': 256, 'wgrad_ctas': 148}


@pytest.mark.parametrize('tunables', [
    {'igemm_halo': 0, 'wgrad_halo': 1}, {'igemm_pair': 0}, {'igemm_resident': 0}, {'igemm_pair': 0, 'igemm_resident': 0},
    {'igemm_mt_ctas': 1}, {'igemm_mt_ctas': 1, 'igemm_pair': 0}, {'igemm_mt': 1}, {'igemm_mt_ctas': 1, 'igemm_resident': 0},
    {'wgrad_slice': 64, 'wgrad_ctas': 40, 'wgrad_kp': 64},
], ids=lambda d: ','.join('%s=%d' % kv for kv in d.items()))
def test_conv_kernel_variants(tunables):
    """Row-shifted taps may share one halo box (igemm_halo / wgrad_halo), a CTA may own two 128-pixel
    accumulators (igemm_mt*), two CTAs may pair up on M=256 tcgen05.mma.cta_group::2 tiles with half a
    weight tile each (igemm_pair), wgrad may slice its columns / stages differently: results must not
    depend on any of it."""
    from margipose_b200._lib import lib
    K.DEV = 'cuda'
    try:
        for k, v in tunables.items():
            assert lib().mp_set_tunable(k.encode(), v) == 0
        for case in K.CASES[:2] + K.CASES[7:8] + K.CASES[10:11]:
            K.check_conv_forward_and_stats(case)
            K.check_conv_dgrad_and_wgrad(case)
        K.check_fused_block_dgrad(1, False)
    finally:
        for k in tunables:
            lib().mp_set_tunable(k.encode(), DEFAULTS[k])


@pytest.mark.parametrize('case', K.CASES[:3] + K.CASES[7:9] + K.CASES[10:12] + K.CASES[16:17], ids=IDS)
def test_split_precision_conv(case):
    """bf16x3 mode on the tensor cores: every operand a bf16 pair, three passes per accumulation chained through
    the epilogue's acc_in -- fp32-grade results (~1e-5) from bf16 MMAs."""
    from tests.conftest import parity_log
    K.DEV = 'cuda'
    err = K.check_split_conv(case)
    parity_log('conv_bf16x3/' + IDS(case), forward_max_err_over_max=err, tolerance='rtol 2e-5 vs the fp64 convolution')
