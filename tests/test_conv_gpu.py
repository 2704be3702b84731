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


@pytest.mark.parametrize('cluster', [2, 4])
def test_conv_forward_with_multicast_clusters(cluster):
    """The weight tile can be fetched once per thread-block cluster and TMA-multicast to its CTAs
    (tunable igemm_cluster); results must not depend on it."""
    from margipose_b200._lib import lib
    K.DEV = 'cuda'
    try:
        assert lib().mp_set_tunable(b'igemm_cluster', cluster) == 0
        K.check_conv_forward_and_stats(K.CASES[0])
        K.check_conv_dgrad_and_wgrad(K.CASES[1])
    finally:
        lib().mp_set_tunable(b'igemm_cluster', 1)
