"""Host-side conv geometry (margipose_b200/convops.py) against torch's CPU convolution, with the
CUDA launches replaced by the CPU emulation of their contracts (tests/emulate.py).  Runs without
a GPU: it proves the tap tables / views / packs the kernels are given are the right ones."""
import pytest

from tests import conv_cases as K
from tests import emulate

IDS = lambda c: 'x'.join(map(str, c))   # noqa: E731
SMALL = [c for c in K.CASES if c[5] * c[6] * c[7] * c[0] * c[1] <= 2 * 32 * 32 * 128 * 192]


@pytest.fixture(autouse=True)
def _cpu(monkeypatch):
    emulate.install(monkeypatch)
    K.DEV = 'cpu'
    yield
    K.DEV = 'cuda'


@pytest.mark.parametrize('case', SMALL, ids=IDS)
def test_conv_forward_and_stats(case):
    K.check_conv_forward_and_stats(case)


@pytest.mark.parametrize('case', SMALL, ids=IDS)
def test_conv_dgrad_and_wgrad(case):
    K.check_conv_dgrad_and_wgrad(case)


@pytest.mark.parametrize('stride,tr', [(1, False), (2, False), (2, True)])
def test_fused_block_dgrad(stride, tr):
    K.check_fused_block_dgrad(stride, tr)


@pytest.mark.parametrize('case', [c for c in SMALL if c[0] >= 64][:8], ids=IDS)
def test_split_precision_conv(case):
    """bf16x3 mode: the three-pass expansion (hi*hi, lo*hi, hi*lo chained through acc_in) in convops."""
    K.check_split_conv(case)
