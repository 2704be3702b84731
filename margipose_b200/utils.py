"""Process-wide algorithm switches, mirror of /root/reference/src/margipose/utils.py:12-24.

The reference's `init_algorithms(deterministic)` selects deterministic cuDNN algorithms; here it selects the
engine's deterministic mode for models created afterwards: no floating-point atomics between thread blocks
(fixed-order BatchNorm statistics and backward reductions, weight gradients without split-K, ordered combiner
gradient), i.e. bit-wise reproducible training steps at some cost in speed.
"""
import random

import torch

DETERMINISTIC = False


def seed_all(seed):
    """Seed all random number generators (utils.py:12-16)."""
    random.seed(seed)
    try:
        import numpy as np
        np.random.seed(seed)
    except ImportError:
        pass
    torch.manual_seed(seed)


def init_algorithms(deterministic=False):
    """utils.py:19-24."""
    global DETERMINISTIC
    DETERMINISTIC = bool(deterministic)
