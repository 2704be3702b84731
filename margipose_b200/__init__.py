"""margipose_b200 -- B200-native (sm_100a) implementation of the MargiPose hot path.

Mirrors the reference's plugin surface for that path and nothing else:
  margipose_b200.models.create_model(model_desc)   <- margipose.models.create_model
  margipose_b200.dsntnn.*                          <- margipose.dsntnn.*
Compute runs in hand-written CUDA kernels behind the C ABI in include/margipose_b200.h
(libmargipose_b200.so, built in-tree by margipose_b200.build).  No CPU fallback.
"""
__version__ = '0.1.0'
