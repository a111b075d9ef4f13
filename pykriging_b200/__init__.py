"""pykriging_b200 -- B200-native implementation of pyKriging's model-fitting / prediction hot path.

Drop-in for ``pyKriging.matrixops`` (the mixin ``kriging`` and ``regression_kriging`` inherit from)
plus batched ``kriging`` / ``regression_kriging`` front ends with the reference's API.  All
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/pykriging_b200.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"
