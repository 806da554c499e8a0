"""sempy_b200 -- B200-native drop-in for SemPy's Poisson hot path.

Same module and function names as the reference ``sempy`` package for the path
this repository covers (elliptic, gradient, kron, iterative, mesh and the 1-D
input modules); the work is done by hand-written CUDA kernels behind the C ABI
of ``include/semb200.h``.  There is no CPU fallback.
"""

__version__ = "0.1.0"
