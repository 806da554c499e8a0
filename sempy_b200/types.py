"""Scalar type of the path.  The reference fixes it to 64-bit floats (sempy/types.py:3); every kernel
of libsemb200 computes in FP64, so the name is kept for callers that import it."""

from numpy import float64 as SEMPY_SCALAR

__all__ = ["SEMPY_SCALAR"]
