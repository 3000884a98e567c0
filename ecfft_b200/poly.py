"""Caller-side polynomial arithmetic on top of the FFTree surface: what a user of the
reference writes around `enter` / `exit` (reference README.md:60-63; the evaluate / interpolate pair timed by
benches/comparison.rs:37-43).  Everything stays on the tree's GPU when the inputs are CUDA tensors.

Coefficient vectors are (k, 4) arrays of u64 limbs, low -> high, Montgomery form — the reference's `&[Fp]`.
"""
import numpy as np

from . import _lib
from ._lib import EcfftError


def _next_pow2(k):
    n = 1
    while n < k:
        n *= 2
    return n


def _pad(x, n):
    """zero-extend a coefficient vector to n elements (numpy array or CUDA tensor)"""
    if len(x) == n:
        return x
    if isinstance(x, np.ndarray):
        out = np.zeros((n, 4), dtype=np.uint64)
        out[: len(x)] = x
        return out
    import torch
    out = torch.zeros((n, 4), dtype=x.dtype, device=x.device)
    out[: len(x)] = x
    return out


def poly_mul(tree, a, b):
    """Product of two polynomials over secp256k1's base field: ENTER both on the smallest subtree that holds
    deg a + deg b + 1 coefficients, multiply the evaluations element-wise, EXIT.  Returns len(a) + len(b) - 1
    coefficients.  Raises EcfftError (tree too small) like `subtree_with_size` panics in the reference."""
    la, lb = len(a), len(b)
    if la == 0 or lb == 0:
        raise EcfftError(_lib.ERR_INVALID_ARG, "empty polynomial")
    n = _next_pow2(la + lb - 1)
    if n > tree.leaves_count:
        raise EcfftError(_lib.ERR_TREE_TOO_SMALL, "FFTree is too small")
    ea = tree.enter(_pad(a, n))
    eb = tree.enter(_pad(b, n))
    return tree.exit(tree.pointwise_mul(ea, eb))[: la + lb - 1]
