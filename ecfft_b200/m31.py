"""FFTree<m31::Fp> on one B200: the Python mirror of the reference's method surface for its second field
(/root/reference/src/lib.rs:190-215, src/fftree.rs:41-497) over the C ABI (include/ecfft_b200.h, "m31").

Elements are numpy uint32 holding the canonical value in [0, 2^31 - 1) — the in-memory form of the reference's
`ark_ff_optimized::fp31::Fp(pub u32)`.  CUDA tensors (int32 / uint32 viewed as 4-byte elements) go through the
`_dev` entry points on torch's current stream.  No CPU fallback."""
import ctypes

import numpy as np

from . import _lib
from ._lib import EcfftError

P = (1 << 31) - 1


def _is_torch_cuda(x):
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _in(x):
    a = np.ascontiguousarray(x, dtype=np.uint32)
    if a.ndim != 1:
        raise EcfftError(_lib.ERR_INVALID_ARG, "expected a 1-D array of u32 field elements")
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class M31FFTree:
    def __init__(self, handle):
        self._h = handle
        self._L = _lib.load()

    @classmethod
    def build(cls, n, device=0):
        """<m31::Fp as FftreeField>::build_fftree(n), src/lib.rs:196-214; EcfftError(ERR_TOO_LARGE) where it returns None"""
        L = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(L.ecfft_m31_tree_build(n, device, ctypes.byref(h)))
        return cls(h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ecfft_m31_tree_free(self._h)
            self._h = None

    @property
    def leaves(self):
        return self._L.ecfft_m31_tree_leaves(self._h)

    def table(self, name, subtree_leaves=None):
        """a pub field of FFTree<F> (src/fftree.rs:23-38) of the subtree with `subtree_leaves` leaves"""
        n = subtree_leaves or self.leaves
        cnt = ctypes.c_size_t()
        _lib.check(self._L.ecfft_m31_tree_table(self._h, n, name.encode(), None, 0, ctypes.byref(cnt)))
        out = np.empty(cnt.value, dtype=np.uint32)
        _lib.check(self._L.ecfft_m31_tree_table(self._h, n, name.encode(), _p(out), out.size, ctypes.byref(cnt)))
        return out

    def eval_domain(self, n=None):
        n = n or self.leaves
        return self.table("f", n)[n:]

    def _host(self, fn, arrays, out_len, extra=()):
        arrays = [_in(a) for a in arrays]
        n = len(arrays[0])
        for a in arrays[1:]:
            if len(a) != n:
                raise EcfftError(_lib.ERR_INVALID_ARG, "operand lengths differ")
        out = np.empty(out_len, dtype=np.uint32)
        _lib.check(fn(self._h, *[_p(a) for a in arrays], n, *extra, _p(out)))
        return out

    def _dev(self, fn, x, extra=()):
        import torch
        if x.dtype not in (torch.int32, torch.uint32) or x.dim() != 1:
            raise EcfftError(_lib.ERR_INVALID_ARG, "expected a 1-D int32/uint32 CUDA tensor")
        x = x.contiguous()
        out = torch.empty_like(x)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(fn(self._h, ctypes.c_void_p(x.data_ptr()), x.shape[0], *extra, ctypes.c_void_p(out.data_ptr()), stream))
        return out

    def enter(self, coeffs):
        if _is_torch_cuda(coeffs):
            return self._dev(self._L.ecfft_m31_enter_dev, coeffs)
        return self._host(self._L.ecfft_m31_enter, [coeffs], len(coeffs))

    def exit(self, evals):
        if _is_torch_cuda(evals):
            return self._dev(self._L.ecfft_m31_exit_dev, evals)
        return self._host(self._L.ecfft_m31_exit, [evals], len(evals))

    def extend(self, evals, moiety):
        if _is_torch_cuda(evals):
            return self._dev(self._L.ecfft_m31_extend_dev, evals, (int(moiety),))
        return self._host(self._L.ecfft_m31_extend, [evals], len(evals), (int(moiety),))

    def mextend(self, evals, moiety):
        return self._host(self._L.ecfft_m31_mextend, [evals], len(evals), (int(moiety),))

    def degree(self, evals):
        a = _in(evals)
        d = ctypes.c_size_t()
        _lib.check(self._L.ecfft_m31_degree(self._h, _p(a), len(a), ctypes.byref(d)))
        return d.value

    def redc_z0(self, evals, a):
        return self._host(self._L.ecfft_m31_redc_z0, [evals, a], len(evals))

    def redc_z1(self, evals, a):
        return self._host(self._L.ecfft_m31_redc_z1, [evals, a], len(evals))

    def modular_reduce(self, evals, a, c):
        return self._host(self._L.ecfft_m31_modular_reduce, [evals, a, c], len(evals))

    def vanish(self, domain):
        return self._host(self._L.ecfft_m31_vanish, [domain], 2 * len(domain))


def build_fftree(n, device=0):
    return M31FFTree.build(n, device)
