"""Host-side mirror of the reference's `FFTree<secp256k1::Fp>` (reference src/fftree.rs:23-497,
src/lib.rs:14-16,39-85) over the C ABI of the CUDA engine.

Same method names, argument meaning and error behaviour as the Rust type: where the reference
panics (`assert!(n.is_power_of_two())`, "FFTree is too small") these raise `EcfftError`;
`build_fftree` returns None when log2 n >= 36 (src/lib.rs:61-64).

Vectors of field elements are numpy uint64 arrays of shape (n, 4): 4 little-endian limbs per
element in Montgomery form — byte for byte the reference's `&[Fp]`.  Methods also accept CUDA
torch tensors of dtype int64/uint64 and shape (n, 4); those run on the tensor's device without
host copies and return a new tensor (enqueued on torch's current stream).
"""
import ctypes
import enum

import numpy as np

from . import _lib
from ._lib import EcfftError

PARTS_FULL = 0
PARTS_ENTER_ONLY = 1


class Moiety(enum.IntEnum):
    """reference src/fftree.rs:17-21"""
    S0 = 0
    S1 = 1


def _is_torch_cuda(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


def _np_in(x):
    a = np.ascontiguousarray(x, dtype=np.uint64)
    if a.ndim == 1:
        if a.size % 4:
            raise EcfftError(_lib.ERR_INVALID_ARG, "flat limb array length must be a multiple of 4")
        a = a.reshape(-1, 4)
    if a.ndim != 2 or a.shape[1] != 4:
        raise EcfftError(_lib.ERR_INVALID_ARG, "expected an (n, 4) array of u64 limbs")
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class FFTree:
    """FFTree<secp256k1::Fp> resident on one B200 (tables stay in HBM for the handle's life)."""

    def __init__(self, handle):
        self._h = handle
        self._L = _lib.load()

    # ---- construction / persistence -----------------------------------------------------
    @classmethod
    def build(cls, n, parts=PARTS_FULL, device=0):
        L = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(L.ecfft_tree_build_secp256k1(n, parts, device, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def new(cls, leaves, rational_maps, parts=PARTS_FULL, device=0):
        """FFTree::new(leaves, rational_maps), src/fftree.rs:42-70.
        rational_maps: list of (numerator_coeffs, denominator_coeffs), each an (k,4) limb array."""
        L = _lib.load()
        lv = _np_in(leaves)
        lens, coeffs = [], []
        for num, den in rational_maps:
            num, den = _np_in(num), _np_in(den)
            lens += [len(num), len(den)]
            coeffs += [num, den]
        cat = np.concatenate(coeffs) if coeffs else np.zeros((0, 4), dtype=np.uint64)
        cat = np.ascontiguousarray(cat)
        lens_arr = (ctypes.c_size_t * max(len(lens), 1))(*lens)
        h = ctypes.c_void_p()
        _lib.check(L.ecfft_tree_new(_p(lv), len(lv), _p(cat), lens_arr, len(rational_maps), parts, device, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def deserialize(cls, data, compressed, device=0):
        """CanonicalDeserialize, src/fftree.rs:602-660"""
        L = _lib.load()
        buf = np.frombuffer(data, dtype=np.uint8)
        h = ctypes.c_void_p()
        _lib.check(L.ecfft_tree_deserialize(_p(buf), len(buf), 1 if compressed else 0, device, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def deserialize_compressed(cls, data, device=0):
        return cls.deserialize(data, True, device)

    @classmethod
    def deserialize_uncompressed(cls, data, device=0):
        return cls.deserialize(data, False, device)

    def serialized_size(self, compressed):
        s = ctypes.c_size_t()
        _lib.check(self._L.ecfft_tree_serialized_size(self._h, 1 if compressed else 0, ctypes.byref(s)))
        return s.value

    def serialize(self, compressed):
        """CanonicalSerialize, src/fftree.rs:510-554"""
        size = self.serialized_size(compressed)
        buf = np.empty(size, dtype=np.uint8)
        w = ctypes.c_size_t()
        _lib.check(self._L.ecfft_tree_serialize(self._h, 1 if compressed else 0, _p(buf), size, ctypes.byref(w)))
        return buf[: w.value].tobytes()

    def serialize_compressed(self):
        return self.serialize(True)

    def serialize_uncompressed(self):
        return self.serialize(False)

    def close(self):
        if self._h:
            self._L.ecfft_tree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- structure ----------------------------------------------------------------------
    @property
    def leaves_count(self):
        return self._L.ecfft_tree_leaves(self._h)

    @property
    def device(self):
        return self._L.ecfft_tree_device(self._h)

    def subtree_with_size(self, n):
        """src/fftree.rs:489-496 — a view; tables are shared with the parent handle."""
        return _SubtreeView(self, n)

    def table(self, name, subtree_leaves=None):
        """One of the pub fields of src/fftree.rs:25-37 as Montgomery limbs."""
        n = self.leaves_count if subtree_leaves is None else subtree_leaves
        cnt = ctypes.c_size_t()
        _lib.check(self._L.ecfft_tree_table(self._h, n, name.encode(), None, 0, ctypes.byref(cnt)))
        out = np.empty((cnt.value, 4), dtype=np.uint64)
        _lib.check(self._L.ecfft_tree_table(self._h, n, name.encode(), _p(out), cnt.value, ctypes.byref(cnt)))
        return out

    def eval_domain(self, subtree_leaves=None):
        """f.leaves() (src/fftree.rs:502-504)"""
        f = self.table("f", subtree_leaves)
        return f[len(f) // 2:]

    # ---- dispatch helpers ---------------------------------------------------------------
    def _dev_call(self, fn, tensors, out_rows, extra=()):
        import torch
        x0 = tensors[0]
        for x in tensors:
            if not _is_torch_cuda(x) or x.device.index != self.device or x.dtype not in (torch.int64, torch.uint64) or x.dim() != 2 or x.shape[1] != 4:
                raise EcfftError(_lib.ERR_INVALID_ARG, "expected (n,4) int64/uint64 CUDA tensors on the tree's device")
            if x.shape[0] != x0.shape[0]:   # the C entry points take ONE n: a shorter operand would be read out of bounds
                raise EcfftError(_lib.ERR_INVALID_ARG, "operand lengths differ")
        tensors = [x.contiguous() for x in tensors]
        out = torch.empty((out_rows, 4), dtype=x0.dtype, device=x0.device)
        stream = torch.cuda.current_stream(x0.device).cuda_stream
        args = [self._h] + [ctypes.c_void_p(x.data_ptr()) for x in tensors] + [x0.shape[0]] + list(extra) + [ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)]
        _lib.check(fn(*args))
        return out

    def _host_call(self, fn, arrays, out_rows, extra=()):
        arrays = [_np_in(a) for a in arrays]
        n = len(arrays[0])
        for a in arrays[1:]:
            if len(a) != n:
                raise EcfftError(_lib.ERR_INVALID_ARG, "operand lengths differ")
        out = np.empty((out_rows, 4), dtype=np.uint64)
        _lib.check(fn(self._h, *[_p(a) for a in arrays], n, *extra, _p(out)))
        return out

    def _call(self, name, arrays, out_rows, extra=()):
        if _is_torch_cuda(arrays[0]):
            return self._dev_call(getattr(self._L, name + "_dev"), list(arrays), out_rows, extra)
        return self._host_call(getattr(self._L, name), arrays, out_rows, extra)

    # ---- the eight algorithms (src/fftree.rs:123-316) -------------------------------------
    def enter(self, coeffs):
        """coefficients (low -> high) -> evaluations at the leaves"""
        return self._call("ecfft_enter", [coeffs], len(coeffs))

    def enter_many(self, coeffs, out=None):
        """`count` coefficient vectors, shape (count, n, 4) host array -> (count, n, 4) evaluations, as one pipelined
        call (uploads and downloads overlap the kernels of the neighbouring vectors).  `out`: optional destination
        (e.g. page-locked memory)."""
        a = np.ascontiguousarray(coeffs, dtype=np.uint64)
        if a.ndim != 3 or a.shape[2] != 4:
            raise EcfftError(_lib.ERR_INVALID_ARG, "expected a (count, n, 4) array of u64 limbs")
        if out is None:
            out = np.empty_like(a)
        elif out.shape != a.shape or out.dtype != np.uint64 or not out.flags.c_contiguous:
            raise EcfftError(_lib.ERR_INVALID_ARG, "out must be a C-contiguous uint64 array of the input's shape")
        _lib.check(self._L.ecfft_enter_many(self._h, _p(a), a.shape[1], a.shape[0], _p(out)))
        return out

    def exit(self, evals):
        """evaluations -> coefficients"""
        return self._call("ecfft_exit", [evals], len(evals))

    def extend(self, evals, moiety):
        """evaluations on one moiety -> evaluations on `moiety` (the target)"""
        return self._call("ecfft_extend", [evals], len(evals), (int(moiety),))

    def mextend(self, evals, moiety):
        return self._call("ecfft_mextend", [evals], len(evals), (int(moiety),))

    def degree(self, evals):
        d = ctypes.c_size_t()
        if _is_torch_cuda(evals):
            import torch
            x = evals
            if x.device.index != self.device or x.dtype not in (torch.int64, torch.uint64) or x.dim() != 2 or x.shape[1] != 4:
                raise EcfftError(_lib.ERR_INVALID_ARG, "expected an (n,4) int64/uint64 CUDA tensor on the tree's device")
            x = x.contiguous()
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(self._L.ecfft_degree_dev(self._h, ctypes.c_void_p(x.data_ptr()), x.shape[0], ctypes.byref(d), ctypes.c_void_p(stream)))
        else:
            a = _np_in(evals)
            _lib.check(self._L.ecfft_degree(self._h, _p(a), len(a), ctypes.byref(d)))
        return d.value

    def redc_z0(self, evals, a):
        return self._call("ecfft_redc_z0", [evals, a], len(evals))

    def redc_z1(self, evals, a):
        return self._call("ecfft_redc_z1", [evals, a], len(evals))

    def modular_reduce(self, evals, a, c):
        return self._call("ecfft_modular_reduce", [evals, a, c], len(evals))

    def vanish(self, vanish_domain):
        return self._call("ecfft_vanish", [vanish_domain], 2 * len(vanish_domain))

    def pointwise_mul(self, a, b):
        """a[i] * b[i] on the tree's GPU (Montgomery limbs in and out, like `*` on two ark-ff values): the step
        between `enter` and `exit` when polynomials are multiplied through the tree"""
        return self._call("ecfft_pointwise_mul", [a, b], len(a))

    def enter_range(self, data, m_lo, m_hi):
        """device-only building block of the multi-GPU ENTER (include/ecfft_b200.h)"""
        return self._dev_call(self._L.ecfft_enter_range_dev, [data], data.shape[0], (m_lo, m_hi))


    # ---- multi-GPU building blocks (device tensors only; include/ecfft_b200.h "fully sharded") ----
    def _mg_out(self, like, rows):
        import torch
        return torch.empty((rows, 4), dtype=like.dtype, device=like.device)

    def _stream(self, x):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)

    def mg_prescale(self, m, pos0, x):
        x = x.contiguous()
        out = self._mg_out(x, x.shape[0])
        _lib.check(self._L.ecfft_mg_prescale_dev(self._h, m, pos0, ctypes.c_void_p(x.data_ptr()), x.shape[0],
                                                 ctypes.c_void_p(out.data_ptr()), self._stream(x)))
        return out

    def mg_cross(self, m, phase, j, role, p_pos0, own, partner):
        own, partner = own.contiguous(), partner.contiguous()
        out = self._mg_out(own, own.shape[0])
        _lib.check(self._L.ecfft_mg_cross_dev(self._h, m, phase, j, role, p_pos0, ctypes.c_void_p(own.data_ptr()),
                                              ctypes.c_void_p(partner.data_ptr()), own.shape[0],
                                              ctypes.c_void_p(out.data_ptr()), self._stream(own)))
        return out

    def mg_local(self, m, x):
        x = x.contiguous()
        out = self._mg_out(x, x.shape[0])
        _lib.check(self._L.ecfft_mg_local_dev(self._h, m, ctypes.c_void_p(x.data_ptr()), x.shape[0],
                                              ctypes.c_void_p(out.data_ptr()), self._stream(x)))
        return out

    def mg_combine(self, m, i0, u0, v0, u1, v1):
        u0, v0, u1, v1 = (t.contiguous() for t in (u0, v0, u1, v1))
        out = self._mg_out(u0, 2 * u0.shape[0])
        _lib.check(self._L.ecfft_mg_combine_dev(self._h, m, i0, *[ctypes.c_void_p(t.data_ptr()) for t in (u0, v0, u1, v1)],
                                                u0.shape[0], ctypes.c_void_p(out.data_ptr()), self._stream(u0)))
        return out


class _SubtreeView:
    """`&FFTree` returned by subtree_with_size: same surface, restricted to n leaves."""

    def __init__(self, parent, n):
        if n <= 0 or n & (n - 1):
            raise EcfftError(_lib.ERR_NOT_POW2, "n is not a power of two")
        if n > parent.leaves_count:
            raise EcfftError(_lib.ERR_TREE_TOO_SMALL, "FFTree is too small")
        self._t = parent
        self.leaves_count = n

    def _guard(self, need):
        if need > self.leaves_count:
            raise EcfftError(_lib.ERR_TREE_TOO_SMALL, "FFTree is too small")

    def subtree_with_size(self, n):
        self._guard(n)
        return _SubtreeView(self._t, n)

    def table(self, name):
        return self._t.table(name, self.leaves_count)

    def eval_domain(self):
        return self._t.eval_domain(self.leaves_count)

    def enter(self, x):
        self._guard(len(x)); return self._t.enter(x)

    def exit(self, x):
        self._guard(len(x)); return self._t.exit(x)

    def extend(self, x, moiety):
        self._guard(2 * len(x)); return self._t.extend(x, moiety)

    def mextend(self, x, moiety):
        self._guard(2 * len(x)); return self._t.mextend(x, moiety)

    def degree(self, x):
        self._guard(len(x)); return self._t.degree(x)

    def redc_z0(self, x, a):
        self._guard(len(x)); return self._t.redc_z0(x, a)

    def redc_z1(self, x, a):
        self._guard(len(x)); return self._t.redc_z1(x, a)

    def modular_reduce(self, x, a, c):
        self._guard(len(x)); return self._t.modular_reduce(x, a, c)

    def vanish(self, x):
        self._guard(2 * len(x)); return self._t.vanish(x)


def build_fftree(n, parts=PARTS_FULL, device=0):
    """<secp256k1::Fp as FftreeField>::build_fftree(n), src/lib.rs:39-85: None when log2 n >= 36."""
    if n <= 0 or n & (n - 1):
        raise EcfftError(_lib.ERR_NOT_POW2, "n is not a power of two")
    try:
        return FFTree.build(n, parts, device)
    except EcfftError as e:
        if e.code == _lib.ERR_TOO_LARGE:
            return None
        raise
