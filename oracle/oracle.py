"""ctypes binding of the C oracle (oracle/ecfft_oracle.c).

ORACLE — TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package.

Field elements travel as numpy uint64 arrays of shape (n, 4): ark-ff memory layout,
little-endian limbs, Montgomery form (reference src/lib.rs:37).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libecfft_oracle.so")

P = 2**256 - 2**32 - 977
R = 2**256 % P
RINV = pow(R, -1, P)


def build(force=False):
    src = os.path.join(_HERE, "ecfft_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.orc_build_fftree.restype = vp
        L.orc_build_fftree.argtypes = [sz, ci]
        L.orc_tree_free.argtypes = [vp]
        L.orc_set_build_threads.argtypes = [ci]
        L.orc_set_build_threads.restype = None
        L.orc_tree_leaves.restype = sz
        L.orc_tree_leaves.argtypes = [vp]
        L.orc_subtree_with_size.restype = vp
        L.orc_subtree_with_size.argtypes = [vp, sz]
        L.orc_tree_table.restype = sz
        L.orc_tree_table.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
        for name in ("orc_enter", "orc_exit"):
            getattr(L, name).argtypes = [vp, vp, sz, vp]
        L.orc_enter_mt.argtypes = [vp, vp, sz, vp, ci]
        L.orc_enter_range.argtypes = [vp, vp, sz, sz, sz, vp]
        for name in ("orc_extend", "orc_mextend"):
            getattr(L, name).argtypes = [vp, vp, sz, ci, vp]
        L.orc_degree.argtypes = [vp, vp, sz, ctypes.POINTER(sz)]
        for name in ("orc_redc_z0", "orc_redc_z1"):
            getattr(L, name).argtypes = [vp, vp, vp, sz, vp]
        L.orc_modular_reduce.argtypes = [vp, vp, vp, vp, sz, vp]
        L.orc_vanish.argtypes = [vp, vp, sz, vp]
        L.orc_serialized_size.restype = sz
        L.orc_serialized_size.argtypes = [vp, ci]
        L.orc_serialize.restype = sz
        L.orc_serialize.argtypes = [vp, ci, vp, sz]
        L.orc_deserialize.restype = vp
        L.orc_deserialize.argtypes = [vp, sz, ci]
        L.orc_fe_mul.argtypes = [vp, vp, vp]
        L.orc_batch_inversion.argtypes = [vp, sz]
        _lib = L
    return _lib


# ---- conversions between Python ints (plain values) and Montgomery limb arrays ----
def to_mont(values):
    """list of ints (plain field values) -> (n,4) uint64 Montgomery limbs"""
    out = np.empty((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        m = (v % P) * R % P
        for k in range(4):
            out[i, k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_mont(arr):
    """(n,4) uint64 Montgomery limbs -> list of plain ints"""
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    res = []
    for row in arr.tolist():
        m = row[0] | (row[1] << 64) | (row[2] << 128) | (row[3] << 192)
        res.append(m * RINV % P)
    return res


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _in(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return a


class OracleError(RuntimeError):
    pass


class OracleTree:
    """FFTree<secp256k1::Fp> on the CPU oracle; method names follow reference src/fftree.rs."""

    def __init__(self, handle, owner=True):
        self._h = handle
        self._owner = owner

    @classmethod
    def build(cls, n, parts=0, threads=1):
        """threads > 1 only shortens the build (element-wise loops cut into ranges); the tables are identical"""
        lib().orc_set_build_threads(threads)
        try:
            h = lib().orc_build_fftree(n, parts)
        finally:
            lib().orc_set_build_threads(1)
        if not h:
            raise OracleError("build_fftree returned None")
        return cls(h)

    @classmethod
    def deserialize(cls, data, compressed):
        buf = np.frombuffer(data, dtype=np.uint8)
        h = lib().orc_deserialize(_ptr(buf), len(buf), 1 if compressed else 0)
        if not h:
            raise OracleError("deserialize failed")
        return cls(h)

    def __del__(self):
        if getattr(self, "_owner", False) and self._h and _lib is not None:
            _lib.orc_tree_free(self._h)
            self._h = None

    @property
    def leaves_count(self):
        return lib().orc_tree_leaves(self._h)

    def subtree_with_size(self, n):
        h = lib().orc_subtree_with_size(self._h, n)
        if not h:
            raise OracleError("FFTree is too small")
        t = OracleTree(h, owner=False)
        t._parent = self
        return t

    def table(self, name):
        p = ctypes.c_void_p()
        cnt = lib().orc_tree_table(self._h, name.encode(), ctypes.byref(p))
        if cnt == 0:
            return np.zeros((0, 4), dtype=np.uint64)
        buf = (ctypes.c_uint64 * (cnt * 4)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint64).reshape(cnt, 4).copy()

    def leaves(self):
        f = self.table("f")
        return f[len(f) // 2:]

    def _run(self, fn, *args):
        rc = fn(*args)
        if rc != 0:
            raise OracleError("oracle call failed (reference would panic)")

    def enter(self, coeffs, threads=1):
        x = _in(coeffs)
        out = np.empty_like(x)
        if threads > 1:
            self._run(lib().orc_enter_mt, self._h, _ptr(x), len(x), _ptr(out), threads)
        else:
            self._run(lib().orc_enter, self._h, _ptr(x), len(x), _ptr(out))
        return out

    def enter_range(self, data, m_lo, m_hi):
        x = _in(data)
        out = np.empty_like(x)
        self._run(lib().orc_enter_range, self._h, _ptr(x), len(x), m_lo, m_hi, _ptr(out))
        return out

    def exit(self, evals):
        x = _in(evals)
        out = np.empty_like(x)
        self._run(lib().orc_exit, self._h, _ptr(x), len(x), _ptr(out))
        return out

    def extend(self, evals, moiety):
        x = _in(evals)
        out = np.empty_like(x)
        self._run(lib().orc_extend, self._h, _ptr(x), len(x), int(moiety), _ptr(out))
        return out

    def mextend(self, evals, moiety):
        x = _in(evals)
        out = np.empty_like(x)
        self._run(lib().orc_mextend, self._h, _ptr(x), len(x), int(moiety), _ptr(out))
        return out

    def degree(self, evals):
        x = _in(evals)
        d = ctypes.c_size_t()
        self._run(lib().orc_degree, self._h, _ptr(x), len(x), ctypes.byref(d))
        return d.value

    def redc_z0(self, evals, a):
        x, y = _in(evals), _in(a)
        out = np.empty_like(x)
        self._run(lib().orc_redc_z0, self._h, _ptr(x), _ptr(y), len(x), _ptr(out))
        return out

    def redc_z1(self, evals, a):
        x, y = _in(evals), _in(a)
        out = np.empty_like(x)
        self._run(lib().orc_redc_z1, self._h, _ptr(x), _ptr(y), len(x), _ptr(out))
        return out

    def modular_reduce(self, evals, a, c):
        x, y, z = _in(evals), _in(a), _in(c)
        out = np.empty_like(x)
        self._run(lib().orc_modular_reduce, self._h, _ptr(x), _ptr(y), _ptr(z), len(x), _ptr(out))
        return out

    def vanish(self, domain):
        x = _in(domain)
        out = np.empty((2 * len(x), 4), dtype=np.uint64)
        self._run(lib().orc_vanish, self._h, _ptr(x), len(x), _ptr(out))
        return out

    def serialize(self, compressed):
        c = 1 if compressed else 0
        size = lib().orc_serialized_size(self._h, c)
        buf = np.empty(size, dtype=np.uint8)
        w = lib().orc_serialize(self._h, c, _ptr(buf), size)
        assert w == size
        return buf.tobytes()


def random_elements(n, seed=1):
    """n uniform field elements as raw Montgomery limbs (splitmix64 counter PRNG)."""
    base = (int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    idx = np.arange(4 * n, dtype=np.uint64) + np.uint64(base)
    with np.errstate(over="ignore"):
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    out = z.reshape(n, 4).copy()
    # reject >= p (probability 2^-224): clear the top limb's top bit in that case
    top = out[:, 3] == np.uint64(0xFFFFFFFFFFFFFFFF)
    out[top, 3] = np.uint64(0x7FFFFFFFFFFFFFFF)
    return out
