"""The device field arithmetic (ecfft_b200/csrc/fp.cuh) compiled for the HOST — the carry-flag PTX
primitives are emulated, everything above them is the exact code the kernels run — fuzzed against
Python big integers.  Catches limb-level logic errors without a GPU."""
import ctypes
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**256 - 2**32 - 977
A8 = ctypes.c_uint32 * 8


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(ROOT, "tools", "_fp_host.so")
    src = os.path.join(ROOT, "tools", "fp_host_shim.cpp")
    hdr = os.path.join(ROOT, "ecfft_b200", "csrc", "fp.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-x", "c++", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def enc(x):
    return A8(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def dec(a):
    return sum(int(a[i]) << (32 * i) for i in range(8))


EDGE = [0, 1, 2, P - 1, P - 2, P, P + 1, 2**256 - 1, 2**256 - 2, 2**255, 977, 2**32 + 977, 2**32, 2**224, 2**256 - 2**32]


def rnd(r):
    c = r.random()
    if c < 0.3:
        return r.choice(EDGE)
    if c < 0.5:
        return 2**256 - 1 - r.getrandbits(r.choice([8, 40, 70]))
    if c < 0.6:
        return r.getrandbits(r.choice([8, 40, 70]))
    return r.getrandbits(256)


def test_mul_reduce_and_lazy_accumulation(shim):
    r = random.Random(1)
    out = A8()
    for _ in range(20000):
        a, b, c, d = rnd(r), rnd(r), rnd(r), rnd(r)   # any 256-bit pattern, not only canonical values
        shim.fph_mul_lazy(enc(a), enc(b), out)
        assert dec(out) % P == a * b % P
        shim.fph_mul(enc(a), enc(b), out)
        assert dec(out) == a * b % P
        shim.fph_dot2(enc(a), enc(b), enc(c), enc(d), out)   # one reduction for a*b + c*d (the butterfly row)
        assert dec(out) % P == (a * b + c * d) % P
        shim.fph_muladd(enc(c), enc(a), enc(b), out)
        assert dec(out) % P == (a * b + c) % P
        shim.fph_canon(enc(a), out)
        assert dec(out) == a % P


def test_add_sub_mont(shim):
    r = random.Random(2)
    out = A8()
    rinv = pow(2**256, -1, P)
    for _ in range(20000):
        a, b = rnd(r) % P, rnd(r) % P
        shim.fph_add(enc(a), enc(b), out)
        assert dec(out) == (a + b) % P
        shim.fph_sub(enc(a), enc(b), out)
        assert dec(out) == (a - b) % P
        l1, l2 = rnd(r), rnd(r)                      # lazy + lazy (any 256-bit patterns)
        shim.fph_add_lazy(enc(l1), enc(l2), out)
        assert dec(out) % P == (l1 + l2) % P
        lazy = rnd(r)                                # any 256-bit pattern minus a canonical value
        shim.fph_sub_lazy(enc(lazy), enc(b), out)
        assert dec(out) % P == (lazy - b) % P
        shim.fph_sub_lazy2(enc(l1), enc(l2), out)   # lazy - lazy (the one-multiplication butterfly)
        assert dec(out) % P == (l1 - l2) % P
        small = r.getrandbits(r.choice([4, 20, 33]))  # a - b just below -p: the double-borrow path
        big = 2**256 - 1 - r.getrandbits(r.choice([2, 20, 33]))
        shim.fph_sub_lazy2(enc(small), enc(big), out)
        assert dec(out) % P == (small - big) % P
        shim.fph_mont_mul(enc(a), enc(b), out)      # the CIOS Montgomery alternative (a*b*R^-1)
        assert dec(out) == a * b * rinv % P


def test_short_add_sub_rare_paths(shim):
    """fp_add_lazy_f / fp_sub_lazy2_f (used by the butterfly kernels) correct only the low two limbs and
    branch on the rare ripple: drive the ripple and double-wrap paths on purpose."""
    r = random.Random(7)
    out = A8()
    M = 2**256
    for it in range(30000):
        a = rnd(r)
        kind = it % 6
        if kind == 0:      # wrapped sum whose low 64 bits are within DELTA of 2^64: the DELTA fold ripples
            tgt = (r.getrandbits(192) << 64) | (2**64 - 1 - r.getrandbits(r.choice([3, 20, 33])))
            b = (tgt + M - a) % M
        elif kind == 1:    # ... and the high limbs are all ones: double wrap
            tgt = ((2**192 - 1) << 64) | (2**64 - 1 - r.getrandbits(r.choice([3, 20, 33])))
            b = (tgt + M - a) % M
        elif kind == 2:    # a - b borrows and the wrapped difference has low 64 bits below DELTA
            tgt = (r.getrandbits(192) << 64) | r.getrandbits(r.choice([3, 20, 33]))
            b = (a - tgt) % M
        elif kind == 3:    # ... with all-zero high limbs: the borrow falls off the top
            tgt = r.getrandbits(r.choice([3, 20, 33]))
            b = (a - tgt) % M
        else:
            b = rnd(r)
        shim.fph_add_lazy_f(enc(a), enc(b), out)
        assert dec(out) % P == (a + b) % P, (kind, hex(a), hex(b))
        shim.fph_sub_lazy2_f(enc(a), enc(b), out)
        assert dec(out) % P == (a - b) % P, (kind, hex(a), hex(b))


def test_inverse_sqrt_pow_and_constants(shim):
    r = random.Random(3)
    out = A8()
    for _ in range(300):
        a = r.choice([0, 1, 2, P - 1, r.getrandbits(256) % P])
        shim.fph_inv(enc(a), out)
        assert dec(out) == (pow(a, -1, P) if a else 0)
        shim.fph_sqrt(enc(a), out)
        assert dec(out) == pow(a, (P + 1) // 4, P)
        e = r.getrandbits(r.choice([1, 5, 22, 64]))
        shim.fph_pow(enc(a), ctypes.c_uint64(e), out)
        assert dec(out) == pow(a, e, P)
    R, RI = A8(), A8()
    shim.fph_consts(R, RI)
    assert dec(R) == 2**256 % P and dec(RI) == pow(2**256, -1, P)


def test_montgomery_api_with_plain_tables(shim):
    """the identity the engine rests on: plain_mul(c, x*R) == (c*x)*R == mont_mul(c*R, x*R)"""
    r = random.Random(4)
    R = 2**256 % P
    out, out2 = A8(), A8()
    for _ in range(2000):
        c, x = r.getrandbits(256) % P, r.getrandbits(256) % P
        shim.fph_mul(enc(c), enc(x * R % P), out)
        shim.fph_mont_mul(enc(c * R % P), enc(x * R % P), out2)
        assert dec(out) == dec(out2) == c * x * R % P
