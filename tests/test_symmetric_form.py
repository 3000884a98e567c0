"""The algebra behind the one-product ("symmetric") butterflies of csrc/sym_kernel.cu, checked on the CPU
with Python big integers against the oracle:

  * every rational map of the secp256k1 chain is psi(x) = (x - b)^2 / x, so the two nodes of a pair are
    s and b^2/s, and g(x) = (x - b)/(x + b) takes opposite values on them;
  * EXTEND = Gamma^tgt . prod_j [[1, g],[1, -g]] . prod_j ([[1, g],[1, -g]]^src)^-1 . (Gamma^src)^-1 with
    Gamma_p = prod_j (s_j(p) + b_j) s_j(p)^(2^j - 1), the halvings folded into the pre-scale and the two
    level-0 butterflies merged into one product with g^tgt/g^src — bit-identical to the reference's 2x2
    matrix network (src/fftree.rs:72-120);
  * ENTER assembled from it with the fused combine constants equals the oracle's ENTER.
"""
import numpy as np
import pytest

from oracle import oracle as O

P = O.P


class SymTables:
    def __init__(self, tree, N):
        self.N, self.h = N, N // 2
        st = tree.subtree_with_size(N)
        self.f = O.from_mont(st.table("f"))
        self.xnn = O.from_mont(st.table("xnn_s"))
        self.L = self.h.bit_length() - 1
        f = self.f
        self.beta = []
        for j in range(self.L):
            B = 2 << j
            s0, s1, parent = f[2 * B], f[2 * B + B], f[B]
            b = (s0 + s1 - parent) * pow(2, -1, P) % P        # psi(x) = t  <=>  x^2 - (2b + t) x + b^2 = 0
            assert b * b % P == s0 * s1 % P
            self.beta.append(b)

    def node(self, mu, j, i, bit):
        B = 2 << j
        return self.f[2 * B + 2 * i + mu + bit * B]

    def g(self, mu, j, i):
        s0, b = self.node(mu, j, i, 0), self.beta[j]
        return (s0 - b) * pow(s0 + b, -1, P) % P

    def gamma(self, mu, p):
        acc = 1
        for j in range(self.L):
            i, bit = p & ((1 << j) - 1), (p >> j) & 1
            s = self.node(mu, j, i, bit)
            acc = acc * (s + self.beta[j]) * pow(s, (1 << j) - 1, P) % P
        return acc

    def extend(self, x, target, scaled=True):
        src, h, L = 1 - target, self.h, self.L
        inv2L = pow(2, -L, P)
        v = [xi * inv2L * pow(self.gamma(src, p), -1, P) % P for p, xi in enumerate(x)]
        for j in range(L - 1, 0, -1):                         # decompose levels above the centre
            for p in range(h):
                if not (p >> j) & 1:
                    q, gi = p + (1 << j), pow(self.g(src, j, p & ((1 << j) - 1)), -1, P)
                    v[p], v[q] = (v[p] + v[q]) % P, (v[p] - v[q]) * gi % P
        if L >= 1:                                            # centre: level 0 of both phases, one product
            c = self.g(target, 0, 0) * pow(self.g(src, 0, 0), -1, P) % P
            for p in range(0, h, 2):
                s, t = (v[p] + v[p + 1]) % P, c * (v[p] - v[p + 1]) % P
                v[p], v[p + 1] = (s + t) % P, (s - t) % P
        for j in range(1, L):
            for p in range(h):
                if not (p >> j) & 1:
                    q, t = p + (1 << j), self.g(target, j, p & ((1 << j) - 1)) * v[p + (1 << j)] % P
                    v[p], v[q] = (v[p] + t) % P, (v[p] - t) % P
        return [vi * self.gamma(target, p) % P for p, vi in enumerate(v)] if scaled else v


@pytest.fixture(scope="module")
def tree():
    return O.OracleTree.build(64)


def test_pairs_are_swapped_by_the_deck_involution(tree):
    t = SymTables(tree, 64)
    for mu in (0, 1):
        for j in range(t.L):
            for i in range(1 << j):
                s0, s1, b = t.node(mu, j, i, 0), t.node(mu, j, i, 1), t.beta[j]
                assert s0 * s1 % P == b * b % P
                assert ((s1 - b) * pow(s1 + b, -1, P) + t.g(mu, j, i)) % P == 0


@pytest.mark.parametrize("N", [4, 8, 64])
def test_symmetric_extend_equals_the_matrix_network(tree, N):
    t = SymTables(tree, N)
    st = tree.subtree_with_size(N)
    x = O.random_elements(N // 2, seed=N)
    xi = O.from_mont(x)
    for target in (0, 1):
        assert t.extend(xi, target) == O.from_mont(st.extend(x, target))


def test_enter_with_fused_combine_constants(tree):
    n = 64
    x = O.random_elements(n, seed=5)
    cur = O.from_mont(x)
    m = 2
    while m <= n:
        t, h = SymTables(tree, m), m // 2
        gam1 = [t.gamma(1, i) for i in range(h)]
        gx = [gam1[i] * t.xnn[2 * i + 1] % P for i in range(h)]
        nxt = []
        for off in range(0, n, m):
            u0, v0 = cur[off:off + h], cur[off + h:off + m]
            u1, v1 = t.extend(u0, 1, scaled=False), t.extend(v0, 1, scaled=False)
            blk = [0] * m
            for i in range(h):
                blk[2 * i] = (u0[i] + v0[i] * t.xnn[2 * i]) % P
                blk[2 * i + 1] = (gam1[i] * u1[i] + gx[i] * v1[i]) % P
            nxt += blk
        cur = nxt
        m *= 2
    assert cur == O.from_mont(tree.enter(x))


def test_enter_with_folded_prescale_tables(tree):
    """Engine::enter_range_serial keeps the data between two depths multiplied by the NEXT depth's pre-scale
    (P = 2^-L / Gamma^0 of the level that reads it) and carries the scale through the combine tables of
    Level::fold_tab: odd {gam1 Pn[2i+1], gx Pn[2i+1]}, even folded->folded {Pn[2i]/P[i], xnn[2i] Pn[2i]/P[i]},
    folded->plain {1/P[i], xnn[2i]/P[i]}, plain->folded {Pn[2i], xnn[2i] Pn[2i]}.  The EXTEND of a depth whose
    input is folded skips its pre-scale.  Same result as the oracle's ENTER, wherever the range is cut."""
    n = 64
    x = O.random_elements(n, seed=7)
    want = O.from_mont(tree.enter(x))

    def prescale(t):   # what SymTables.extend applies first: 2^-L / Gamma^src, src = S0
        inv2L = pow(2, -t.L, P)
        return [inv2L * pow(t.gamma(0, p), -1, P) % P for p in range(t.h)]

    def extend_unscaled_from_prescaled(t, v):   # SymTables.extend(target=1, scaled=False) minus its first line
        h, L, src, target = t.h, t.L, 0, 1
        v = list(v)
        for j in range(L - 1, 0, -1):
            for p in range(h):
                if not (p >> j) & 1:
                    q, gi = p + (1 << j), pow(t.g(src, j, p & ((1 << j) - 1)), -1, P)
                    v[p], v[q] = (v[p] + v[q]) % P, (v[p] - v[q]) * gi % P
        if L >= 1:
            c = t.g(target, 0, 0) * pow(t.g(src, 0, 0), -1, P) % P
            for p in range(0, h, 2):
                s, tt = (v[p] + v[p + 1]) % P, c * (v[p] - v[p + 1]) % P
                v[p], v[p + 1] = (s + tt) % P, (s - tt) % P
        for j in range(1, L):
            for p in range(h):
                if not (p >> j) & 1:
                    q, tt = p + (1 << j), t.g(target, j, p & ((1 << j) - 1)) * v[p + (1 << j)] % P
                    v[p], v[q] = (v[p] + tt) % P, (v[p] - tt) % P
        return v

    for cuts in ([], [8], [2, 32]):                       # block sizes after which the data returns to plain form
        cur, folded = O.from_mont(x), False
        m = 2
        while m <= n:
            t, h = SymTables(tree, m), m // 2
            Pm = prescale(t)
            out_folded = m < n and m not in cuts
            Pn = prescale(SymTables(tree, 2 * m)) if out_folded else None
            gam1 = [t.gamma(1, i) for i in range(h)]
            gx = [gam1[i] * t.xnn[2 * i + 1] % P for i in range(h)]
            nxt = []
            for off in range(0, n, m):
                u0, v0 = cur[off:off + h], cur[off + h:off + m]
                pu = u0 if folded else [a * b % P for a, b in zip(u0, Pm)]
                pv = v0 if folded else [a * b % P for a, b in zip(v0, Pm)]
                u1, v1 = extend_unscaled_from_prescaled(t, pu), extend_unscaled_from_prescaled(t, pv)
                blk = [0] * m
                for i in range(h):
                    e0 = (pow(Pm[i], -1, P) if folded else 1) * (Pn[2 * i] if out_folded else 1) % P
                    e1 = e0 * t.xnn[2 * i] % P
                    so = Pn[2 * i + 1] if out_folded else 1
                    blk[2 * i] = (e0 * u0[i] + e1 * v0[i]) % P
                    blk[2 * i + 1] = (gam1[i] * so * u1[i] + gx[i] * so * v1[i]) % P
                nxt += blk
            cur, folded = nxt, out_folded
            m *= 2
        assert cur == want, cuts
