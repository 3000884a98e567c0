"""Independent Python big-integer checker (ORACLE — TEST INFRASTRUCTURE ONLY).

Nothing here shares code with the C oracle or the CUDA path: plain `int` arithmetic mod p.
It re-creates the *definitions* the reference's tests use (src/lib.rs:108-186: Horner
evaluation at the leaves) and the definitions of the operations the reference leaves
untested (SURVEY.md 8c properties 3-8), plus an independent derivation of the leaves
(x-coordinates of coset_offset + i*G on the Good Curve of src/lib.rs:45-59).
"""
P = 2**256 - 2**32 - 977

A = 31172306031375832341232376275243462303334845584808513005362718476441963632613
BB = 45508371059383884471556188660911097844526467659576498497548207627741160623272
OFFSET = (105623886150579165427389078198493427091405550492761682382732004625374789850161,
          7709812624542158994629670452026922591039826164720902911013234773380889499231)
GEN = (41293412487153066667050767300223451435019201659857889215769525847559135483332,
       73754924733368840065089190002333366411120578552679996887076912271884749237510)
GEN_LOG_ORDER = 36


def horner(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def ec_add(p1, p2, a=A, bb=BB):
    """affine addition on y^2 = x^3 + a x^2 + bb x; None is the point at infinity"""
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2 and (y1 + y2) % P == 0:
        return None
    if x1 == x2:
        lam = (3 * x1 * x1 + 2 * a * x1 + bb) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - a - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def ec_mul(k, pt):
    res, acc = None, pt
    while k:
        if k & 1:
            res = ec_add(res, acc)
        acc = ec_add(acc, acc)
        k >>= 1
    return res


def leaves(n):
    """x(OFFSET + i*G_n), G_n = 2^(36-log n) * GEN  (reference src/lib.rs:66-78)"""
    g = ec_mul(1 << (GEN_LOG_ORDER - (n.bit_length() - 1)), GEN)
    out, acc = [], None
    for _ in range(n):
        out.append(ec_add(OFFSET, acc)[0])
        acc = ec_add(acc, g)
    return out


def poly_mul(a, b):
    res = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                res[i + j] = (res[i + j] + x * y) % P
    return res


def poly_from_roots(roots):
    poly = [1]
    for r in roots:
        poly = poly_mul(poly, [(-r) % P, 1])
    return poly


def poly_divmod(num, den):
    num = list(num)
    dd = len(den) - 1
    inv = pow(den[-1], -1, P)
    q = [0] * max(len(num) - dd, 1)
    for i in range(len(num) - 1, dd - 1, -1):
        c = num[i] * inv % P
        q[i - dd] = c
        if c:
            for j, d in enumerate(den):
                num[i - dd + j] = (num[i - dd + j] - c * d) % P
    return q, num[:dd]
