"""T0 oracle: CPU restatement of zk-toolkit's BLS12-381 MSM / Groth16 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path: it
may be imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, as the checker and never as the
thing that is shipped or measured as the GPU result.

Every function restates the algorithm of the reference source it cites
(paths relative to /root/reference/src).  The reference is Rust on
``num-bigint`` (exact integers), so Python ``int`` reproduces it bit for bit.
Parity is pinned by the reference's own known-answer tests, extracted into
``tests/golden/*.json`` by ``tests/golden/extract_kats.py`` and checked in
``tests/test_oracle_kats.py``.

Conventions kept from the reference:
  * Fq2(u1, u0), Fq6(v2, v1, v0), Fq12(w1, w0): highest-degree coefficient first
    (building_block/curves/bls12_381/fq2.rs:15-24, fq6.rs:15-19, fq12.rs:17-28).
  * Points are affine ``(x, y)`` tuples or the singleton ``INF`` (``AtInfinity``,
    g1_point.rs:33-36, g2_point.rs:31-34).
  * Randomness (CRS trapdoor, r, s) is an explicit argument: the reference draws
    it from OS entropy (field/prime_field.rs:73-85) so its proofs are not
    reproducible; the algorithm is otherwise identical.
"""
from __future__ import annotations

# --------------------------------------------------------------------------- params
# building_block/curves/bls12_381/params.rs:8-16
Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
EMBEDDING_DEGREE = 12

INF = None  # AtInfinity


# --------------------------------------------------------------------------- Fp / Fr
def ext_euclid_inv(v: int, order: int) -> int:
    """field/prime_field_elem.rs:379-432 (safe_inv): extended Euclid on signed ints."""
    if v == 0:
        raise ZeroDivisionError("Cannot find inverse of zero")
    r0, r1 = v, order
    x0, y0, x1, y1 = 1, 0, 0, 1
    while r1 != 0:
        q = r0 // r1
        r2 = r0 % r1
        x2 = x0 - x1 * q
        y2 = y0 - y1 * q
        r0, r1 = r1, r2
        x0, y0, x1, y1 = x1, y1, x2, y2
    new_v = x0
    if new_v < 0:
        while new_v < 0:
            new_v += order
    elif new_v >= order:
        new_v %= order
    return new_v


class Fq1:
    """PrimeFieldElem over the base field q (field/prime_field_elem.rs:33-37, fq1.rs:13)."""
    __slots__ = ("e",)
    P = Q

    def __init__(self, e: int):
        self.e = e % self.P  # prime_field_elem.rs:263-272

    def __add__(self, o):  # plus, :278-286
        e = self.e + o.e
        if e >= self.P:
            e -= self.P
        return self._raw(e)

    def __sub__(self, o):  # minus, :288-300
        if self.e < o.e:
            return self._raw(self.P - (o.e - self.e))
        return self._raw(self.e - o.e)

    def __mul__(self, o):  # times, :302-308
        return self._raw((self.e * o.e) % self.P)

    def __neg__(self):  # negate, :448-457
        return self._raw(0 if self.e == 0 else self.P - self.e)

    def sq(self):  # :330-335
        return self._raw((self.e * self.e) % self.P)

    def inv(self):  # :434-436
        return self._raw(ext_euclid_inv(self.e, self.P))

    def is_zero(self):
        return self.e == 0

    def __eq__(self, o):
        return isinstance(o, Fq1) and type(o).P == type(self).P and self.e == o.e

    def __hash__(self):
        return hash(self.e)

    def __repr__(self):
        return f"Fq1({self.e})"

    @classmethod
    def _raw(cls, e):
        o = object.__new__(cls)
        o.e = e
        return o

    @classmethod
    def zero(cls):
        return cls._raw(0)

    @classmethod
    def from_int(cls, n):
        return cls(n)


class Fr(Fq1):
    """PrimeFieldElem over the subgroup order r (params.rs:13-16)."""
    __slots__ = ()
    P = R

    def __repr__(self):
        return f"Fr({self.e})"


# --------------------------------------------------------------------------- Fq2
class Fq2:
    """Fq[u]/(u^2+1); fq2.rs:15-24 (note the (u1, u0) argument order)."""
    __slots__ = ("u1", "u0")

    def __init__(self, u1: Fq1, u0: Fq1):
        self.u1, self.u0 = u1, u0

    def __add__(self, o):  # fq2.rs:96-113
        return Fq2(self.u1 + o.u1, self.u0 + o.u0)

    def __sub__(self, o):  # fq2.rs:115-132
        return Fq2(self.u1 - o.u1, self.u0 - o.u0)

    def __mul__(self, o):  # fq2.rs:134-151
        return Fq2(self.u1 * o.u0 + self.u0 * o.u1, self.u0 * o.u0 - self.u1 * o.u1)

    def __neg__(self):  # fq2.rs:82-94
        return Fq2.zero() - self

    def sq(self):  # fq2.rs:34-36
        return self * self

    def inv(self):  # fq2.rs:26-32
        factor = (self.u1 * self.u1 + self.u0 * self.u0).inv()
        return Fq2((-self.u1) * factor, self.u0 * factor)

    def reduce(self):  # fq2.rs:52-59: multiply by (1 + u)
        return Fq2(self.u1 + self.u0, self.u0 - self.u1)

    def is_zero(self):  # fq2.rs:39-41
        return self.u0.is_zero() and self.u1.is_zero()

    def __eq__(self, o):
        return isinstance(o, Fq2) and self.u1 == o.u1 and self.u0 == o.u0

    def __hash__(self):
        return hash((self.u1.e, self.u0.e))

    def __repr__(self):
        return f"Fq2(u1={self.u1.e}, u0={self.u0.e})"

    @staticmethod
    def zero():
        return Fq2(Fq1.zero(), Fq1.zero())

    @staticmethod
    def from_int(n):  # fq2.rs:69-74
        return Fq2(Fq1.zero(), Fq1(n))


# --------------------------------------------------------------------------- Fq6 / Fq12 (verifier only)
class Fq6:
    """Fq2[v]/(v^3 - (1+u)); fq6.rs:15-19,64-72."""
    __slots__ = ("v2", "v1", "v0")

    def __init__(self, v2, v1, v0):
        self.v2, self.v1, self.v0 = v2, v1, v0

    def __add__(self, o):
        return Fq6(self.v2 + o.v2, self.v1 + o.v1, self.v0 + o.v0)

    def __sub__(self, o):
        return Fq6(self.v2 - o.v2, self.v1 - o.v1, self.v0 - o.v0)

    def __neg__(self):
        return Fq6.zero() - self

    def __mul__(self, o):  # fq6.rs:148-171
        t0 = self.v0 * o.v0
        t1 = self.v0 * o.v1 + self.v1 * o.v0
        t2 = self.v0 * o.v2 + self.v1 * o.v1 + self.v2 * o.v0
        t3 = (self.v1 * o.v2 + self.v2 * o.v1).reduce()
        t4 = (self.v2 * o.v2).reduce()
        return Fq6(t2, t1 + t4, t0 + t3)

    def inv(self):  # fq6.rs:23-37
        t0 = self.v0 * self.v0 - (self.v1 * self.v2).reduce()
        t1 = (self.v2 * self.v2).reduce() - self.v0 * self.v1
        t2 = self.v1 * self.v1 - self.v0 * self.v2
        factor = (self.v0 * t0 + (self.v2 * t1).reduce() + (self.v1 * t2).reduce()).inv()
        return Fq6(t2 * factor, t1 * factor, t0 * factor)

    def reduce(self):  # fq6.rs:54-61: multiply by v
        return Fq6(self.v1, self.v0, self.v2.reduce())

    def __eq__(self, o):
        return self.v2 == o.v2 and self.v1 == o.v1 and self.v0 == o.v0

    @staticmethod
    def zero():
        return Fq6(Fq2.zero(), Fq2.zero(), Fq2.zero())

    @staticmethod
    def from_int(n):  # fq6.rs:82-90
        return Fq6(Fq2.zero(), Fq2.zero(), Fq2.from_int(n))


class Fq12:
    """Fq6[w]/(w^2 - v); fq12.rs:17-28."""
    __slots__ = ("w1", "w0")

    def __init__(self, w1, w0):
        self.w1, self.w0 = w1, w0

    def __add__(self, o):
        return Fq12(self.w1 + o.w1, self.w0 + o.w0)

    def __sub__(self, o):
        return Fq12(self.w1 - o.w1, self.w0 - o.w0)

    def __neg__(self):
        return Fq12.zero() - self

    def __mul__(self, o):  # fq12.rs:135-152
        return Fq12(self.w1 * o.w0 + self.w0 * o.w1,
                    self.w0 * o.w0 + (self.w1 * o.w1).reduce())

    def inv(self):  # fq12.rs:31-40
        factor = (self.w0 * self.w0 - (self.w1 * self.w1).reduce()).inv()
        return Fq12((-self.w1) * factor, self.w0 * factor)

    def pow(self, exp: int):  # fq12.rs:42-57
        base, acc = self, Fq12.from_int(1)
        while exp != 0:
            if exp & 1:
                acc = acc * base
            base = base * base
            exp >>= 1
        return acc

    def __eq__(self, o):
        return self.w1 == o.w1 and self.w0 == o.w0

    @staticmethod
    def zero():
        return Fq12(Fq6.zero(), Fq6.zero())

    @staticmethod
    def from_int(n):  # fq12.rs:59-66
        return Fq12(Fq6.zero(), Fq6.from_int(n))

    @staticmethod
    def from_fq1(x: Fq1):
        return Fq12.from_int(x.e)


# --------------------------------------------------------------------------- group law (G1, G2)
def affine_add(p, q):
    """impl_affine_add!, curves/macros.rs:35-163.  Works over Fq1 (G1) and Fq2 (G2)."""
    if p is INF and q is INF:          # :44-46
        return INF
    if p is INF:                       # :47-49
        return q
    if q is INF:                       # :50-52
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2 and y1 != y2:          # :53-56 vertical line
        return INF
    if x1 == x2 and y1 == y2:          # :57-108 tangent
        if y1.is_zero():               # :61-63
            return INF
        x1_sq = x1.sq()
        m1 = x1_sq + x1_sq + x1_sq
        m2 = y1 + y1
        m = m1 * m2.inv()
        p3x = m.sq() - (x1 + x1)
        p3y_neg = m * (x1 - p3x) - y1
        return (p3x, p3y_neg)
    m = (y2 - y1) * (x2 - x1).inv()    # :109-152 chord
    p3x = m.sq() - x1 - x2
    p3y = m * (p3x - x1) + y1
    return (p3x, -p3y)


def point_neg(p):
    """g1_point.rs:178-195 / g2_point.rs (Neg)."""
    if p is INF:
        return INF
    return (p[0], -p[1])


def scalar_mul(p, n: int):
    """impl_scalar_mul_point!, curves/macros.rs:2-32: LSB-first double-and-add on the raw integer."""
    res = INF
    pt = p
    while n != 0:
        if n & 1:
            res = affine_add(res, pt)
        pt = affine_add(pt, pt)
        n >>= 1
    return res


def msm(points, scalars):
    """Polynomial::eval_with_g{1,2}_hidings, field/polynomial.rs:272-293.

    Uses the first len(scalars) points; IndexError if there are fewer points
    (the reference panics on the slice index, polynomial.rs:278)."""
    s = INF
    for i in range(len(scalars)):
        s = affine_add(s, scalar_mul(points[i], int(scalars[i])))
    return s


# generators: g1_point.rs:38-47, g2_point.rs:36-46
G1_GEN = (
    Fq1(0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb),
    Fq1(0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
)
G2_GEN = (
    Fq2(Fq1(0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
        Fq1(0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8)),
    Fq2(Fq1(0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be),
        Fq1(0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801)),
)


def g1_is_on_curve(p):
    """g1_point.rs:101-107: y^2 = x^3 + 4."""
    if p is INF:
        return True
    x, y = p
    return y.sq() == x * x * x + Fq1(4)


def g2_is_on_curve(p):
    """g2_point.rs:76-81: y^2 = x^3 + 4(1+u)."""
    if p is INF:
        return True
    x, y = p
    return y.sq() == x * x * x + Fq2.from_int(4).reduce()


def g1(x: int, y: int):
    return (Fq1(x), Fq1(y))


def g2(x1: int, x0: int, y1: int, y0: int):
    return (Fq2(Fq1(x1), Fq1(x0)), Fq2(Fq1(y1), Fq1(y0)))


# --------------------------------------------------------------------------- pairing (verifier only)
def _untwist(p):
    """From<&G2Point> for G12Point, g12_point.rs:47-67."""
    x, y = p
    one = Fq2.from_int(1)
    root = Fq6(Fq2.zero(), one, Fq2.zero())
    x6 = Fq6(Fq2.zero(), Fq2.zero(), x)
    y6 = Fq6(Fq2.zero(), Fq2.zero(), y)
    x12 = Fq12(Fq6.zero(), x6) * Fq12(Fq6.zero(), root).inv()
    y12 = Fq12(Fq6.zero(), y6) * Fq12(root, Fq6.zero()).inv()
    return (x12, y12)


def _g1_to_12(p):
    """From<&G1Point> for G12Point, g12_point.rs:29-44."""
    return (Fq12.from_fq1(p[0]), Fq12.from_fq1(p[1]))


class _Line:
    """RationalFunction, rational_function.rs:11-102."""
    __slots__ = ("x", "y", "slope")

    def __init__(self, p, q, to12):
        if p is INF or q is INF:
            raise ValueError("Both points need to be rational")
        x1, y1 = to12(p)
        if p[0] == q[0] and p[1] == q[1]:                # tangent :71-83
            two, three = Fq12.from_int(2), Fq12.from_int(3)
            self.x, self.y = x1, y1
            self.slope = three * x1 * x1 * (two * y1).inv()
        elif q[0] == p[0] and q[1] == -p[1]:             # vertical :85-89
            self.x, self.y, self.slope = x1, None, None
        else:                                            # chord :91-101
            x2, y2 = to12(q)
            self.x, self.y = x1, y1
            self.slope = (y2 - y1) * (x2 - x1).inv()

    def eval(self, q12):  # :45-62
        X, Y = q12
        if self.slope is None:
            return X + (-self.x)
        return (-self.slope) * X + Y + (-self.y) + self.slope * self.x


def _l_bits():
    """Pairing::new, pairing.rs:58-73: bits of r-1, MSB first, MSB dropped."""
    l = R - 1
    bits = []
    while l:
        bits.append(bool(l & 1))
        l >>= 1
    bits.reverse()
    return bits[1:]


_L_BITS = _l_bits()


def miller_g1_g2(p, q):
    """impl_miller_algorithm!(G1Point, G2Point, ...), pairing.rs:20-55."""
    f = Fq12.from_int(1)
    V = p
    q12 = _untwist(q)
    for bit in _L_BITS:
        v2 = affine_add(V, V)
        g_num = _Line(V, V, _g1_to_12)
        g_den = _Line(v2, point_neg(v2), _g1_to_12)
        f = (f * f) * g_num.eval(q12) * g_den.eval(q12).inv()
        V = v2
        if bit:
            vp = affine_add(V, p)
            g_num = _Line(V, p, _g1_to_12)
            g_den = _Line(vp, point_neg(vp), _g1_to_12)
            f = f * g_num.eval(q12) * g_den.eval(q12).inv()
            V = vp
    return f


def miller_g2_g1(p, q):
    """impl_miller_algorithm!(G2Point, G1Point, ...), pairing.rs:57."""
    f = Fq12.from_int(1)
    V = p
    q12 = _g1_to_12(q)
    for bit in _L_BITS:
        v2 = affine_add(V, V)
        g_num = _Line(V, V, _untwist)
        g_den = _Line(v2, point_neg(v2), _untwist)
        f = (f * f) * g_num.eval(q12) * g_den.eval(q12).inv()
        V = v2
        if bit:
            vp = affine_add(V, p)
            g_num = _Line(V, p, _untwist)
            g_den = _Line(vp, point_neg(vp), _untwist)
            f = f * g_num.eval(q12) * g_den.eval(q12).inv()
            V = vp
    return f


_TATE_EXP = (Q ** EMBEDDING_DEGREE - 1) // R


def tate(p1, p2) -> Fq12:
    """Pairing::tate, pairing.rs:86-100."""
    return miller_g1_g2(p1, p2).pow(_TATE_EXP)


def weil(p1, p2) -> Fq12:
    """Pairing::weil, pairing.rs:75-84."""
    return miller_g1_g2(p1, p2) * miller_g2_g1(p2, p1).inv()


# --------------------------------------------------------------------------- polynomials over Fr
class Polynomial:
    """field/polynomial.rs:32-36; coeffs[i] multiplies x^i; ints mod r."""
    __slots__ = ("coeffs",)

    def __init__(self, coeffs, normalize=True):
        if len(coeffs) == 0:
            raise ValueError("coeffs is empty")                       # :120
        c = [int(x) % R for x in coeffs]
        if normalize:                                                  # :139-152
            n = len(c)
            while n > 1 and c[n - 1] == 0:
                n -= 1
            c = c[:n]
        self.coeffs = c

    @staticmethod
    def zero():
        return Polynomial([0])

    def is_zero(self):
        return len(self.coeffs) == 1 and self.coeffs[0] == 0

    def __len__(self):
        return len(self.coeffs)

    def plus(self, o):  # :154-171
        a, b = (self.coeffs, o.coeffs) if len(self.coeffs) < len(o.coeffs) else (o.coeffs, self.coeffs)
        return Polynomial([(a[i] + b[i]) % R if i < len(a) else b[i] for i in range(len(b))])

    def multiply_by(self, o):  # :173-190 (result is NOT normalized in the reference)
        out = [0] * (len(self.coeffs) + len(o.coeffs) - 1)
        for i, a in enumerate(self.coeffs):
            for j, b in enumerate(o.coeffs):
                out[i + j] = (out[i + j] + a * b) % R
        return Polynomial(out, normalize=False)

    def minus(self, o):  # :192-202
        assert len(self.coeffs) >= len(o.coeffs)
        c = list(self.coeffs)
        for i, b in enumerate(o.coeffs):
            c[i] = (c[i] - b) % R
        return Polynomial(c)

    def scale(self, k: int):  # Mul<PrimeFieldElem>, :351-361
        return Polynomial([(c * k) % R for c in self.coeffs])

    def divide_by(self, divisor):
        """:204-238.  Returns (quotient, remainder-or-None); quotient is not normalized."""
        dividend = Polynomial(self.coeffs, normalize=False)
        qdeg = len(dividend) - len(divisor)
        dc = divisor.coeffs[-1]
        assert dc != 0
        dc_inv = ext_euclid_inv(dc, R)
        quot = [0] * (qdeg + 1)
        while not dividend.is_zero() and len(dividend) >= len(divisor):
            term_coeff = (dividend.coeffs[-1] * dc_inv) % R
            term_degree = len(dividend) - len(divisor)
            quot[term_degree] = term_coeff
            term = [0] * term_degree + [term_coeff]
            sub = divisor.multiply_by(Polynomial(term))
            dividend = dividend.minus(sub)
        q = Polynomial(quot, normalize=False)
        return (q, None) if dividend.is_zero() else (q, dividend)

    def eval_at(self, x: int) -> int:  # :240-249
        mult, s = 1, 0
        for c in self.coeffs:
            s = (s + c * mult) % R
            mult = (mult * x) % R
        return s

    def eval_with_g1_hidings(self, powers):  # :272-281
        return msm(powers, self.coeffs)

    def eval_with_g2_hidings(self, powers):  # :284-293
        return msm(powers, self.coeffs)


# --------------------------------------------------------------------------- R1CS -> QAP
def qap_build_polynomial(target_vals):
    """QAP::build_polynomial, zk/w_trusted_setup/qap/qap.rs:33-97: Lagrange over x = 1..n."""
    n = len(target_vals)
    polys = []
    for tx in range(1, n + 1):
        tv = target_vals[tx - 1] % R
        if tv == 0:
            polys.append(Polynomial([0]))
            continue
        acc = Polynomial([1]).multiply_by(Polynomial([tv]))
        denom = 1
        for i in range(1, n + 1):
            if i == tx:
                continue
            acc = acc.multiply_by(Polynomial([(-i) % R, 1]))
            denom = (denom * ((tx - i) % R)) % R
        acc = acc.multiply_by(Polynomial([ext_euclid_inv(denom, R)]))
        polys.append(acc)
    res = polys[0]
    for p in polys[1:]:
        res = res.plus(p)
    return res


def qap_build_t(n: int):
    """QAP::build_t, qap.rs:115-135."""
    acc = Polynomial([1])
    for i in range(1, n + 1):
        acc = acc.multiply_by(Polynomial([(-i) % R, 1]))
    return acc


class QAP:
    """QAP::build, qap.rs:137-203.  rows_{a,b,c}: list (one per constraint) of {wire: coeff}."""

    def __init__(self, rows_a, rows_b, rows_c, num_wires):
        n = len(rows_a)
        col = lambda rows, w: [rows[k].get(w, 0) for k in range(n)]
        self.vi = [qap_build_polynomial(col(rows_a, w)) for w in range(num_wires)]
        self.wi = [qap_build_polynomial(col(rows_b, w)) for w in range(num_wires)]
        self.yi = [qap_build_polynomial(col(rows_c, w)) for w in range(num_wires)]
        self.num_constraints = n

    def build_p(self, witness):  # qap.rs:99-112
        v = w = y = Polynomial.zero()
        for i, wit in enumerate(witness):
            v = v.plus(self.vi[i].scale(wit))
            w = w.plus(self.wi[i].scale(wit))
            y = y.plus(self.yi[i].scale(wit))
        return v.multiply_by(w).minus(y)


def r1cs_validate(rows_a, rows_b, rows_c, witness):
    """R1CS::validate, qap/r1cs.rs:61-74."""
    dot = lambda row: sum(c * witness[w] for w, c in row.items()) % R
    for a, b, c in zip(rows_a, rows_b, rows_c):
        if (dot(a) * dot(b)) % R != dot(c):
            raise ValueError("constraint doesn't hold")


# The one circuit the reference proves: "(x * x * x) + x + 5 == 35"
# (groth16/zktoolkit_based/prover.rs:162-177).  Gates pinned by qap/gate.rs:195-235,
# wire order [1, x, out, t1, t2, t3, t4] by qap/r1cs_tmpl.rs:22-51, Num overwrites wire 0
# (r1cs_tmpl.rs:82-84).
CONFIG1 = dict(
    rows_a=[{1: 1}, {1: 1}, {1: 1, 0: 5}, {4: 1, 5: 1}, {6: 1}],
    rows_b=[{1: 1}, {3: 1}, {0: 1}, {0: 1}, {0: 1}],
    rows_c=[{3: 1}, {4: 1}, {5: 1}, {6: 1}, {2: 1}],
    witness=[1, 3, 35, 9, 27, 8, 35],
    mid_beg=3,
)


# --------------------------------------------------------------------------- Groth16
class Prover:
    """groth16/zktoolkit_based/prover.rs:35-147."""

    def __init__(self, rows_a, rows_b, rows_c, witness, mid_beg):  # Prover::new, :49-93
        r1cs_validate(rows_a, rows_b, rows_c, witness)
        qap = QAP(rows_a, rows_b, rows_c, len(witness))
        self.n = len(rows_a)
        self.t = qap_build_t(self.n)
        p = qap.build_p(witness)
        h, rem = p.divide_by(self.t)
        if rem is not None:
            raise ValueError("p should be divisible by t")            # :69
        self.h = h
        self.l = mid_beg - 1                                            # :73-76
        self.m = len(witness) - 1                                       # :77
        self.wires = [w % R for w in witness]
        self.ui, self.vi, self.wi = qap.vi, qap.wi, qap.yi              # :89-91

    def statement(self):  # wires.rs:27-30
        return self.wires[: self.l + 1]

    def prove(self, crs, r: int, s: int):
        """Prover::prove, :96-147, with r and s supplied by the caller."""
        sum_A = sum_B = sum_Bg1 = INF
        for i in range(self.m + 1):                                     # :108-117
            ai = self.wires[i]
            sum_A = affine_add(sum_A, scalar_mul(self.ui[i].eval_with_g1_hidings(crs.g1_xi), ai))
            sum_B = affine_add(sum_B, scalar_mul(self.vi[i].eval_with_g2_hidings(crs.g2_xi), ai))
            sum_Bg1 = affine_add(sum_Bg1, scalar_mul(self.vi[i].eval_with_g1_hidings(crs.g1_xi), ai))
        A = affine_add(affine_add(crs.g1_alpha, sum_A), scalar_mul(crs.g1_delta, r))       # :118
        B = affine_add(affine_add(crs.g2_beta, sum_B), scalar_mul(crs.g2_delta, s))        # :119
        B_g1 = affine_add(affine_add(crs.g1_beta, sum_Bg1), scalar_mul(crs.g1_delta, s))   # :120
        acc = INF
        wit_beg = self.l + 1
        for i in range(wit_beg, self.m + 1):                            # :128-131
            acc = affine_add(acc, scalar_mul(crs.g1_uvw_wit[i - wit_beg], self.wires[i]))
        ht = self.h.eval_with_g1_hidings(crs.g1_xt_by_delta)            # :133
        C = affine_add(acc, ht)                                         # :135-139
        C = affine_add(C, scalar_mul(A, s))
        C = affine_add(C, scalar_mul(B_g1, r))
        C = affine_add(C, point_neg(scalar_mul(scalar_mul(crs.g1_delta, r), s)))
        return (A, B, C)


class CRS:
    """groth16/zktoolkit_based/crs.rs:17-146; trapdoor supplied by the caller."""

    def __init__(self, prover: Prover, alpha, beta, gamma, delta, x, with_pairing=True):
        g, h = G1_GEN, G2_GEN
        inv = lambda a: ext_euclid_inv(a % R, R)

        def uvw_div(lo, hi, div):                                      # :65-83
            out = []
            for i in range(lo, hi + 1):
                ui = (beta * prover.ui[i].eval_at(x)) % R
                vi = (alpha * prover.vi[i].eval_at(x)) % R
                wi = prover.wi[i].eval_at(x)
                out.append(scalar_mul(g, ((ui + vi + wi) * div) % R))
            return out

        self.g1_uvw_stmt = uvw_div(0, prover.l, inv(gamma))            # :85
        self.g1_uvw_wit = uvw_div(prover.l + 1, prover.m, inv(delta))  # :86

        def n_pows(gen):                                               # :88-102
            ys, xp = [], 1
            for _ in range(prover.n):
                ys.append(scalar_mul(gen, xp))
                xp = (xp * x) % R
            return ys

        self.g1_xi = n_pows(g)
        t = qap_build_t(prover.n).eval_at(x)                           # :106-116
        xs, xp = [], 1
        for _ in range(prover.n):
            xs.append(scalar_mul(g, (xp * t % R) * inv(delta) % R))
            xp = (xp * x) % R
        self.g1_xt_by_delta = xs
        self.g1_alpha = scalar_mul(g, alpha)
        self.g1_beta = scalar_mul(g, beta)
        self.g1_delta = scalar_mul(g, delta)
        self.g2_xi = n_pows(h)
        self.g2_beta = scalar_mul(h, beta)
        self.g2_gamma = scalar_mul(h, gamma)
        self.g2_delta = scalar_mul(h, delta)
        self.gt_alpha_beta = tate(self.g1_alpha, self.g2_beta) if with_pairing else None  # :137-139


def verify(proof, crs: CRS, stmt_wires) -> bool:
    """Verifier::verify, groth16/zktoolkit_based/verifier.rs:30-54."""
    A, B, C = proof
    lhs = tate(A, B)
    sum_term = INF
    for i, ai in enumerate(stmt_wires):
        sum_term = affine_add(sum_term, scalar_mul(crs.g1_uvw_stmt[i], ai))
    rhs = crs.gt_alpha_beta * tate(sum_term, crs.g2_gamma) * tate(C, crs.g2_delta)
    return lhs == rhs


# --------------------------------------------------------------------------- flat-limb helpers (ABI layout)
def g1_to_limbs(p):
    """(x, y) -> 24 little-endian u32 limbs (x then y), canonical; INF -> zeros."""
    if p is INF:
        return [0] * 24
    out = []
    for v in (p[0].e, p[1].e):
        out += [(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
    return out


def g2_to_limbs(p):
    """ABI order x.u0, x.u1, y.u0, y.u1 (12 limbs each)."""
    if p is INF:
        return [0] * 48
    out = []
    for v in (p[0].u0.e, p[0].u1.e, p[1].u0.e, p[1].u1.e):
        out += [(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
    return out


# --------------------------------------------------------------------------- Pinocchio (second consumer of the MSM seam)
class PinocchioProver:
    """zk/w_trusted_setup/pinocchio/prover.rs:35-93 (Prover::new) on explicit R1CS rows."""

    def __init__(self, rows_a, rows_b, rows_c, witness, mid_beg):
        r1cs_validate(rows_a, rows_b, rows_c, witness)
        qap = QAP(rows_a, rows_b, rows_c, len(witness))
        self.num_constraints = len(rows_a)
        self.t = qap_build_t(self.num_constraints)
        self.p = qap.build_p(witness)                                              # :64
        polys = qap.vi + qap.wi + qap.yi + [self.p, self.t]
        self.max_degree = max(len(x) - 1 for x in polys) + 1                       # :66-76
        self.witness = [w % R for w in witness]
        self.mid_beg = mid_beg
        self.vi, self.wi, self.yi = qap.vi, qap.wi, qap.yi

    def io(self):   # witness.rs:20-23
        return self.witness[: self.mid_beg]

    def mid(self):  # witness.rs:25-27
        return self.witness[self.mid_beg:]


class PinocchioCRS:
    """pinocchio/crs.rs:55-161 with the trapdoor supplied by the caller."""

    def __init__(self, p: PinocchioProver, r_v, r_w, alpha_v, alpha_w, alpha_y, beta, gamma, s):
        g1, g2 = G1_GEN, G2_GEN
        m = scalar_mul
        r_y = (r_v * r_w) % R
        g1_v, g1_w, g2_w, g1_y = m(g1, r_v), m(g1, r_w), m(g2, r_w), m(g1, r_y)
        mid = list(range(p.mid_beg, len(p.witness)))
        io = list(range(p.mid_beg))
        ev = lambda poly: poly.eval_at(s)
        self.vk_mid = [m(g1_v, ev(p.vi[i])) for i in mid]
        self.g1_wk_mid = [m(g1_w, ev(p.wi[i])) for i in mid]
        self.g2_wk_mid = [m(g2_w, ev(p.wi[i])) for i in mid]
        self.yk_mid = [m(g1_y, ev(p.yi[i])) for i in mid]
        self.alpha_vk_mid = [m(m(g1_v, alpha_v), ev(p.vi[i])) for i in mid]
        self.alpha_wk_mid = [m(m(g1_w, alpha_w), ev(p.wi[i])) for i in mid]
        self.alpha_yk_mid = [m(m(g1_y, alpha_y), ev(p.yi[i])) for i in mid]
        spow, self.si = 1, []
        for _ in range(p.max_degree):                                              # s.pow_seq(max_degree), crs.rs:95-96
            self.si.append(m(g2, spow))
            spow = (spow * s) % R
        self.beta_vwy_k_mid = [
            affine_add(affine_add(m(m(g1_v, beta), ev(p.vi[i])), m(m(g1_w, beta), ev(p.wi[i]))), m(m(g1_y, beta), ev(p.yi[i])))
            for i in mid]
        self.one_g1, self.one_g2 = m(g1, 1), m(g2, 1)
        self.alpha_v, self.alpha_w, self.alpha_y = m(g2, alpha_v), m(g1, alpha_w), m(g2, alpha_y)
        self.gamma, self.beta_gamma = m(g2, gamma), m(m(g2, gamma), beta)
        self.t = m(g1_y, ev(p.t))
        self.vk_io = [m(g1_v, ev(p.vi[i])) for i in io]
        self.wk_io = [m(g2_w, ev(p.wi[i])) for i in io]
        self.yk_io = [m(g1_y, ev(p.yi[i])) for i in io]
        self.alpha_v_t, self.alpha_y_t, self.beta_t = m(self.t, alpha_v), m(self.t, alpha_y), m(self.t, beta)


def pinocchio_prove(p: PinocchioProver, crs: PinocchioCRS, delta_v, delta_y):
    """Prover::prove, pinocchio/prover.rs:95-171, with delta_v / delta_y supplied by the caller.
    Returns a dict with the nine proof elements (proof.rs)."""
    m, add = scalar_mul, affine_add
    mid = p.mid()
    v_mid_s = m(crs.t, delta_v)
    g1_w_mid_s = INF
    g2_w_mid_s = INF
    y_mid_s = m(crs.t, delta_y)
    alpha_v_mid_s = m(crs.alpha_v_t, delta_v)
    alpha_w_mid_s = INF
    alpha_y_mid_s = m(crs.alpha_y_t, delta_y)
    beta_vwy_mid_s = add(m(crs.beta_t, delta_v), m(crs.beta_t, delta_y))
    for i, w in enumerate(mid):                                                    # :118-128, the inline MSM loops
        v_mid_s = add(v_mid_s, m(crs.vk_mid[i], w))
        g1_w_mid_s = add(g1_w_mid_s, m(crs.g1_wk_mid[i], w))
        g2_w_mid_s = add(g2_w_mid_s, m(crs.g2_wk_mid[i], w))
        y_mid_s = add(y_mid_s, m(crs.yk_mid[i], w))
        alpha_v_mid_s = add(alpha_v_mid_s, m(crs.alpha_vk_mid[i], w))
        alpha_w_mid_s = add(alpha_w_mid_s, m(crs.alpha_wk_mid[i], w))
        alpha_y_mid_s = add(alpha_y_mid_s, m(crs.alpha_yk_mid[i], w))
        beta_vwy_mid_s = add(beta_vwy_mid_s, m(crs.beta_vwy_k_mid[i], w))
    h, rem = p.p.divide_by(p.t)                                                    # :131-134
    if rem is not None:
        raise ValueError("p should be divisible by t")
    h_s = h.eval_with_g2_hidings(crs.si)                                           # :136
    w_s = g2_w_mid_s
    for i, w in enumerate(p.io()[: len(crs.wk_io)]):                               # :139-142
        w_s = add(w_s, m(crs.wk_io[i], w))
    adj_h_s = add(add(h_s, m(w_s, delta_v)), point_neg(m(crs.one_g2, delta_y)))    # :144
    return dict(v_mid_s=v_mid_s, g1_w_mid_s=g1_w_mid_s, g2_w_mid_s=g2_w_mid_s, y_mid_s=y_mid_s, h_s=adj_h_s,
                alpha_v_mid_s=alpha_v_mid_s, alpha_w_mid_s=alpha_w_mid_s, alpha_y_mid_s=alpha_y_mid_s,
                beta_vwy_mid_s=beta_vwy_mid_s)


def pinocchio_verify(proof, crs: PinocchioCRS, witness_io) -> bool:
    """Verifier::verify, pinocchio/verifier.rs:27-87."""
    e, add, m = tate, affine_add, scalar_mul
    pr = proof
    vwy = add(add(pr["v_mid_s"], pr["g1_w_mid_s"]), pr["y_mid_s"])
    if e(pr["beta_vwy_mid_s"], crs.gamma) != e(vwy, crs.beta_gamma):
        return False
    if e(pr["alpha_v_mid_s"], crs.one_g2) != e(pr["v_mid_s"], crs.alpha_v):
        return False
    if e(pr["alpha_w_mid_s"], crs.one_g2) != e(crs.alpha_w, pr["g2_w_mid_s"]):
        return False
    if e(pr["alpha_y_mid_s"], crs.one_g2) != e(pr["y_mid_s"], crs.alpha_y):
        return False
    v_s, w_s, y_s = pr["v_mid_s"], pr["g2_w_mid_s"], pr["y_mid_s"]
    for i, w in enumerate(witness_io):
        v_s = add(v_s, m(crs.vk_io[i], w))
        w_s = add(w_s, m(crs.wk_io[i], w))
        y_s = add(y_s, m(crs.yk_io[i], w))
    return e(v_s, w_s) == e(crs.t, pr["h_s"]) * e(y_s, crs.one_g2)
