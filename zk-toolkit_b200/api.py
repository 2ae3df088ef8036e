"""Host-side mirror of the reference's interface on the MSM path (same names, argument meaning
and error behaviour), computing on the GPU through the C ABI.

reference                                                        here
---------------------------------------------------------------  ------------------------------
G1Point / G2Point  (curves/bls12_381/g1_point.rs:33-36,          G1Point / G2Point
                    g2_point.rs:31-34; Rational{x,y}|AtInfinity)
&G1Point + &G1Point (impl_affine_add!, curves/macros.rs:35-163)  p + q
-&G1Point           (g1_point.rs:178-195)                         -p
&G1Point * &Fq1     (impl_scalar_mul_point!, macros.rs:2-32)      p * k
Polynomial::eval_with_g1_hidings(&self, &[G1Point])              Polynomial.eval_with_g1_hidings
Polynomial::eval_with_g2_hidings(&self, &[G2Point])              Polynomial.eval_with_g2_hidings
(field/polynomial.rs:272-293)

`powers` may be a Python list of points (uploaded on every call, as the reference's slice
argument) or a resident G1Points / G2Points set (the CRS use: upload once, multiply many times).
There is no CPU implementation behind any of this: without the CUDA library and a B200 every
operation raises ZkmsmError.
"""
import os

import numpy as np

from . import _lib as L
from .context import Context, PointSet

Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001

_default_ctx = None


def default_context() -> Context:
    """One context per process on cuda:LOCAL_RANK (one process per GPU)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("ZKMSM_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _default_ctx


def _limbs(v, n=12):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def _int(a):
    return sum(int(x) << (32 * i) for i, x in enumerate(a))


def scalars_to_array(scalars):
    """ints -> (n, 8) uint32.  Field elements are < r; anything >= 2^255 is rejected by the device."""
    try:
        raw = b"".join(int(s).to_bytes(32, "little") for s in scalars)
    except OverflowError:
        raise ValueError("scalar out of range")
    return np.frombuffer(raw, dtype=np.uint32).reshape(-1, 8).copy() if raw else np.zeros((0, 8), dtype=np.uint32)


class _Point:
    """Affine point or AtInfinity; equality is coordinate equality (g1_point.rs:163-174)."""
    GROUP = 0
    WORDS = 0
    __slots__ = ("coords",)  # None = AtInfinity, else tuple of canonical ints

    def __init__(self, coords):
        self.coords = None if coords is None else tuple(int(c) % Q for c in coords)

    @classmethod
    def zero(cls):
        return cls(None)

    def is_zero(self):
        return self.coords is None

    def __eq__(self, o):
        return type(o) is type(self) and self.coords == o.coords

    def __hash__(self):
        return hash((self.GROUP, self.coords))

    def limbs(self):
        out = np.zeros(self.WORDS, dtype=np.uint32)
        if self.coords is not None:
            for k, c in enumerate(self.coords):
                out[12 * k:12 * k + 12] = _limbs(c)
        return out

    @classmethod
    def from_limbs(cls, xy, is_inf):
        if is_inf:
            return cls(None)
        return cls([_int(xy[12 * k:12 * k + 12]) for k in range(cls.WORDS // 12)])

    @classmethod
    def pack(cls, points):
        xy = np.zeros((len(points), cls.WORDS), dtype=np.uint32)
        inf = np.zeros(len(points), dtype=np.uint8)
        for i, p in enumerate(points):
            if p.coords is None:
                inf[i] = 1
            else:
                xy[i] = p.limbs()
        return xy, inf

    def __add__(self, o):
        if type(o) is not type(self):
            return NotImplemented
        xy, inf = self.pack([self, o])
        out, oinf = default_context().msm_oneshot(self.GROUP, xy, inf, scalars_to_array([1, 1]))
        return self.from_limbs(out, oinf)

    def __mul__(self, k):
        """raw integer multiple, not reduced mod r (macros.rs:10-21)"""
        k = int(getattr(k, "e", k))
        if k < 0:
            raise ValueError("negative scalar")
        if self.coords is None:
            return type(self)(None)
        if k >> 256:
            # the reference multiplies by any BigUint (macros.rs:10-21); the device takes 256-bit raw integers, so a
            # longer one goes in 256-bit chunks, Horner over 2^256 = 2 * 2^255
            chunks = []
            while k:
                chunks.append(k & ((1 << 256) - 1))
                k >>= 256
            acc = None
            for c in reversed(chunks):
                if acc is not None:
                    acc = (acc * (1 << 255)) * 2
                t = self * c
                acc = t if acc is None else acc + t
            return acc
        out, inf = default_context().mul_base(self.GROUP, self.limbs(), scalars_to_array([k]))
        return self.from_limbs(out[0], inf[0])

    __rmul__ = __mul__


class G1Point(_Point):
    GROUP, WORDS = 1, 24
    __slots__ = ()

    @classmethod
    def new(cls, x, y):
        return cls((x, y))

    @classmethod
    def g(cls):  # g1_point.rs:38-47
        return cls((0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
                    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1))

    @property
    def x(self):
        return self.coords[0]

    @property
    def y(self):
        return self.coords[1]

    def __neg__(self):  # g1_point.rs:178-195
        if self.coords is None:
            return G1Point(None)
        return G1Point((self.coords[0], (Q - self.coords[1]) % Q))

    def is_on_curve(self):  # g1_point.rs:101-107
        return self.coords is None or (self.y * self.y - self.x ** 3 - 4) % Q == 0

    def __repr__(self):
        return "G1Point(AtInfinity)" if self.coords is None else f"G1Point(x={self.x:#x}, y={self.y:#x})"


class G2Point(_Point):
    """coords = (x.u0, x.u1, y.u0, y.u1), the ABI order; new() takes the reference's (u1, u0) pairs."""
    GROUP, WORDS = 2, 48
    __slots__ = ()

    @classmethod
    def new(cls, x_u1, x_u0, y_u1, y_u0):  # Fq2::new(u1, u0), fq2.rs:22
        return cls((x_u0, x_u1, y_u0, y_u1))

    @classmethod
    def g(cls):  # g2_point.rs:36-46
        return cls.new(
            0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e,
            0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
            0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be,
            0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801)

    def __neg__(self):
        if self.coords is None:
            return G2Point(None)
        c = self.coords
        return G2Point((c[0], c[1], (Q - c[2]) % Q, (Q - c[3]) % Q))

    def __repr__(self):
        return "G2Point(AtInfinity)" if self.coords is None else f"G2Point{tuple(hex(c) for c in self.coords)}"


class _Points:
    """Device-resident slice of points (a CRS vector)."""
    POINT = None

    def __init__(self, points=None, precompute=False, ctx=None, _set=None, in_subgroup=False):
        """in_subgroup=True asserts that every point has order r (any CRS vector does: its points are multiples
        of the generator); scalars must then be field elements (< r) and the device may evaluate s P as
        (r - s)(-P).  Leave it False for arbitrary curve points / raw integer scalars (macros.rs:10-21)."""
        self.ctx = ctx or default_context()
        if _set is not None:
            self.set = _set
        else:
            xy, inf = self.POINT.pack(points)
            self.set = self.ctx.load_points(self.POINT.GROUP, xy, inf if inf.any() else None, precompute=precompute,
                                            in_subgroup=in_subgroup)

    @classmethod
    def from_arrays(cls, xy, inf=None, precompute=False, ctx=None, in_subgroup=False):
        ctx = ctx or default_context()
        return cls(ctx=ctx, _set=ctx.load_points(cls.POINT.GROUP, xy, inf, precompute=precompute, in_subgroup=in_subgroup))

    @classmethod
    def generator_multiples(cls, scalars, precompute=False, ctx=None, in_subgroup=True):
        """[k * g for k in scalars] computed on the device (CRS::new's calc_n_pows, crs.rs:88-104).
        Multiples of the generator have order r, so in_subgroup defaults to True here."""
        ctx = ctx or default_context()
        sc = scalars if isinstance(scalars, np.ndarray) else scalars_to_array(scalars)
        return cls(ctx=ctx, _set=ctx.points_from_scalars(cls.POINT.GROUP, cls.POINT.g().limbs(), sc, precompute=precompute,
                                                         in_subgroup=in_subgroup))

    def __len__(self):
        return len(self.set)

    def __getitem__(self, i):
        xy, inf = self.set.read(i, 1)
        return self.POINT.from_limbs(xy[0], inf[0])

    def to_list(self):
        xy, inf = self.set.read()
        return [self.POINT.from_limbs(xy[i], inf[i]) for i in range(len(self))]


class G1Points(_Points):
    POINT = G1Point


class G2Points(_Points):
    POINT = G2Point


class Polynomial:
    """field/polynomial.rs:32-36 restricted to what the MSM seam needs: coefficients in Fr,
    coeffs[i] multiplies x^i, trailing zeros trimmed (normalize, :139-152)."""

    def __init__(self, coeffs):
        if len(coeffs) == 0:
            raise ValueError("coeffs is empty")  # polynomial.rs:120
        c = [int(getattr(x, "e", x)) % R for x in coeffs]
        n = len(c)
        while n > 1 and c[n - 1] == 0:
            n -= 1
        self.coeffs = c[:n]

    def __len__(self):
        return len(self.coeffs)

    def _eval(self, powers, point_cls, points_cls):
        sc = scalars_to_array(self.coeffs)
        n = len(self.coeffs)
        if isinstance(powers, _Points):
            if powers.POINT is not point_cls:
                raise TypeError("wrong group")
            if n > len(powers):
                raise IndexError("index out of bounds: more coefficients than powers")  # polynomial.rs:278 panics
            out, inf = powers.ctx.msm(powers.set, sc)
            return point_cls.from_limbs(out, inf)
        if n > len(powers):
            raise IndexError("index out of bounds: more coefficients than powers")
        xy, pinf = point_cls.pack(list(powers[:n]))
        out, inf = default_context().msm_oneshot(point_cls.GROUP, xy, pinf if pinf.any() else None, sc)
        return point_cls.from_limbs(out, inf)

    def eval_with_g1_hidings(self, powers):  # polynomial.rs:272-281
        return self._eval(powers, G1Point, G1Points)

    def eval_with_g2_hidings(self, powers):  # polynomial.rs:284-293
        return self._eval(powers, G2Point, G2Points)
