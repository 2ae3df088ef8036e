"""Limb algorithms of csrc/{fp,fp2,ec}.cuh checked on the CPU.

The device headers are compiled with g++ (ptx.cuh emulates each PTX carry-chain
instruction), so Montgomery multiplication, the XYZZ formulas and every exceptional
case of the group law are compared with Python integers / the T0 oracle without a GPU.
The same comparisons run on the real kernels in test_gpu_*.py."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import zkt_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "host_emu")
Q, R = O.Q, O.R


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(EMU, "libemu_field.so")
    src = os.path.join(EMU, "emu_field.cpp")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def limbs(v, n):
    return (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(arr):
    return sum(int(x) << (32 * i) for i, x in enumerate(arr))


def field_op(lib, field, op, a, b=None):
    n, p = (12, Q) if field == 0 else (8, R)
    out = (ctypes.c_uint32 * n)()
    rc = lib.emu_field_op(field, op, limbs(a, n), limbs(b, n) if b is not None else None, out)
    assert rc == 0
    return val(out)


@pytest.mark.parametrize("field", [0, 1])
def test_mont_field_ops(lib, field):
    n, p = (12, Q) if field == 0 else (8, R)
    Rm = 1 << (32 * n)
    rnd = random.Random(1234 + field)
    specials = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, Rm % p, (1 << (32 * n - 33)) % p]
    vals = specials + [rnd.randrange(p) for _ in range(60)]
    for a in vals:
        am = field_op(lib, field, 5, a)                 # to_mont
        assert am == a * Rm % p
        assert field_op(lib, field, 6, am) == a         # from_mont
        assert field_op(lib, field, 3, a) == (-a) % p   # neg (form-agnostic)
    for _ in range(300):
        a, b = rnd.choice(vals), rnd.choice(vals)
        assert field_op(lib, field, 0, a, b) == (a + b) % p
        assert field_op(lib, field, 1, a, b) == (a - b) % p
        # Montgomery product: mont(a,b) = a*b/R
        assert field_op(lib, field, 2, a, b) == a * b * pow(Rm, -1, p) % p
    rinv = pow(Rm, -1, p)
    allones = [((1 << (32 * n)) - 1) % p, ((1 << (32 * n - 1)) - 1) % p, (0xFFFFFFFF << (32 * (n - 1))) % p]
    for a in vals + allones:
        assert field_op(lib, field, 8, a) == a * a * rinv % p            # dedicated squaring == mont(a, a)
    for _ in range(50):
        a, b = rnd.choice(vals + allones), rnd.choice(vals)
        assert field_op(lib, field, 9, a, b) == 2 * a * b * rinv % p     # lockstep pair
    for a in vals[1:40]:
        am = a * Rm % p
        assert field_op(lib, field, 4, am) == pow(a, -1, p) * Rm % p      # division steps in batches of 30
        assert field_op(lib, field, 10, am) == pow(a, -1, p) * Rm % p     # binary extended Euclid cross-check
    for a in vals[1:8]:
        am = a * Rm % p
        assert field_op(lib, field, 7, am) == pow(a, -1, p) * Rm % p      # Fermat cross-check
    assert field_op(lib, field, 4, 0) == 0


@pytest.mark.parametrize("field", [0, 1])
def test_inverse_many(lib, field):
    """finv (batched division steps, fp.cuh) against pow(a, -1, p): random values, small values, values near p and
    near powers of two (long runs of zero bits drive the variable-length inner loop), p - small."""
    n, p = (12, Q) if field == 0 else (8, R)
    Rm = 1 << (32 * n)
    rnd = random.Random(99 + field)
    vals = [rnd.randrange(1, p) for _ in range(1500)]
    vals += list(range(1, 40)) + [p - k for k in range(1, 40)]
    vals += [(1 << k) % p for k in range(0, 32 * n, 7)] + [((1 << k) - 1) % p for k in range(1, 32 * n, 11)]
    vals += [(p >> k) for k in range(1, 60)] + [(p + 1) // 2, (p - 1) // 2, (p - 1) // 3]
    for a in vals:
        if a == 0:
            continue
        # the function sees the Montgomery form aR; also feed a itself (then the result is (a/R)^-1 R = a^-1 R^2)
        assert field_op(lib, field, 4, a * Rm % p) == pow(a, -1, p) * Rm % p
        assert field_op(lib, field, 4, a) == pow(a, -1, p) * Rm * Rm % p


def test_fp2_ops(lib):
    rnd = random.Random(7)
    Rm = 1 << 384
    def enc(x):  # O.Fq2 -> 24 Montgomery limbs c0|c1
        arr = (ctypes.c_uint32 * 24)()
        for k, v in enumerate((x.u0.e, x.u1.e)):
            m = v * Rm % Q
            for i in range(12):
                arr[12 * k + i] = (m >> (32 * i)) & 0xFFFFFFFF
        return arr
    def dec(arr):
        c0 = val(arr[:12]) * pow(Rm, -1, Q) % Q
        c1 = val(arr[12:]) * pow(Rm, -1, Q) % Q
        return O.Fq2(O.Fq1(c1), O.Fq1(c0))
    for _ in range(40):
        a = O.Fq2(O.Fq1(rnd.randrange(Q)), O.Fq1(rnd.randrange(Q)))
        b = O.Fq2(O.Fq1(rnd.randrange(Q)), O.Fq1(rnd.randrange(Q)))
        for op, exp in ((0, a + b), (1, a - b), (2, a * b), (3, a.sq()), (4, a.inv())):
            out = (ctypes.c_uint32 * 24)()
            assert lib.emu_fp2_op(op, enc(a), enc(b), out) == 0
            assert dec(out) == exp, op


def g1_limbs(p):
    return (ctypes.c_uint32 * 24)(*O.g1_to_limbs(p))


def g1_from(arr):
    x, y = val(arr[:12]), val(arr[12:])
    return O.INF if (x == 0 and y == 0) else O.g1(x, y)


def g2_limbs(p):
    return (ctypes.c_uint32 * 48)(*O.g2_to_limbs(p))


def g2_from(arr):
    v = [val(arr[12 * k:12 * k + 12]) for k in range(4)]
    return O.INF if not any(v) else O.g2(v[1], v[0], v[3], v[2])


def check_group(lib, fn, to_l, from_l, nl, gen):
    ks = [1, 2, 3, 5, 12345, O.R - 1, O.R - 2, 0x1234567890ABCDEF]
    pts = [O.scalar_mul(gen, k) for k in ks] + [O.INF]
    for op in (0, 1):
        for p in pts:
            for q in pts + [O.point_neg(p)]:
                out = (ctypes.c_uint32 * nl)()
                assert fn(op, to_l(p), to_l(q), out) == 0
                assert from_l(out) == O.affine_add(p, q), (op, p, q)
    for p in pts:
        for op in (2, 3):
            out = (ctypes.c_uint32 * nl)()
            assert fn(op, to_l(p), None, out) == 0
            assert from_l(out) == O.affine_add(p, p)


def test_g1_group_law_all_cases(lib):
    check_group(lib, lib.emu_g1_op, g1_limbs, g1_from, 24, O.G1_GEN)


def test_g2_group_law_all_cases(lib):
    check_group(lib, lib.emu_g2_op, g2_limbs, g2_from, 48, O.G2_GEN)
