"""The C restatement of the reference algorithm (oracle/c/zkt_ref.c: the CPU baseline of bench.py)
against the reference's golden vectors and against the Python T0 oracle."""
import ctypes
import json
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import zkt_oracle as O
from tests import util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
u32p = ctypes.POINTER(ctypes.c_uint32)


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "c")], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libzkt_ref.so"))


def arr(v):
    return np.ascontiguousarray(v, dtype=np.uint32)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fq_op(lib, op, a, b=None):
    out = np.zeros(12, dtype=np.uint32)
    aa = arr(U.int_to_limbs(a, 12))
    bb = arr(U.int_to_limbs(b, 12)) if b is not None else None
    rc = lib.zkt_fq_op_ref(op, P(aa), P(bb) if bb is not None else None, P(out))
    return rc, U.limbs_to_int(out)


def test_fq_ops_vs_python(lib):
    rnd = random.Random(3)
    vals = [0, 1, 2, O.Q - 1, O.Q - 2, (O.Q - 1) // 2] + [rnd.randrange(O.Q) for _ in range(200)]
    for _ in range(500):
        a, b = rnd.choice(vals), rnd.choice(vals)
        assert fq_op(lib, 0, a, b) == (0, (a + b) % O.Q)
        assert fq_op(lib, 1, a, b) == (0, (a - b) % O.Q)
        assert fq_op(lib, 2, a, b) == (0, (a * b) % O.Q)
    for a in vals:
        assert fq_op(lib, 3, a) == (0, (-a) % O.Q)
        if a:
            assert fq_op(lib, 4, a) == (0, O.ext_euclid_inv(a, O.Q))
    assert fq_op(lib, 4, 0)[0] == -1                      # "Cannot find inverse of zero"


def g1_call(lib, fn, p, *rest):
    xy, inf = U.g1_points_to_array([p])
    return xy[0], int(inf[0])


def g1_add(lib, p, q):
    (a, ai), (b, bi) = g1_call(lib, None, p), g1_call(lib, None, q)
    out = np.zeros(24, dtype=np.uint32)
    oi = ctypes.c_int(0)
    lib.zkt_g1_add_ref(P(a), ai, P(b), bi, P(out), ctypes.byref(oi))
    return U.g1_from_array(out, oi.value)


def g1_mul(lib, p, k):
    a, ai = g1_call(lib, None, p)
    kk = arr(U.int_to_limbs(k, 8))
    out = np.zeros(24, dtype=np.uint32)
    oi = ctypes.c_int(0)
    lib.zkt_g1_mul_ref(P(a), ai, P(kk), P(out), ctypes.byref(oi))
    return U.g1_from_array(out, oi.value)


def test_g1_golden_vectors(lib):
    k = json.load(open(os.path.join(G, "ref_g1.json")))
    g = O.G1_GEN
    gs = [O.g1(int(p["x"]), int(p["y"])) for p in k["g_multiples"]["points"]]
    assert g1_add(lib, g, g) == O.g1(int(k["add_same_point"]["x"]), int(k["add_same_point"]["y"]))
    for n in range(1, 11):
        assert g1_mul(lib, g, n) == gs[n - 1]
    for c in k["scalar_mul_gen_pubkey"]["cases"]:
        assert g1_mul(lib, g, int(c["multiple"]) % O.Q) == O.g1(int(c["x"]), int(c["y"]))
    for a, b, c in k["add_different_points"]["cases"]:
        assert g1_add(lib, gs[a - 1], gs[b - 1]) == gs[c - 1]
    assert g1_add(lib, g, O.point_neg(g)) is O.INF
    assert g1_add(lib, g, O.INF) == g and g1_add(lib, O.INF, g) == g and g1_add(lib, O.INF, O.INF) is O.INF
    assert g1_mul(lib, g, 0) is O.INF and g1_mul(lib, g, O.R) is O.INF


def test_g2_golden_vectors(lib):
    k = json.load(open(os.path.join(G, "ref_g2.json")))
    mk = lambda p: O.g2(int(p["x1"]), int(p["x0"]), int(p["y1"]), int(p["y0"]))
    gs = [mk(p) for p in k["g_multiples"]["points"]]

    def call(fn, p, q=None, kk=None):
        xy, inf = U.g2_points_to_array([p] + ([q] if q is not None else []))
        out = np.zeros(48, dtype=np.uint32)
        oi = ctypes.c_int(0)
        if q is not None:
            lib.zkt_g2_add_ref(P(xy[0]), int(inf[0]), P(xy[1]), int(inf[1]), P(out), ctypes.byref(oi))
        else:
            s = arr(U.int_to_limbs(kk, 8))
            lib.zkt_g2_mul_ref(P(xy[0]), int(inf[0]), P(s), P(out), ctypes.byref(oi))
        return U.g2_from_array(out, oi.value)

    g = O.G2_GEN
    assert call(None, g, g) == mk(k["add_same_point"])
    for n in range(1, 11):
        assert call(None, g, kk=n) == gs[n - 1]
    for c in k["scalar_mul_gen_pubkey"]["cases"]:
        assert call(None, g, kk=int(c["multiple"])) == mk(c)
    for a, b, c in k["add_different_points"]["cases"]:
        assert call(None, gs[a - 1], gs[b - 1]) == gs[c - 1]
    assert call(None, g, O.point_neg(g)) is O.INF


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_msm_ref_vs_python_oracle(lib, threads):
    rnd = random.Random(11)
    n = 9
    pts = [O.scalar_mul(O.G1_GEN, rnd.randrange(1, O.R)) for _ in range(n)] + [O.INF]
    sc = [rnd.randrange(O.R) for _ in range(n)] + [5]
    sc[2] = 0
    xy, inf = U.g1_points_to_array(pts)
    s = U.scalars_to_array(sc)
    out = np.zeros(24, dtype=np.uint32)
    oi = ctypes.c_int(0)
    lib.zkt_g1_msm_ref(P(xy), P(inf), P(s), len(sc), threads, P(out), ctypes.byref(oi))
    assert U.g1_from_array(out, oi.value) == O.msm(pts, sc)
    lib.zkt_g1_msm_ref(P(xy), P(inf), P(s), 0, threads, P(out), ctypes.byref(oi))
    assert oi.value == 1
    pts2 = [O.scalar_mul(O.G2_GEN, rnd.randrange(1, O.R)) for _ in range(3)]
    xy2, inf2 = U.g2_points_to_array(pts2)
    out2 = np.zeros(48, dtype=np.uint32)
    lib.zkt_g2_msm_ref(P(xy2), None, P(s), 3, threads, P(out2), ctypes.byref(oi))
    assert U.g2_from_array(out2, oi.value) == O.msm(pts2, sc[:3])
