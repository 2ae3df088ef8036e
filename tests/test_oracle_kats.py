"""Pin the T0 oracle against the reference's own known-answer tests.

Vectors: tests/golden/ref_*.json, extracted verbatim from the reference's #[test]
functions by tests/golden/extract_kats.py (file:line recorded in each JSON)."""
import json
import os

import pytest

from oracle import zkt_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, f"ref_{name}.json")) as f:
        return json.load(f)


def test_params_and_generators():
    p = load("params")
    assert int(p["q"], 16) == O.Q and int(p["r"], 16) == O.R
    assert (int(p["g1"][0], 16), int(p["g1"][1], 16)) == (O.G1_GEN[0].e, O.G1_GEN[1].e)
    x1, x0, y1, y0 = (int(v, 16) for v in p["g2_x1_x0_y1_y0"])
    assert O.G2_GEN == O.g2(x1, x0, y1, y0)
    assert O.g1_is_on_curve(O.G1_GEN) and O.g2_is_on_curve(O.G2_GEN)


# ---- G1: g1_point.rs:203-412
def g1_multiples():
    return [O.g1(int(p["x"]), int(p["y"])) for p in load("g1")["g_multiples"]["points"]]


def test_g1_add_same_point():
    k = load("g1")["add_same_point"]
    assert O.affine_add(O.G1_GEN, O.G1_GEN) == O.g1(int(k["x"]), int(k["y"]))


def test_g1_edge_cases():
    g = O.G1_GEN
    assert O.affine_add(g, O.point_neg(g)) is O.INF          # negate, add_vertical_line
    assert O.affine_add(g, O.INF) == g and O.affine_add(O.INF, g) == g
    assert O.affine_add(O.INF, O.INF) is O.INF
    assert O.scalar_mul(g, 1) == g
    assert O.scalar_mul(g, 2) == O.affine_add(g, g)
    assert O.scalar_mul(g, 3) == O.affine_add(O.affine_add(g, g), g)
    assert O.scalar_mul(g, 0) is O.INF
    assert O.scalar_mul(g, O.R) is O.INF


def test_g1_scalar_mul_smaller_nums():
    gs = g1_multiples()
    for n in range(1, 11):
        assert O.scalar_mul(O.G1_GEN, n) == gs[n - 1]


def test_g1_scalar_mul_gen_pubkey():
    for c in load("g1")["scalar_mul_gen_pubkey"]["cases"]:
        # the reference wraps the multiple in the BASE field (g1_point.rs:361), i.e. mod q, not mod r
        k = int(c["multiple"]) % O.Q
        assert O.scalar_mul(O.G1_GEN, k) == O.g1(int(c["x"]), int(c["y"]))


def test_g1_add_different_points():
    gs = g1_multiples()
    for a, b, c in load("g1")["add_different_points"]["cases"]:
        assert O.affine_add(gs[a - 1], gs[b - 1]) == gs[c - 1]


# ---- G2: g2_point.rs:178-444
def g2_pt(p):
    return O.g2(int(p["x1"]), int(p["x0"]), int(p["y1"]), int(p["y0"]))


def g2_multiples():
    return [g2_pt(p) for p in load("g2")["g_multiples"]["points"]]


def test_g2_add_same_point():
    assert O.affine_add(O.G2_GEN, O.G2_GEN) == g2_pt(load("g2")["add_same_point"])


def test_g2_edge_cases():
    g = O.G2_GEN
    assert O.affine_add(g, O.point_neg(g)) is O.INF
    assert O.affine_add(g, O.INF) == g and O.affine_add(O.INF, g) == g
    assert O.affine_add(O.INF, O.INF) is O.INF
    assert O.scalar_mul(g, 2) == O.affine_add(g, g)
    assert O.scalar_mul(g, O.R) is O.INF


def test_g2_scalar_mul_smaller_nums():
    gs = g2_multiples()
    for n in range(1, 11):
        assert O.scalar_mul(O.G2_GEN, n) == gs[n - 1]


def test_g2_scalar_mul_gen_pubkey():
    for c in load("g2")["scalar_mul_gen_pubkey"]["cases"]:
        assert O.scalar_mul(O.G2_GEN, int(c["multiple"]) % O.R) == g2_pt(c)


def test_g2_add_different_points():
    gs = g2_multiples()
    for a, b, c in load("g2")["add_different_points"]["cases"]:
        assert O.affine_add(gs[a - 1], gs[b - 1]) == gs[c - 1]


# ---- tower: fq_test_helper.rs:9-34 builds operands from -3, -5, -7, -9
def fq1_values():
    return tuple(-O.Fq1(v) for v in (3, 5, 7, 9))


def fq2_values():
    a1, b1, c1, d1 = fq1_values()
    return O.Fq2(a1, b1), O.Fq2(b1, c1), O.Fq2(c1, d1), O.Fq2(d1, a1)


def fq6_values():
    a2, b2, c2, d2 = fq2_values()
    return O.Fq6(a2, b2, c2), O.Fq6(b2, c2, d2), O.Fq6(c2, d2, a2), O.Fq6(d2, a2, b2)


def s2(x):
    return [str(x.u1.e), str(x.u0.e)]


def s6(x):
    return s2(x.v2) + s2(x.v1) + s2(x.v0)


def s12(x):
    return s6(x.w1) + s6(x.w0)


def test_fq2_kats():  # fq2.rs:165-235
    k = load("fq2")
    a1, b1, c1, d1 = fq1_values()
    a2, b2 = O.Fq2(a1, b1), O.Fq2(c1, d1)
    assert s2(a2 + b2) == k["test_add"]
    assert s2(a2 - b2) == k["test_sub"]
    assert s2(a2 * b2) == k["test_mul"]
    assert s2(a2.inv()) + s2(b2.inv()) == k["test_inv"]
    assert s2((a2 * b2).reduce()) == k["test_reduce"]
    for v in fq2_values():
        assert (-v) + v == O.Fq2.zero()


def test_fq6_kats():  # fq6.rs:189-275
    k = load("fq6")
    a2, b2, c2, d2 = fq2_values()
    a6, b6 = O.Fq6(a2, b2, c2), O.Fq6(b2, c2, d2)
    assert s6(a6 + b6) == k["test_add"]
    assert s6(a6 - b6) == k["test_sub"]
    assert s6(a6 * b6) == k["test_mul"]
    assert s6(a6.inv()) + s6(b6.inv()) == k["test_inv"]
    assert s6((a6 * b6).reduce()) == k["test_reduce"]


def test_fq12_kats():  # fq12.rs:197-329
    k = load("fq12")
    a6, b6, c6, d6 = fq6_values()
    a12, b12 = O.Fq12(a6, b6), O.Fq12(c6, d6)
    assert s12(a12 + b12) == k["test_add"]
    assert s12(a12 - b12) == k["test_sub"]
    assert s12(a12 * b12) == k["test_mul"]
    assert s12(a12.inv()) + s12(b12.inv()) == k["test_inv"]
    assert O.Fq12.from_int(3).pow(4) == O.Fq12.from_int(81)


def test_fp_inverse_tables():  # prime_field_elem.rs:625-821 (exhaustive small-prime inverses)
    for p in (97, 53, 11):
        for v in range(1, p):
            assert (O.ext_euclid_inv(v, p) * v) % p == 1
    with pytest.raises(ZeroDivisionError):
        O.ext_euclid_inv(0, 97)


# ---- MSM seam: polynomial.rs:1250-1285 (self-consistency, toy scalars 2..5)
def test_eval_with_hidings_matches_written_out_sum():
    g = O.G1_GEN
    pts = [O.scalar_mul(g, k) for k in (1, 2, 3, 4)]
    poly = O.Polynomial([2, 3, 4, 5])
    exp = O.INF
    for p, s in zip(pts, (2, 3, 4, 5)):
        exp = O.affine_add(exp, O.scalar_mul(p, s))
    assert poly.eval_with_g1_hidings(pts) == exp == O.scalar_mul(g, 2 + 6 + 12 + 20)
    h = O.G2_GEN
    pts2 = [O.scalar_mul(h, k) for k in (1, 2, 3, 4)]
    assert poly.eval_with_g2_hidings(pts2) == O.scalar_mul(h, 40)
    assert O.msm(pts, []) is O.INF
    with pytest.raises(IndexError):
        O.msm(pts[:2], [1, 2, 3])


# ---- pairing: pairing.rs:173-196
@pytest.mark.slow
def test_tate_bilinear_generators():
    p1, p2 = O.G1_GEN, O.G2_GEN
    ten = O.scalar_mul(p1, 10)
    assert O.tate(O.affine_add(p1, ten), p2) == O.tate(p1, p2) * O.tate(ten, p2)


@pytest.mark.slow
def test_tate_plus_to_mul():
    p = O.affine_add(O.G1_GEN, O.G1_GEN)
    a = O.tate(p, O.G2_GEN)
    assert O.tate(O.affine_add(p, p), O.G2_GEN) == a * a


# ---- Groth16 end to end: groth16/zktoolkit_based/prover.rs:158-192
@pytest.mark.slow
def test_groth16_config1_prove_verify():
    prover = O.Prover(**O.CONFIG1)
    assert (prover.n, prover.l, prover.m) == (5, 2, 6)
    assert len(prover.h) - 1 == 3
    crs = O.CRS(prover, alpha=0x1111, beta=0x2222_3333, gamma=0x4444_5555_6666,
                delta=0x7777_8888_9999_aaaa, x=0xbbbb_cccc_dddd_eeee_ffff)
    proof = prover.prove(crs, r=0x1234_5678_9abc, s=0xfedc_ba98_7654)
    assert O.verify(proof, crs, prover.statement())
    bad = (O.affine_add(proof[0], O.G1_GEN), proof[1], proof[2])
    assert not O.verify(bad, crs, prover.statement())


# ---- Pinocchio end to end: pinocchio/prover.rs:178-211
@pytest.mark.slow
def test_pinocchio_config1_prove_verify():
    op = O.PinocchioProver(**O.CONFIG1)
    assert op.max_degree == 9 and op.mid() == [9, 27, 8, 35] and op.io() == [1, 3, 35]
    crs = O.PinocchioCRS(op, r_v=0x1357, r_w=0x2468ace, alpha_v=0x1111, alpha_w=0x2222, alpha_y=0x3333, beta=0x4444,
                         gamma=0x5555, s=0x66667777)
    proof = O.pinocchio_prove(op, crs, 0xabcdef, 0x123457)
    assert O.pinocchio_verify(proof, crs, op.io())
    bad = dict(proof, y_mid_s=O.affine_add(proof["y_mid_s"], O.G1_GEN))
    assert not O.pinocchio_verify(bad, crs, op.io())
