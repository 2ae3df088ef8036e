"""Groth16 prove on the GPU (zk-toolkit_b200/groth16.py) against the oracle's restatement of
prover.rs:96-147 and the reference verifier equation (verifier.rs:36-53)."""
import random

import pytest

from oracle import zkt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def z():
    import zk_toolkit_b200 as z
    return z


def g1_to_api(z, p):
    return z.G1Point.zero() if p is O.INF else z.G1Point.new(p[0].e, p[1].e)


def g2_to_api(z, p):
    return z.G2Point.zero() if p is O.INF else z.G2Point.new(p[0].u1.e, p[0].u0.e, p[1].u1.e, p[1].u0.e)


def g1_to_o(p):
    return O.INF if p.is_zero() else O.g1(p.x, p.y)


def g2_to_o(p):
    if p.is_zero():
        return O.INF
    c = p.coords
    return O.g2(c[1], c[0], c[3], c[2])


@pytest.mark.parametrize("precompute", [False, True])
def test_config1_proof_identical_to_oracle_and_verifies(z, precompute):
    """the reference's only proof: (x*x*x)+x+5 == 35 with witness x = 3 (prover.rs:158-192)"""
    from importlib import import_module
    G = import_module("zk-toolkit_b200.groth16")
    op = O.Prover(**O.CONFIG1)
    crs = O.CRS(op, alpha=0x1111, beta=0x22223333, gamma=0x444455556666, delta=0x777788889999aaaa,
                x=0xbbbbccccddddeeeeffff)
    r, s = 0x123456789abc, 0xfedcba987654
    want = op.prove(crs, r, s)
    dcrs = G.DeviceCRS(g1_to_api(z, crs.g1_alpha), g1_to_api(z, crs.g1_beta), g1_to_api(z, crs.g1_delta),
                       [g1_to_api(z, p) for p in crs.g1_xi], [g1_to_api(z, p) for p in crs.g1_uvw_wit],
                       [g1_to_api(z, p) for p in crs.g1_xt_by_delta], g2_to_api(z, crs.g2_beta),
                       g2_to_api(z, crs.g2_delta), [g2_to_api(z, p) for p in crs.g2_xi], precompute=precompute)
    gp = G.Prover.from_per_wire([p.coeffs for p in op.ui], [p.coeffs for p in op.vi], op.h.coeffs, op.wires, op.l)
    proof = gp.prove(dcrs, r, s)
    # the same with the witness aggregation AND the quotient polynomial h computed on the device
    gq = G.Prover.from_qap([p.coeffs for p in op.ui], [p.coeffs for p in op.vi], [p.coeffs for p in op.wi], op.wires, op.l, op.n)
    assert gq.h[:len(op.h.coeffs)] == op.h.coeffs
    proof_q = gq.prove(dcrs, r, s)
    assert (proof_q.A, proof_q.B, proof_q.C) == (proof.A, proof.B, proof.C)
    got = (g1_to_o(proof.A), g2_to_o(proof.B), g1_to_o(proof.C))
    assert got == want                                      # identical canonical affine coordinates
    assert O.verify(got, crs, op.statement())               # and the reference verifier equation holds
    # the proof split over 1, 2 and 4 "devices" (zkmsm_groth16_prove_partial / _combine, all on this GPU here)
    if precompute:
        for world in (1, 2, 4):
            parts = [gp.prove_partial(dcrs, r, s, k, world) for k in range(world)]
            pd = G.combine_partials(parts)
            assert (pd.A, pd.B, pd.C) == (proof.A, proof.B, proof.C)
    # r, s are field elements: anything else is rejected, not reduced silently on the device
    import numpy as np
    ctx = dcrs.ctx
    u, v, h, wit = gp._arrays(dcrs)
    bad_r = z.scalars_to_array([O.R])
    ok_s = z.scalars_to_array([5])
    out = np.zeros(96, dtype=np.uint32)
    import ctypes
    inf = (ctypes.c_int * 3)()
    from zk_toolkit_b200 import ZkmsmError  # noqa: F401
    L = import_module("zk-toolkit_b200._lib")
    rc = ctx.lib.zkmsm_groth16_prove(ctx.h, dcrs.handle, L.dptr(u), L.dptr(v), L.dptr(h), L.dptr(wit), L.dptr(bad_r), L.dptr(ok_s),
                                     L.dptr(out), inf)
    assert rc == -3


def test_synthetic_instance_closed_form_and_verifier(z):
    """configs[4] construction at n = 2^10: proof elements equal their closed-form discrete logs times the
    generators, and the pairing check of verifier.rs:36-53 passes."""
    from importlib import import_module
    S = import_module("zk-toolkit_b200.synthetic")
    inst = S.build(1 << 10, 1 << 10, seed=77)
    rnd = random.Random(5)
    r, s = rnd.randrange(1, O.R), rnd.randrange(1, O.R)
    proof = inst["prover"].prove(inst["crs"], r, s)
    a, b, c = S.expected_dlogs(inst, r, s)
    assert g1_to_o(proof.A) == O.scalar_mul(O.G1_GEN, a)
    assert g2_to_o(proof.B) == O.scalar_mul(O.G2_GEN, b)
    assert g1_to_o(proof.C) == O.scalar_mul(O.G1_GEN, c)

    class _CRS:  # the fields Verifier::verify reads
        pass
    crs = _CRS()
    crs.g1_uvw_stmt = [g1_to_o(p) for p in inst["uvw_stmt"]]
    crs.g2_gamma, crs.g2_delta = g2_to_o(inst["g2_gamma"]), g2_to_o(inst["g2_delta"])
    crs.gt_alpha_beta = O.tate(g1_to_o(inst["g1_alpha"]), g2_to_o(inst["g2_beta"]))
    got = (g1_to_o(proof.A), g2_to_o(proof.B), g1_to_o(proof.C))
    assert O.verify(got, crs, inst["stmt_wires"])
    bad = (got[0], got[1], O.affine_add(got[2], O.G1_GEN))
    assert not O.verify(bad, crs, inst["stmt_wires"])


def test_fr_aggregate_matches_python(z):
    import numpy as np
    ctx = z.default_context()
    rnd = random.Random(3)
    for n_wires, n in ((1, 1), (7, 5), (200, 1000)):
        polys = [[rnd.randrange(O.R) for _ in range(n)] for _ in range(n_wires)]
        wires = [rnd.randrange(O.R) for _ in range(n_wires)]
        mat = np.stack([z.scalars_to_array(p) for p in polys])
        out = ctx.fr_aggregate(mat, z.scalars_to_array(wires))
        exp = [sum(a * p[j] for a, p in zip(wires, polys)) % O.R for j in range(n)]
        assert [sum(int(w) << (32 * k) for k, w in enumerate(r)) for r in out] == exp


def test_fr_quotient_matches_oracle(z):
    """h = (u v - w) / t on the device vs the oracle's Prover::new (prover.rs:64-71)"""
    from importlib import import_module
    G = import_module("zk-toolkit_b200.groth16")
    op = O.Prover(**O.CONFIG1)
    u = G.aggregate([p.coeffs for p in op.ui], op.wires)
    v = G.aggregate([p.coeffs for p in op.vi], op.wires)
    w = G.aggregate([p.coeffs for p in op.wi], op.wires)
    h = G.quotient(u, v, w, op.n)
    assert h[:len(op.h.coeffs)] == op.h.coeffs and not any(h[len(op.h.coeffs):])
    w[0] = (w[0] + 1) % O.R
    with pytest.raises(ValueError):
        G.quotient(u, v, w, op.n)
    rnd = random.Random(23)
    n = 300
    t = O.qap_build_t(n)
    uu = [rnd.randrange(O.R) for _ in range(n)]
    vv = [rnd.randrange(O.R) for _ in range(n)]
    q, rem = O.Polynomial(uu, normalize=False).multiply_by(O.Polynomial(vv, normalize=False)).divide_by(t)
    remc = ((rem.coeffs if rem is not None else [0]) + [0] * n)[:n]
    assert G.quotient(uu, vv, remc, n) == (q.coeffs + [0] * n)[:n - 1]


def test_fr_quotient_transforms_vs_schoolbook_and_identity(z, monkeypatch):
    """the transform path of zkmsm_fr_quotient (n >= 32) against the reference-style schoolbook path forced by
    ZKMSM_QUOTIENT_SCHOOLBOOK, and the defining identity u v - w == h t checked with Python integers at n = 1024"""
    ctx = z.default_context()
    rnd = random.Random(77)
    for n in (32, 33, 1000, 2048):
        u, v, w = ([rnd.randrange(O.R) for _ in range(n)] for _ in range(3))
        arrs = [z.scalars_to_array(x) for x in (u, v, w)]
        h1, exact1 = ctx.fr_quotient(*arrs)
        ctx.set_option("quotient_schoolbook", 1)
        h2, exact2 = ctx.fr_quotient(*arrs)
        ctx.set_option("quotient_schoolbook", 0)
        assert h1.tolist() == h2.tolist() and exact1 == exact2 == False   # random w: Euclidean quotient + remainder
    n = 1024
    t = O.qap_build_t(n).coeffs
    u, v = ([rnd.randrange(O.R) for _ in range(n)] for _ in range(2))
    h0, _ = ctx.fr_quotient(z.scalars_to_array(u), z.scalars_to_array(v), z.scalars_to_array([0] * n))
    h0 = [sum(int(x) << (32 * k) for k, x in enumerate(r)) for r in h0]
    p = [0] * (2 * n - 1)
    for i, a in enumerate(u):
        for j, b in enumerate(v):
            p[i + j] += a * b
    for i, a in enumerate(h0):
        for j, b in enumerate(t):
            p[i + j] -= a * b
    p = [x % O.R for x in p]
    assert not any(p[n:])                                        # u v - h t has degree < n: h0 is the quotient
    h, exact = ctx.fr_quotient(z.scalars_to_array(u), z.scalars_to_array(v), z.scalars_to_array(p[:n]))
    assert exact and [sum(int(x) << (32 * k) for k, x in enumerate(r)) for r in h] == h0


def test_pinocchio_proof_identical_to_oracle_and_verifies(z):
    """pinocchio/prover.rs:178-211: the same circuit through Pinocchio; all nine proof elements must equal the
    oracle's restatement and pass the restated verifier (verifier.rs:27-87)."""
    from importlib import import_module
    P = import_module("zk-toolkit_b200.pinocchio")
    G = import_module("zk-toolkit_b200.groth16")
    op = O.PinocchioProver(**O.CONFIG1)
    crs = O.PinocchioCRS(op, r_v=0x1357, r_w=0x2468ace, alpha_v=0x1111, alpha_w=0x2222, alpha_y=0x3333, beta=0x4444,
                         gamma=0x5555, s=0x66667777)
    dv, dy = 0xabcdef, 0x123457
    want = O.pinocchio_prove(op, crs, dv, dy)
    keys = P.DeviceKeys(crs, lambda p: g1_to_api(z, p), lambda p: g2_to_api(z, p))
    agg = lambda polys: G.aggregate([p.coeffs for p in polys], op.witness)
    h = G.quotient(agg(op.vi), agg(op.wi), agg(op.yi), op.num_constraints)
    got = P.prove(keys, op.mid(), op.io(), h, dv, dy)
    conv = {k: (g2_to_o(v) if isinstance(v, z.G2Point) else g1_to_o(v)) for k, v in got.items()}
    assert conv == want
    assert O.pinocchio_verify(conv, crs, op.io())


def test_crs_new_matches_oracle_crs_and_proves(z):
    """CRS::new (crs.rs:49-146) on the device for config 1: every CRS vector equals the oracle's, and the resident CRS
    built from it gives the oracle's proof"""
    from importlib import import_module
    G = import_module("zk-toolkit_b200.groth16")
    op = O.Prover(**O.CONFIG1)
    td = dict(alpha=0x1111, beta=0x22223333, gamma=0x444455556666, delta=0x777788889999aaaa, x=0xbbbbccccddddeeeeffff)
    ocrs = O.CRS(op, with_pairing=False, **td)
    crs = G.CRS([p.coeffs for p in op.ui], [p.coeffs for p in op.vi], [p.coeffs for p in op.wi], op.l, op.n, **td)
    for name in ("g1_xi", "g1_uvw_stmt", "g1_uvw_wit", "g1_xt_by_delta"):
        want = [O.g1_to_limbs(p) for p in getattr(ocrs, name)]
        assert getattr(crs, name).tolist() == want, name
        assert not getattr(crs, name + "_inf").any()
    assert crs.g2_xi.tolist() == [O.g2_to_limbs(p) for p in ocrs.g2_xi]
    for name in ("g1_alpha", "g1_beta", "g1_delta"):
        assert getattr(crs, name).tolist() == O.g1_to_limbs(getattr(ocrs, name)), name
    for name in ("g2_beta", "g2_gamma", "g2_delta"):
        assert getattr(crs, name).tolist() == O.g2_to_limbs(getattr(ocrs, name)), name
    r, s = 0x5555, 0x7777
    gp = G.Prover.from_qap([p.coeffs for p in op.ui], [p.coeffs for p in op.vi], [p.coeffs for p in op.wi], op.wires, op.l, op.n)
    proof = gp.prove(crs.device(), r, s)
    assert (g1_to_o(proof.A), g2_to_o(proof.B), g1_to_o(proof.C)) == op.prove(ocrs, r, s)
