"""Synthetic Groth16 instance in aggregated-coefficient space (BASELINE.json configs[4]).

The reference's own pipeline cannot produce a 2^18-constraint instance: `QAP::build` is O(n^3) over the
integer domain 1..n (qap/qap.rs:33-97).  SURVEY.md section 7 gives a construction the reference verifier
(verifier.rs:36-53) still accepts; it only sees CRS points, so the proof verifies iff all of the prover's MSMs
are right:

  * random trapdoor (alpha, beta, gamma, delta, x), random coefficient vectors u, v (degree < n) and
    h (degree n-2), random wires a_0 = 1 (the statement) and a_1..a_m (the witness);
  * per-wire evaluations (U_i, V_i, W_i) random for all wires but the last, whose values are solved from
    sum a_i U_i = u(x), sum a_i V_i = v(x), sum a_i W_i = u(x) v(x) - h(x) t(x),  t(x) = prod_{k=1..n} (x - k);
  * CRS points exactly as crs.rs:65-135 builds them from those scalars, computed on the device by the
    vector fixed-base multiplication (`&G1Point * &Fq1` for many scalars, macros.rs:2-32).
All scalar arithmetic is exact (Python integers mod r)."""
import random

import numpy as np

from .api import G1Point, G1Points, G2Point, G2Points, R, default_context, scalars_to_array
from .groth16 import DeviceCRS, Prover


def _inv(a):
    return pow(a % R, -1, R)


def _eval(coeffs, xpows):
    return sum(c * p for c, p in zip(coeffs, xpows)) % R


def build(n, m_wit, seed=0x5EED0004, precompute=True, ctx=None):
    """Returns dict(prover, crs (DeviceCRS), trapdoor, stmt_wires, uvw_stmt (list[G1Point]), g2_gamma, g2_delta,
    scalars of A/B/C for the closed-form check)."""
    ctx = ctx or default_context()
    rnd = random.Random(seed)
    nz = lambda: rnd.randrange(1, R)
    alpha, beta, gamma, delta, x = nz(), nz(), nz(), nz(), nz()
    u = [rnd.randrange(R) for _ in range(n)]
    v = [rnd.randrange(R) for _ in range(n)]
    h = [rnd.randrange(R) for _ in range(n - 1)]
    xp = [1] * n
    for j in range(1, n):
        xp[j] = xp[j - 1] * x % R
    t = 1
    for k in range(1, n + 1):
        t = t * ((x - k) % R) % R
    ux, vx, hx = _eval(u, xp), _eval(v, xp), _eval(h, xp)
    wx = (ux * vx - hx * t) % R
    m = m_wit                      # wires 0..m, wire 0 is the statement "1"
    a = [1] + [nz() for _ in range(m)]
    U = [rnd.randrange(R) for _ in range(m)]
    V = [rnd.randrange(R) for _ in range(m)]
    Wv = [rnd.randrange(R) for _ in range(m)]
    am_inv = _inv(a[m])
    solve = lambda target, vals: (target - sum(ai * vi for ai, vi in zip(a[:m], vals))) % R * am_inv % R
    U.append(solve(ux, U)); V.append(solve(vx, V)); Wv.append(solve(wx, Wv))
    ginv, dinv = _inv(gamma), _inv(delta)
    comb = lambda i: (beta * U[i] + alpha * V[i] + Wv[i]) % R
    stmt_scalars = [comb(0) * ginv % R]
    wit_scalars = [comb(i) * dinv % R for i in range(1, m + 1)]
    xt_scalars = [xp[j] * t % R * dinv % R for j in range(n)]

    g1 = lambda ks: G1Points.generator_multiples(scalars_to_array(ks), ctx=ctx)
    g2 = lambda ks: G2Points.generator_multiples(scalars_to_array(ks), ctx=ctx)
    singles1 = g1([alpha, beta, delta] + stmt_scalars).set.read()[0]
    singles2 = g2([beta, gamma, delta]).set.read()[0]
    arrs = {
        "g1_xi": g1(xp).set.read()[0], "g1_uvw_wit": g1(wit_scalars).set.read()[0],
        "g1_xt_by_delta": g1(xt_scalars).set.read()[0][: n - 1], "g2_xi": g2(xp).set.read()[0],
        "g1_alpha": singles1[0], "g1_beta": singles1[1], "g1_delta": singles1[2],
        "g2_beta": singles2[0], "g2_delta": singles2[2],
    }
    crs = DeviceCRS.from_arrays(arrs, precompute=precompute, ctx=ctx)
    prover = Prover(u, v, h, a[1:])
    return {
        "prover": prover, "crs": crs, "n": n, "m_wit": m,
        "trapdoor": dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, x=x),
        "stmt_wires": [1], "uvw_stmt": [G1Point.from_limbs(singles1[3], False)],
        "g1_alpha": G1Point.from_limbs(singles1[0], False), "g2_beta": G2Point.from_limbs(singles2[0], False),
        "g2_gamma": G2Point.from_limbs(singles2[1], False), "g2_delta": G2Point.from_limbs(singles2[2], False),
        "evals": dict(ux=ux, vx=vx, hx=hx, t=t, wit_sum=sum(ai * si for ai, si in zip(a[1:], wit_scalars)) % R),
    }


def expected_dlogs(inst, r, s):
    """discrete logs of the proof elements w.r.t. the generators (closed form from the trapdoor)"""
    td, ev = inst["trapdoor"], inst["evals"]
    A = (td["alpha"] + ev["ux"] + r * td["delta"]) % R
    B = (td["beta"] + ev["vx"] + s * td["delta"]) % R
    C = (ev["wit_sum"] + ev["hx"] * ev["t"] % R * _inv(td["delta"]) + A * s + B * r - r * s % R * td["delta"]) % R
    return A, B, C
