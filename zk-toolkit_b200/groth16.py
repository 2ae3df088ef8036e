"""Groth16 proof generation on the GPU: the five MSMs of `Prover::prove`
(src/zk/w_trusted_setup/groth16/zktoolkit_based/prover.rs:96-147).

The reference evaluates, per wire i, `ui[i].eval_with_g1_hidings(xi) * a_i` and sums the results
(prover.rs:108-117): (m+1) nested MSMs.  Because the group is commutative that equals ONE MSM with the
aggregated coefficients  sum_i a_i * u_{i,j}  (the canonical affine result is
identical).  The blinding terms are folded into the same MSMs by appending the single CRS points to
the resident sets:

    A    = alpha   + sum_j u_j [x^j]_1 + r delta        = MSM(xi_1 ++ [alpha_1, delta_1], u ++ [1, r])
    B    = beta_2  + sum_j v_j [x^j]_2 + s delta_2      = MSM(xi_2 ++ [beta_2,  delta_2], v ++ [1, s])
    B_g1 = beta_1  + sum_j v_j [x^j]_1 + s delta_1      = MSM(xi_1 ++ [beta_1,  delta_1], v ++ [1, s])
    C    = sum_{i>l} a_i uvw_wit_i + sum_j h_j [x^j t/delta]_1 - (r s) delta_1   (one MSM)
           + s A + r B_g1                                                          (one 2-term MSM)

All scalar work here is exact integer arithmetic mod r on the host (the reference's Fr aggregation,
qap.rs:99-109, is the next row of the scope table); every group operation runs in libzkmsm.so.
"""
from dataclasses import dataclass

import numpy as np

from .api import G1Point, G1Points, G2Point, G2Points, R, default_context, scalars_to_array


@dataclass
class Proof:  # proof.rs:7-11
    A: G1Point
    B: G2Point
    C: G1Point


class DeviceCRS:
    """CRS vectors of crs.rs:17-43 resident on one GPU, with the single points appended (see above)."""

    def __init__(self, g1_alpha, g1_beta, g1_delta, g1_xi, g1_uvw_wit, g1_xt_by_delta, g2_beta, g2_delta, g2_xi,
                 precompute=True, ctx=None):
        self.ctx = ctx or default_context()
        self.n = len(g1_xi)
        self.n_wit = len(g1_uvw_wit)
        self.n_xt = len(g1_xt_by_delta)
        self.g1_delta = g1_delta
        mk1 = lambda pts: G1Points(pts, precompute=precompute, ctx=self.ctx, in_subgroup=True)  # CRS points have order r
        self.set_A = mk1(list(g1_xi) + [g1_alpha, g1_delta])
        self.set_Bg1 = mk1(list(g1_xi) + [g1_beta, g1_delta])
        self.set_B = G2Points(list(g2_xi) + [g2_beta, g2_delta], precompute=precompute, ctx=self.ctx, in_subgroup=True)
        self.set_C = mk1(list(g1_uvw_wit) + list(g1_xt_by_delta) + [g1_delta])

    def contexts(self, k):
        """k contexts on this CRS's device (the first is the CRS's own), created on first use"""
        from .context import Context
        pool = getattr(self, "_pool", None)
        if pool is None:
            pool = self._pool = [self.ctx]
        while len(pool) < k:
            pool.append(Context(self.ctx.device))
        return pool[:k]

    @classmethod
    def from_arrays(cls, arrs, precompute=True, ctx=None):
        """arrs: dict of canonical limb arrays (for large synthetic instances built on the device):
        g1_xi (n,24), g1_uvw_wit, g1_xt_by_delta, g2_xi (n,48), and single points g1_alpha, g1_beta,
        g1_delta (24,), g2_beta, g2_delta (48,)."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self.n, self.n_wit, self.n_xt = len(arrs["g1_xi"]), len(arrs["g1_uvw_wit"]), len(arrs["g1_xt_by_delta"])
        self.g1_delta = G1Point.from_limbs(arrs["g1_delta"], False)
        cat = lambda *a: np.concatenate([np.asarray(x, dtype=np.uint32).reshape(-1, np.asarray(a[0]).shape[-1]) for x in a])
        self.set_A = G1Points.from_arrays(cat(arrs["g1_xi"], arrs["g1_alpha"], arrs["g1_delta"]), precompute=precompute, ctx=self.ctx, in_subgroup=True)
        self.set_Bg1 = G1Points.from_arrays(cat(arrs["g1_xi"], arrs["g1_beta"], arrs["g1_delta"]), precompute=precompute, ctx=self.ctx, in_subgroup=True)
        self.set_B = G2Points.from_arrays(cat(arrs["g2_xi"], arrs["g2_beta"], arrs["g2_delta"]), precompute=precompute, ctx=self.ctx, in_subgroup=True)
        self.set_C = G1Points.from_arrays(cat(arrs["g1_uvw_wit"], arrs["g1_xt_by_delta"], arrs["g1_delta"]), precompute=precompute, ctx=self.ctx, in_subgroup=True)
        return self


def aggregate(polys, wires, ctx=None):
    """sum_i a_i * poly_i as a coefficient list mod r (Prover's per-wire loop, prover.rs:108-117, made one
    vector) -- computed on the device by zkmsm_fr_aggregate (Montgomery Fr kernel)."""
    ctx = ctx or default_context()
    n = max(len(p) for p in polys)
    mat = np.zeros((len(polys), n, 8), dtype=np.uint32)
    for i, p in enumerate(polys):
        mat[i, :len(p)] = scalars_to_array([int(c) % R for c in p])
    out = ctx.fr_aggregate(mat, scalars_to_array([int(a) % R for a in wires]))
    return [sum(int(w) << (32 * k) for k, w in enumerate(row)) for row in out]


def quotient(u_agg, v_agg, w_agg, n, ctx=None):
    """h = (u v - w) / t on the device (Prover::new, prover.rs:64-71); raises where the reference panics"""
    ctx = ctx or default_context()
    pad = lambda p: scalars_to_array(_pad(p, n))
    h, exact = ctx.fr_quotient(pad(u_agg), pad(v_agg), pad(w_agg))
    if not exact:
        raise ValueError("p should be divisible by t")   # prover.rs:69
    return [sum(int(x) << (32 * k) for k, x in enumerate(row)) for row in h]


def _pad(v, n):
    v = [int(x) % R for x in v]
    if len(v) > n:
        if any(v[n:]):
            raise IndexError("index out of bounds: more coefficients than powers")  # polynomial.rs:278
        v = v[:n]
    return v + [0] * (n - len(v))


class Prover:
    """prover.rs:35-46 restricted to what `prove` reads: aggregated u, v coefficient vectors, h, witness wires."""

    def __init__(self, u_agg, v_agg, h, witness_wires):
        self.u, self.v, self.h, self.wit = u_agg, v_agg, h, witness_wires
        self._limbs = {}   # scalar vectors marshalled once to the ABI layout (padding depends on the CRS)

    def _arr(self, ctx, name, parts, extra):
        """scalar vector = concatenation of padded `parts` [(vec, n), ...] plus `extra` trailing slots that change
        per proof; kept in pinned host memory so that only the trailing rows are rewritten for each proof"""
        key = (name, tuple(n for _, n in parts), extra)
        if key not in self._limbs:
            total = sum(n for _, n in parts) + extra
            arr = ctx.pinned_array((total, 8))
            pos = 0
            for vec, n in parts:
                arr[pos:pos + n] = scalars_to_array(_pad(vec, n))
                pos += n
            self._limbs[key] = arr
        return self._limbs[key]

    @classmethod
    def from_per_wire(cls, ui, vi, h, wires, l):
        """ui, vi: per-wire coefficient lists (prover.ui / prover.vi), wires a_0..a_m, l = last statement wire"""
        return cls(aggregate(ui, wires), aggregate(vi, wires), list(h), list(wires[l + 1:]))

    @classmethod
    def from_qap(cls, ui, vi, wi, wires, l, n, ctx=None):
        """Prover::new's arithmetic part (prover.rs:64-93) on the device: aggregate the per-wire polynomials
        (zkmsm_fr_aggregate), then h = (u v - w) / t (zkmsm_fr_quotient).  ui, vi, wi: per-wire coefficient
        lists of the QAP (prover.ui / vi / wi), wires a_0..a_m, l = last statement wire, n = constraints."""
        u, v, w = (aggregate(p, wires, ctx) for p in (ui, vi, wi))
        return cls(u, v, quotient(u, v, w, n, ctx), list(wires[l + 1:]))

    def prove(self, crs: DeviceCRS, r: int, s: int) -> Proof:
        ctx = crs.ctx
        r, s = int(r) % R, int(s) % R
        n = crs.n
        su = self._arr(ctx, "u", [(self.u, n)], 2)
        su[n:] = scalars_to_array([1, r])
        sv = self._arr(ctx, "v", [(self.v, n)], 2)
        sv[n:] = scalars_to_array([1, s])
        if len(self.wit) != crs.n_wit:
            raise IndexError("witness length does not match crs.g1.uvw_wit")
        sc = self._arr(ctx, "c", [(self.wit, crs.n_wit), (self.h, crs.n_xt)], 1)
        sc[crs.n_wit + crs.n_xt:] = scalars_to_array([(-(r * s)) % R])
        # the four large MSMs are independent: one context (stream + workspace) each, all in flight at once, so
        # the latency-bound reduction tail of one overlaps the accumulation of the others
        cA, cB, cBg1, cC = crs.contexts(4)
        cB.msm_begin(crs.set_B.set, sv)                                           # prover.rs:119 (G2, the longest)
        cA.msm_begin(crs.set_A.set, su)                                           # :118
        cBg1.msm_begin(crs.set_Bg1.set, sv)                                       # :120
        cC.msm_begin(crs.set_C.set, sc)                                           # :128-133 and -(delta r) s
        A = G1Point.from_limbs(*cA.msm_result(1))
        B_g1 = G1Point.from_limbs(*cBg1.msm_result(1))
        C_main = G1Point.from_limbs(*cC.msm_result(1))
        B = G2Point.from_limbs(*cB.msm_result(2))
        # s A + r B_g1 needs A and B_g1; it is a 2-term MSM, run once the device is idle again
        xy, inf = G1Point.pack([A, B_g1])
        C_blind = G1Point.from_limbs(*cA.msm_oneshot(1, xy, inf if inf.any() else None, scalars_to_array([s, r])))  # :137-138
        return Proof(A, B, C_main + C_blind)
