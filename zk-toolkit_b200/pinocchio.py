"""Pinocchio proof generation on the GPU: the second consumer of the MSM seam in the reference
(src/zk/w_trusted_setup/pinocchio/prover.rs:95-171).  Its eight inline loops
`x_mid_s += &ek.x_mid[i] * w` (prover.rs:118-128) are eight MSMs over the intermediate witness, and
`h.eval_with_g2_hidings(&ek.si)` (prover.rs:136) a ninth; the blinding terms (vk.t * delta_v, ...) are folded in
by appending the single key points to the resident sets, exactly as groth16.py does."""
from .api import G1Point, G1Points, G2Point, G2Points, R, default_context, scalars_to_array


class DeviceKeys:
    """Evaluation / verification key vectors of pinocchio/crs.rs:17-52 resident on one GPU."""

    def __init__(self, crs, to_g1, to_g2, precompute=False, ctx=None):
        """crs: object with the fields of crs.rs (lists of points in any representation); to_g1 / to_g2 convert one
        point to G1Point / G2Point."""
        self.ctx = ctx or default_context()
        g1 = lambda pts: G1Points([to_g1(p) for p in pts], precompute=precompute, ctx=self.ctx, in_subgroup=True)
        g2 = lambda pts: G2Points([to_g2(p) for p in pts], precompute=precompute, ctx=self.ctx, in_subgroup=True)
        self.n_mid = len(crs.vk_mid)
        self.v = g1(list(crs.vk_mid) + [crs.t])
        self.w1 = g1(crs.g1_wk_mid)
        self.w2 = g2(list(crs.g2_wk_mid) + list(crs.wk_io))
        self.y = g1(list(crs.yk_mid) + [crs.t])
        self.av = g1(list(crs.alpha_vk_mid) + [crs.alpha_v_t])
        self.aw = g1(crs.alpha_wk_mid)
        self.ay = g1(list(crs.alpha_yk_mid) + [crs.alpha_y_t])
        self.b = g1(list(crs.beta_vwy_k_mid) + [crs.beta_t])
        self.si = g2(crs.si)
        self.n_io = len(crs.wk_io)
        self.one_g2 = to_g2(crs.one_g2)


def prove(keys: DeviceKeys, witness_mid, witness_io, h, delta_v, delta_y):
    """prover.rs:95-171.  witness_mid / witness_io: the two slices of the witness (witness.rs:20-27), h: quotient
    coefficients (from groth16.quotient or the caller).  Returns the nine proof elements (proof.rs) as a dict."""
    ctx = keys.ctx
    dv, dy = int(delta_v) % R, int(delta_y) % R
    mid = [int(w) % R for w in witness_mid]
    io = [int(w) % R for w in witness_io][: keys.n_io]
    g1 = lambda pts, sc: G1Point.from_limbs(*ctx.msm(pts.set, scalars_to_array(sc)))
    g2 = lambda pts, sc: G2Point.from_limbs(*ctx.msm(pts.set, scalars_to_array(sc)))
    proof = dict(
        v_mid_s=g1(keys.v, mid + [dv]),                       # vk.t * delta_v + sum vk_mid[i] * w_i
        g1_w_mid_s=g1(keys.w1, mid),
        g2_w_mid_s=g2(keys.w2, mid),                          # uses the first n_mid points
        y_mid_s=g1(keys.y, mid + [dy]),
        alpha_v_mid_s=g1(keys.av, mid + [dv]),
        alpha_w_mid_s=g1(keys.aw, mid),
        alpha_y_mid_s=g1(keys.ay, mid + [dy]),
        beta_vwy_mid_s=g1(keys.b, mid + [(dv + dy) % R]),     # beta_t * delta_v + beta_t * delta_y
    )
    hc = [int(c) % R for c in h]
    while len(hc) > 1 and hc[-1] == 0:
        hc.pop()
    h_s = g2(keys.si, hc)                                     # h.eval_with_g2_hidings(&ek.si), prover.rs:136
    w_s = g2(keys.w2, mid + io)                               # g2_w_mid_s + sum wk_io[i] * io_i, prover.rs:138-142
    proof["h_s"] = h_s + w_s * dv + (-(keys.one_g2 * dy))     # prover.rs:144
    return proof
