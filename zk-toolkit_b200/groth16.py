"""Groth16 proof generation on the GPU: `Prover::prove`
(src/zk/w_trusted_setup/groth16/zktoolkit_based/prover.rs:96-147) through zkmsm_groth16_prove.

The reference evaluates, per wire i, `ui[i].eval_with_g1_hidings(xi) * a_i` and sums the results
(prover.rs:108-117): (m+1) nested MSMs.  Because the group is commutative that equals ONE MSM with the
aggregated coefficients  sum_i a_i * u_{i,j}  (the canonical affine result is identical).
The library (csrc/zkmsm.cu, csrc/fr_ops.cuh Groth16Scalars) then runs three concurrent MSMs:

    A = alpha  + sum_j u_j [x^j]_1 + r delta           = MSM(xi_1 ++ [alpha_1, delta_1],  u ++ [1, r])
    B = beta_2 + sum_j v_j [x^j]_2 + s delta_2         = MSM(xi_2 ++ [beta_2,  delta_2],  v ++ [1, s])
    C = sum_{i>l} a_i uvw_wit_i + sum_j h_j [x^j t/delta]_1 + s A + r B_g1 - r s delta          (prover.rs:128-145)
      = MSM(uvw_wit ++ xt_by_delta ++ xi_1 ++ [alpha_1, beta_1, delta_1],  wit ++ h ++ (s u + r v) ++ [s, r, r s])

(expand A and B_g1 = beta_1 + sum v_j [x^j]_1 + s delta_1), so B_g1 and the two 255-bit scalar multiplications
s A, r B_g1 are never formed.  This module only marshals Python integers to the ABI's limb layout; every field
and group operation runs in libzkmsm.so.
"""
import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .api import G1Point, G2Point, R, default_context, scalars_to_array


@dataclass
class Proof:  # proof.rs:7-11
    A: G1Point
    B: G2Point
    C: G1Point


def _proof_from_limbs(out, inf):
    return Proof(G1Point.from_limbs(out[:24], inf[0]), G2Point.from_limbs(out[24:72], inf[1]),
                 G1Point.from_limbs(out[72:96], inf[2]))


class DeviceCRS:
    """CRS vectors of crs.rs:17-43 resident on one GPU (zkmsm_crs_load).  in_subgroup=True asserts that every CRS point
    has order r -- true for any CRS built as multiples of the generators (crs.rs:65-135); check_subgroup=True makes
    the device verify it at load time (r P = AtInfinity for every point, ZkmsmError otherwise).  Pass
    in_subgroup=False for points of unknown provenance: the device then assumes nothing (macros.rs:10-21)."""

    def __init__(self, g1_alpha, g1_beta, g1_delta, g1_xi, g1_uvw_wit, g1_xt_by_delta, g2_beta, g2_delta, g2_xi,
                 precompute=True, ctx=None, in_subgroup=True, check_subgroup=False):
        p1 = lambda pts: G1Point.pack(list(pts))
        xi, xi_inf = p1(g1_xi)
        wit, wit_inf = p1(g1_uvw_wit)
        xt, xt_inf = p1(g1_xt_by_delta)
        xi2, xi2_inf = G2Point.pack(list(g2_xi))
        for single in (g1_alpha, g1_beta, g1_delta, g2_beta, g2_delta):
            if single.is_zero():
                raise ValueError("alpha, beta, delta must not be AtInfinity")
        self._load(dict(g1_alpha=g1_alpha.limbs(), g1_beta=g1_beta.limbs(), g1_delta=g1_delta.limbs(), g1_xi=xi,
                        g1_uvw_wit=wit, g1_xt_by_delta=xt, g2_beta=g2_beta.limbs(), g2_delta=g2_delta.limbs(), g2_xi=xi2),
                   dict(g1_xi=xi_inf, g1_uvw_wit=wit_inf, g1_xt_by_delta=xt_inf, g2_xi=xi2_inf), precompute, ctx, in_subgroup,
                   check_subgroup)

    @classmethod
    def from_arrays(cls, arrs, precompute=True, ctx=None, in_subgroup=True, check_subgroup=False):
        """arrs: dict of canonical limb arrays (for large synthetic instances built on the device):
        g1_xi (n,24), g1_uvw_wit, g1_xt_by_delta, g2_xi (n,48), and single points g1_alpha, g1_beta,
        g1_delta (24,), g2_beta, g2_delta (48,)."""
        self = cls.__new__(cls)
        self._load(arrs, {}, precompute, ctx, in_subgroup, check_subgroup)
        return self

    def _load(self, arrs, infs, precompute, ctx, in_subgroup, check_subgroup=False):
        self.ctx = ctx or default_context()
        self.handle = None
        keep = {}
        d = L.CrsDesc()
        for name, words in (("g1_alpha", 24), ("g1_beta", 24), ("g1_delta", 24), ("g1_xi", 24), ("g1_uvw_wit", 24),
                            ("g1_xt_by_delta", 24), ("g2_beta", 48), ("g2_delta", 48), ("g2_xi", 48)):
            a = np.ascontiguousarray(np.asarray(arrs[name], dtype=np.uint32).reshape(-1, words))
            keep[name] = a
            setattr(d, name, a.ctypes.data if a.size else None)
        for name in ("g1_xi", "g1_uvw_wit", "g1_xt_by_delta", "g2_xi"):
            f = infs.get(name)
            if f is not None and np.any(f):
                f = np.ascontiguousarray(f, dtype=np.uint8)
                keep[name + "_inf"] = f
                setattr(d, name + "_inf", f.ctypes.data)
        self.n, self.n_wit, self.n_xt = len(keep["g1_xi"]), len(keep["g1_uvw_wit"]), len(keep["g1_xt_by_delta"])
        assert len(keep["g2_xi"]) == self.n
        d.n, d.n_wit, d.n_xt = self.n, self.n_wit, self.n_xt
        self.g1_delta = G1Point.from_limbs(keep["g1_delta"][0], False)
        flags = (L.PRECOMPUTE if precompute else 0) | (L.SUBGROUP if in_subgroup else 0) | (L.CHECK_SUBGROUP if check_subgroup else 0)
        h = ctypes.c_void_p()
        self.ctx._check(self.ctx.lib.zkmsm_crs_load(self.ctx.h, ctypes.byref(d), flags, ctypes.byref(h)))
        self.handle = h

    def free(self):
        if getattr(self, "handle", None) is not None and self.ctx.h:
            self.ctx.lib.zkmsm_crs_free(self.ctx.h, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


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
        self._limbs = {}   # (context id, CRS sizes) -> the four vectors marshalled once to the ABI layout, pinned

    def _arrays(self, crs):
        """u, v (n), h (n_xt), witness (n_wit) as canonical limb arrays in pinned host memory of the CRS's context
        (keyed by that context: the memory is owned by it)"""
        key = (id(crs.ctx), crs.n, crs.n_xt, crs.n_wit)
        if key not in self._limbs:
            if len(self.wit) != crs.n_wit:
                raise IndexError("witness length does not match crs.g1.uvw_wit")
            out = []
            for vec, n in ((self.u, crs.n), (self.v, crs.n), (self.h, crs.n_xt), (self.wit, crs.n_wit)):
                arr = crs.ctx.pinned_array((max(n, 1), 8))
                arr[:n] = scalars_to_array(_pad(vec, n))
                out.append(arr)
            self._limbs[key] = (crs.ctx, out)      # the context reference keeps the pinned memory alive
        return self._limbs[key][1]

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
        """Prover::prove (prover.rs:96-147) with the blinding scalars supplied by the caller"""
        u, v, h, wit = self._arrays(crs)
        rs = scalars_to_array([int(r) % R, int(s) % R])
        out = np.zeros(96, dtype=np.uint32)
        inf = (ctypes.c_int * 3)()
        ctx = crs.ctx
        ctx._check(ctx.lib.zkmsm_groth16_prove(ctx.h, crs.handle, L.dptr(u), L.dptr(v), L.dptr(h), L.dptr(wit), L.dptr(rs[0:1]),
                                               L.dptr(rs[1:2]), L.dptr(out), inf))
        return _proof_from_limbs(out, inf)

    def prove_partial(self, crs: DeviceCRS, r: int, s: int, rank: int, world: int):
        """this device's 1/world share of the three MSMs (bucket-range split): a (192,) uint32 blob for
        combine_partials (configs[4]: one proof over the GPUs of a box)"""
        u, v, h, wit = self._arrays(crs)
        rs = scalars_to_array([int(r) % R, int(s) % R])
        out = np.zeros(L.GROTH16_PARTIAL_WORDS, dtype=np.uint32)
        ctx = crs.ctx
        ctx._check(ctx.lib.zkmsm_groth16_prove_partial(ctx.h, crs.handle, L.dptr(u), L.dptr(v), L.dptr(h), L.dptr(wit),
                                                       L.dptr(rs[0:1]), L.dptr(rs[1:2]), rank, world, L.dptr(out)))
        return out


def _prove_partial_gathered(self, crs, r, s, rank, world):
    import torch
    import torch.distributed as dist
    u, v, h, wit = self._arrays(crs)
    key = ("gathered", id(crs.ctx), crs.n, crs.n_xt, crs.n_wit, world)
    if key not in self._limbs:
        dev = torch.device("cuda", crs.ctx.device)
        bufs = []
        for vec, m in ((u, crs.n), (v, crs.n), (h, crs.n_xt), (wit, crs.n_wit)):
            per = max(1, -(-m // world))
            lo, hi = min(m, rank * per), min(m, (rank + 1) * per)
            src = torch.from_numpy(vec[lo:hi].view(np.int32)) if hi > lo else None
            d_slice = torch.zeros(per, 8, dtype=torch.int32, device=dev)
            d_full = torch.empty(world * per, 8, dtype=torch.int32, device=dev)
            bufs.append((src, hi - lo, d_slice, d_full))
        self._limbs[key] = (crs.ctx, bufs)
    bufs = self._limbs[key][1]
    for src, cnt, d_slice, d_full in bufs:
        if cnt:
            d_slice[:cnt].copy_(src, non_blocking=True)
        dist.all_gather_into_tensor(d_full.view(-1), d_slice.view(-1))
    torch.cuda.current_stream().synchronize()      # the library runs on its own stream
    rs = scalars_to_array([int(r) % R, int(s) % R])
    out = np.zeros(L.GROTH16_PARTIAL_WORDS, dtype=np.uint32)
    ctx = crs.ctx
    P = ctypes.c_void_p
    ctx._check(ctx.lib.zkmsm_groth16_prove_partial(ctx.h, crs.handle, P(bufs[0][3].data_ptr()), P(bufs[1][3].data_ptr()),
                                                   P(bufs[2][3].data_ptr()), P(bufs[3][3].data_ptr()), L.dptr(rs[0:1]), L.dptr(rs[1:2]),
                                                   rank, world, L.dptr(out)))
    return out


Prover._prove_partial_gathered = _prove_partial_gathered


def combine_partials(partials, ctx=None) -> Proof:
    """sum of the per-rank shares of prove_partial, in rank order (zkmsm_groth16_combine)"""
    ctx = ctx or default_context()
    parts = np.ascontiguousarray(np.asarray(partials, dtype=np.uint32).reshape(-1, L.GROTH16_PARTIAL_WORDS))
    out = np.zeros(96, dtype=np.uint32)
    inf = (ctypes.c_int * 3)()
    ctx._check(ctx.lib.zkmsm_groth16_combine(ctx.h, L.dptr(parts), parts.shape[0], L.dptr(out), inf))
    return _proof_from_limbs(out, inf)


def prove_distributed(prover: Prover, crs: DeviceCRS, r: int, s: int, timings=None) -> Proof:
    """One proof over all ranks of the default torch.distributed group (one process per GPU, every rank holds the
    CRS): each rank proves its share, ONE all-gather of 192 words per rank, every rank combines (rank order, so the
    proof is identical everywhere and identical to Prover.prove on one GPU).  timings: optional dict that receives
    the wall-clock split of this call in ms (share, exchange, combine)."""
    import time
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return prover.prove(crs, r, s)
    t0 = time.perf_counter()
    if dist.get_backend() == "nccl":
        # every rank uploads 1/world of each coefficient vector and the slices are all-gathered over NVLink: the
        # scalars cross PCIe once per proof instead of once per GPU (8 concurrent 32 MB uploads measured 1.7 ms each)
        mine = prover._prove_partial_gathered(crs, r, s, rank, world)
    else:
        mine = prover.prove_partial(crs, r, s, rank, world)
    t1 = time.perf_counter()
    if dist.get_backend() == "nccl":
        # the blob is already on the host (the share call waits for its three MSMs): pinned staging both ways
        st = getattr(crs, "_xchg", None)
        if st is None:
            dev = torch.device("cuda", crs.ctx.device)
            st = crs._xchg = (torch.empty(L.GROTH16_PARTIAL_WORDS, dtype=torch.int32).pin_memory(),
                              torch.empty(L.GROTH16_PARTIAL_WORDS, dtype=torch.int32, device=dev),
                              torch.empty(world * L.GROTH16_PARTIAL_WORDS, dtype=torch.int32, device=dev),
                              torch.empty(world * L.GROTH16_PARTIAL_WORDS, dtype=torch.int32).pin_memory())
        h_in, d_in, d_out, h_out = st
        h_in.numpy()[:] = mine.view(np.int32)
        d_in.copy_(h_in, non_blocking=True)
        dist.all_gather_into_tensor(d_out, d_in)
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        parts = h_out.numpy().view(np.uint32)
    else:
        t = torch.from_numpy(mine.view(np.int32))
        out = torch.empty(world * L.GROTH16_PARTIAL_WORDS, dtype=torch.int32)
        dist.all_gather_into_tensor(out, t)
        parts = out.numpy().view(np.uint32)
    t2 = time.perf_counter()
    proof = combine_partials(parts, crs.ctx)
    if timings is not None:
        t3 = time.perf_counter()
        timings.update(share_ms=(t1 - t0) * 1e3, exchange_ms=(t2 - t1) * 1e3, combine_ms=(t3 - t2) * 1e3)
    return proof


class CRS:
    """CRS::new (groth16/zktoolkit_based/crs.rs:49-146) with the trapdoor supplied by the caller: the scalars of
    crs.rs:65-116 are formed on the host (setup, exact Python integers mod r), every point is a device fixed-base
    multiplication (`&G1Point * &Fq1` for a vector of scalars, zkmsm_g{1,2}_mul_base), and `device()` makes the
    resident prover-side CRS.  Fields as the reference's: g1.alpha, beta, delta, xi, uvw_stmt, uvw_wit, xt_by_delta;
    g2.beta, gamma, delta, xi (canonical limb arrays + AtInfinity flags)."""

    def __init__(self, ui, vi, wi, l, n, alpha, beta, gamma, delta, x, ctx=None):
        self.ctx = ctx or default_context()
        ev = lambda poly: sum(int(c) * pow(x, j, R) for j, c in enumerate(poly)) % R       # Polynomial::eval_at
        inv = lambda a: pow(a % R, -1, R)
        m = len(ui) - 1

        def uvw_div(lo, hi, div):                                                          # crs.rs:65-83
            return [((beta * ev(ui[i]) + alpha * ev(vi[i]) + ev(wi[i])) * div) % R for i in range(lo, hi + 1)]

        xp = [pow(x, j, R) for j in range(n)]                                              # crs.rs:88-102
        t = 1
        for k in range(1, n + 1):                                                          # QAP::build_t at x, qap.rs:115-135
            t = t * ((x - k) % R) % R
        g1 = lambda ks: self.ctx.mul_base(1, G1Point.g().limbs(), scalars_to_array([k % R for k in ks]))
        g2 = lambda ks: self.ctx.mul_base(2, G2Point.g().limbs(), scalars_to_array([k % R for k in ks]))
        self.n, self.l, self.m = n, l, m
        self.g1_uvw_stmt, self.g1_uvw_stmt_inf = g1(uvw_div(0, l, inv(gamma)))             # crs.rs:85
        self.g1_uvw_wit, self.g1_uvw_wit_inf = g1(uvw_div(l + 1, m, inv(delta)))           # crs.rs:86
        self.g1_xi, self.g1_xi_inf = g1(xp)
        self.g1_xt_by_delta, self.g1_xt_by_delta_inf = g1([p * t % R * inv(delta) % R for p in xp])   # crs.rs:104-116
        singles1, _ = g1([alpha, beta, delta])
        self.g1_alpha, self.g1_beta, self.g1_delta = singles1
        self.g2_xi, self.g2_xi_inf = g2(xp)
        singles2, _ = g2([beta, gamma, delta])
        self.g2_beta, self.g2_gamma, self.g2_delta = singles2

    def device(self, precompute=True) -> DeviceCRS:
        arrs = {k: getattr(self, k) for k in ("g1_alpha", "g1_beta", "g1_delta", "g1_xi", "g1_uvw_wit", "g1_xt_by_delta",
                                              "g2_beta", "g2_delta", "g2_xi")}
        crs = DeviceCRS.__new__(DeviceCRS)
        crs._load(arrs, {k: getattr(self, k + "_inf") for k in ("g1_xi", "g1_uvw_wit", "g1_xt_by_delta", "g2_xi")},
                  precompute, self.ctx, True)
        return crs
