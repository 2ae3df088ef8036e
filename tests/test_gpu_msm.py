"""Parity of the CUDA path (through the C ABI) with the T0 oracle and the reference's golden
vectors.  Bit-exact: every comparison is on canonical affine coordinates."""
import json
import os
import random

import numpy as np
import pytest

from oracle import zkt_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def z():
    import zk_toolkit_b200 as z
    return z


@pytest.fixture(scope="module")
def ctx(z):
    return z.default_context()


def load(name):
    with open(os.path.join(G, f"ref_{name}.json")) as f:
        return json.load(f)


def to_o1(p):
    return O.INF if p.is_zero() else O.g1(p.x, p.y)


def to_o2(p):
    if p.is_zero():
        return O.INF
    c = p.coords
    return O.g2(c[1], c[0], c[3], c[2])


# ---------------------------------------------------------------- reference KATs through the GPU API
def test_g1_kats_through_api(z):
    k = load("g1")
    g = z.G1Point.g()
    gs = [z.G1Point.new(int(p["x"]), int(p["y"])) for p in k["g_multiples"]["points"]]
    assert g + g == z.G1Point.new(int(k["add_same_point"]["x"]), int(k["add_same_point"]["y"]))   # g1_point.rs:223-237
    for n in range(1, 11):                                                                         # :347-356
        assert g * n == gs[n - 1]
    for c in k["scalar_mul_gen_pubkey"]["cases"]:                                                  # :352-371
        assert g * (int(c["multiple"]) % O.Q) == z.G1Point.new(int(c["x"]), int(c["y"]))
    for a, b, c in k["add_different_points"]["cases"]:                                             # :389-412
        assert gs[a - 1] + gs[b - 1] == gs[c - 1]
    inf = z.G1Point.zero()
    assert (g + (-g)).is_zero() and g + inf == g and inf + g == g and (inf + inf).is_zero()       # :239-296
    assert (g * 0).is_zero() and (g * O.R).is_zero() and g * 1 == g
    assert g * 2 == g + g and g * 3 == g + g + g                                                   # :203-221


def test_g2_kats_through_api(z):
    k = load("g2")
    mk = lambda p: z.G2Point.new(int(p["x1"]), int(p["x0"]), int(p["y1"]), int(p["y0"]))
    g = z.G2Point.g()
    gs = [mk(p) for p in k["g_multiples"]["points"]]
    assert g + g == mk(k["add_same_point"])                                                        # g2_point.rs:199-230
    for n in range(1, 11):
        assert g * n == gs[n - 1]
    for c in k["scalar_mul_gen_pubkey"]["cases"]:
        assert g * int(c["multiple"]) == mk(c)
    for a, b, c in k["add_different_points"]["cases"]:
        assert gs[a - 1] + gs[b - 1] == gs[c - 1]
    inf = z.G2Point.zero()
    assert (g + (-g)).is_zero() and g + inf == g and inf + g == g and (inf + inf).is_zero()
    assert (g * O.R).is_zero()


# ---------------------------------------------------------------- MSM vs the reference's serial loop
@pytest.fixture(scope="module")
def small_g1(z):
    rnd = random.Random(42)
    dlogs = [rnd.randrange(1, O.R) for _ in range(40)]
    opts = [O.scalar_mul(O.G1_GEN, k) for k in dlogs]
    pts = [z.G1Point.new(p[0].e, p[1].e) for p in opts]
    return dlogs, opts, pts


def test_msm_seam_small_vs_oracle(z, small_g1):
    dlogs, opts, pts = small_g1
    rnd = random.Random(1)
    for n in (1, 2, 3, 7, 16):
        sc = U.rand_scalars(rnd, n)
        exp = O.msm(opts[:n], sc)
        assert to_o1(z.Polynomial(sc).eval_with_g1_hidings(pts[:n])) == exp       # host slice, uploaded per call
        dev = z.G1Points(pts[:n])
        assert to_o1(z.Polynomial(sc).eval_with_g1_hidings(dev)) == exp
    # polynomial.rs:1250-1285: 4 terms, toy scalars
    g = z.G1Point.g()
    pw = [g * 1, g * 2, g * 3, g * 4]
    assert z.Polynomial([2, 3, 4, 5]).eval_with_g1_hidings(pw) == g * 40
    with pytest.raises(IndexError):
        z.Polynomial([1, 2, 3]).eval_with_g1_hidings(pw[:2])                  # polynomial.rs:278 panics
    assert z.Polynomial([0]).eval_with_g1_hidings(pw).is_zero()


def test_msm_edge_cases(z, ctx, small_g1):
    dlogs, opts, pts = small_g1
    n = 12
    P, D, OP = pts[:n], dlogs[:n], opts[:n]
    dev = z.G1Points(P)
    pre = z.G1Points(P, precompute=True)

    def run(s, points=dev, cnt=None):
        out, inf = ctx.msm(points.set, z.scalars_to_array(s), n=cnt)
        return U.g1_from_array(out, inf)

    exp = lambda s, d=D: U.expected_from_dlogs(O.G1_GEN, d, s)
    assert run([]) is O.INF                                                    # n = 0
    assert run([0] * n) is O.INF
    s = [0, 1, O.R - 1, 2, 0, O.R - 2, 1, 1, 5, 0, 7, O.R - 1]
    assert run(s) == exp(s) == run(s, pre)
    s = [0x1234567890ABCDEF1234567890ABCDEF] * n                               # all-equal scalars
    assert run(s) == exp(s) == run(s, pre)
    dup = z.G1Points([P[0]] * n)                                               # duplicate points -> doubling inside a bucket
    assert run([3] * n, dup) == O.scalar_mul(OP[0], 3 * n)
    pm = z.G1Points([P[0], -P[0], P[1], -P[1]])                                # P + (-P)
    assert run([9, 9, 11, 11], pm) is O.INF
    assert run([9, 9, 11, 10], pm) == OP[1]
    wi = z.G1Points([P[0], z.G1Point.zero(), P[1], z.G1Point.zero()])          # AtInfinity inputs
    assert run([5, 6, 7, 8], wi) == O.msm([OP[0], O.INF, OP[1], O.INF], [5, 6, 7, 8])
    assert run([5, 6, 7], dev, 3) == exp([5, 6, 7], D[:3]) == run([5, 6, 7], pre, 3)   # extra points ignored
    s = [(1 << 255) - 1, (1 << 255) - 19]
    assert run(s, z.G1Points(P[:2])) == O.msm(OP[:2], s)
    assert run([O.R], z.G1Points(P[:1])) is O.INF
    with pytest.raises(z.ZkmsmError) as e:                                     # bit 255 set: rejected
        run([1 << 255, 1], z.G1Points(P[:2]))
    assert e.value.code == -3
    half = (O.R - 1) // 2                                                      # subgroup sets: s -> r - s, P -> -P
    s = [half, half + 1, O.R - 1, 1, 0, half - 1, O.R - 2, 2, 3, 4, 5, 6]
    for pre_ in (False, True):
        assert run(s, z.G1Points(P, precompute=pre_, in_subgroup=True)) == exp(s)
    with pytest.raises(z.ZkmsmError) as e:                                     # ... and s >= r is rejected there
        run([O.R, 1], z.G1Points(P[:2], in_subgroup=True))
    assert e.value.code == -3
    with pytest.raises(z.ZkmsmError) as e:                                     # more scalars than points
        ctx.msm(z.G1Points(P[:2]).set, z.scalars_to_array([1, 2, 3]))
    assert e.value.code == -5


@pytest.mark.parametrize("logn", [6, 10, 12, 16])
@pytest.mark.parametrize("precompute", [False, True])
def test_msm_random_vs_dlog_identity(z, ctx, logn, precompute):
    """points = k_i * g made on the device (spot-checked against the oracle), so the full-size
    result has a closed form: (sum s_i k_i mod r) * g, one oracle scalar multiplication."""
    n = 1 << logn
    rnd = random.Random(1000 + logn)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    pts = z.G1Points.generator_multiples(dlogs, precompute=precompute)
    for i in (0, 1, n // 2, n - 1):
        assert to_o1(pts[i]) == O.scalar_mul(O.G1_GEN, dlogs[i])
    for trial in range(2):
        sc = U.rand_scalars(rnd, n)
        out, inf = ctx.msm(pts.set, z.scalars_to_array(sc))
        assert U.g1_from_array(out, inf) == U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
    # ragged: 2^k +- 1 terms of the same set
    for m in (n - 1, n // 2 + 1):
        sc = U.rand_scalars(rnd, m)
        out, inf = ctx.msm(pts.set, z.scalars_to_array(sc))
        assert U.g1_from_array(out, inf) == U.expected_from_dlogs(O.G1_GEN, dlogs[:m], sc)


@pytest.mark.parametrize("c", [4, 9, 13, 16])
def test_msm_forced_windows(z, ctx, c):
    n = 3000
    rnd = random.Random(c)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    exp = U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
    ctx.set_window(c)
    try:
        for pre in (False, True):
            pts = z.G1Points.generator_multiples(dlogs, precompute=pre)
            out, inf = ctx.msm(pts.set, z.scalars_to_array(sc))
            assert U.g1_from_array(out, inf) == exp
    finally:
        ctx.set_window(0)


def test_msm_sharded_partials(z, ctx):
    """the multi-GPU path on one device: k shards -> k partial points -> combine; partition independent"""
    n = 5000
    rnd = random.Random(77)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    exp = U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
    sca = z.scalars_to_array(sc)
    for k in (1, 2, 3, 8):
        parts = []
        for r in range(k):
            lo, hi = n * r // k, n * (r + 1) // k
            shard = z.G1Points.generator_multiples(dlogs[lo:hi])
            parts.append(ctx.msm_partial(shard.set, sca[lo:hi]))
        out, inf = ctx.combine(1, np.stack(parts))
        assert U.g1_from_array(out, inf) == exp
    out, inf = ctx.combine(1, np.zeros((3, 48), dtype=np.uint32))
    assert inf


def test_g2_msm(z, ctx):
    rnd = random.Random(9)
    n = 600
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    exp = U.expected_from_dlogs(O.G2_GEN, dlogs, sc)
    for pre in (False, True):
        pts = z.G2Points.generator_multiples(dlogs, precompute=pre)
        assert to_o2(pts[5]) == O.scalar_mul(O.G2_GEN, dlogs[5])
        out, inf = ctx.msm(pts.set, z.scalars_to_array(sc))
        assert U.g2_from_array(out, inf) == exp
    small = [z.G2Point.g() * k for k in (1, 2, 3)]
    got = z.Polynomial([4, 5, 6]).eval_with_g2_hidings(small)
    assert to_o2(got) == O.scalar_mul(O.G2_GEN, 4 + 10 + 18)


def test_enqueue_result_split_and_launch_count(z, ctx):
    import torch
    n = 4096
    rnd = random.Random(3)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    pts = z.G1Points.generator_multiples(dlogs)
    d_sc = torch.from_numpy(z.scalars_to_array(sc).view(np.int32)).cuda()
    torch.cuda.synchronize()
    ctx.msm_enqueue(pts.set, d_sc.data_ptr(), n)
    assert ctx.last_launch_count() >= 8
    out, inf = ctx.msm_result(1)
    assert U.g1_from_array(out, inf) == U.expected_from_dlogs(O.G1_GEN, dlogs, sc)


def test_g2_edge_cases(z, ctx):
    """Appendix B cases for G2: zero scalars, duplicates (tangent inside a bucket), P / -P, AtInfinity inputs"""
    h = z.G2Point.g()
    P = [h * k for k in (3, 5, 7)]
    OP = [O.scalar_mul(O.G2_GEN, k) for k in (3, 5, 7)]

    def run(points, scalars, **kw):
        out, inf = ctx.msm(z.G2Points(points, **kw).set, z.scalars_to_array(scalars))
        return U.g2_from_array(out, inf)

    assert run(P, [0, 0, 0]) is O.INF
    assert run([P[0]] * 5, [2] * 5) == O.scalar_mul(OP[0], 10)
    assert run([P[0], -P[0], P[1]], [9, 9, 4]) == O.scalar_mul(OP[1], 4)
    assert run([P[0], z.G2Point.zero(), P[2]], [1, 5, O.R - 1]) == O.msm([OP[0], O.INF, OP[2]], [1, 5, O.R - 1])
    half = (O.R - 1) // 2
    for pre in (False, True):
        assert run(P, [half, half + 1, O.R - 1], precompute=pre, in_subgroup=True) == O.msm(OP, [half, half + 1, O.R - 1])
    out, inf = ctx.msm(z.G2Points(P).set, z.scalars_to_array([]))
    assert inf


def test_begin_result_and_points_info(z, ctx):
    rnd = random.Random(8)
    n = 2000
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    exp = U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
    pts = z.G1Points.generator_multiples(dlogs, precompute=True)
    info = pts.set.info()
    assert info["precomputed"] and info["subgroup"] and info["windows"] == 254 // info["c"] + 1
    plain = z.G1Points.generator_multiples(dlogs, in_subgroup=False)
    assert plain.set.info() == {"c": 0, "windows": 0, "precomputed": False, "subgroup": False}
    arr = ctx.pinned_array((n, 8))
    arr[:] = z.scalars_to_array(sc)
    other = z.Context(0)                       # a second context (stream + workspace) on the same device
    ctx.msm_begin(pts.set, arr)
    other.msm_begin(plain.set, arr)
    a = ctx.msm_result(1)
    b = other.msm_result(1)
    assert U.g1_from_array(*a) == exp == U.g1_from_array(*b)
    ctx.profile(True)
    ctx.msm(pts.set, arr)
    names = [nm for nm, _, _ in ctx.profile_read()]
    ctx.profile(False)
    assert "accumulate_buckets" in names and "bucket_reduce" in names and "finish" in names


def test_batched_affine_prereduction_option(z, ctx):
    """batched-affine pre-reduction rounds forced on small inputs (options batch_rounds / batch_T): identical
    results, including P + P (tangent), P + (-P), AtInfinity operands and odd leftovers inside the rounds"""
    n = 6000
    rnd = random.Random(31)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    exp = U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
    for pre in (False, True):
        pts = z.G1Points.generator_multiples(dlogs, precompute=pre)
        for rounds, T in ((1, 128), (3, 7), (4, 64)):
            ctx.set_option("batch_rounds", rounds)
            ctx.set_option("batch_T", T)
            for no_bucket_acc in (0, 1):                 # bucket sums in one launch / chunked accumulation + fix-up tree
                ctx.set_option("no_bucket_acc", no_bucket_acc)
                out, inf = ctx.msm(pts.set, z.scalars_to_array(sc))
                assert U.g1_from_array(out, inf) == exp
            ctx.set_option("no_bucket_acc", 0)
        ctx.set_option("batch_rounds", -1)
    try:
        _rare_cases_inside_rounds(z, ctx)
    finally:
        ctx.set_option("batch_rounds", -1)
        ctx.set_option("batch_T", 0)


def _rare_cases_inside_rounds(z, ctx):
    ctx.set_option("batch_rounds", 2)
    ctx.set_option("batch_T", 5)
    g = z.G1Point.g()
    dup = z.G1Points([g * 7] * 40)
    out, inf = ctx.msm(dup.set, z.scalars_to_array([5] * 40))
    assert U.g1_from_array(out, inf) == O.scalar_mul(O.G1_GEN, 7 * 5 * 40)
    P, Q = g * 11, g * 13
    opp = z.G1Points([P, -P, Q, -Q, P, -P])
    out, inf = ctx.msm(opp.set, z.scalars_to_array([9, 9, 11, 11, 3, 3]))
    assert U.g1_from_array(out, inf) == O.INF
    withinf = z.G1Points([P, z.G1Point.zero(), Q, z.G1Point.zero(), g * 17, P])
    scs = [5, 5, 5, 5, 5, 5]
    out, inf = ctx.msm(withinf.set, z.scalars_to_array(scs))
    assert U.g1_from_array(out, inf) == O.scalar_mul(O.G1_GEN, 5 * (11 + 13 + 17 + 11))


def test_default_batched_path_at_2p19(z, ctx):
    """n = 2^19 with resident CRS tables takes the default batched-affine rounds (msm_default_batch_rounds): the
    stage profile shows them and the result has its closed form; duplicates of one point in one bucket too"""
    n = 1 << 19
    rnd = random.Random(519)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    pts = z.G1Points.generator_multiples(dlogs, precompute=True)
    sc = U.rand_scalars(rnd, n)
    ctx.profile(True)
    out, inf = ctx.msm(pts.set, z.scalars_to_array(sc))
    names = [name for name, _, _ in ctx.profile_read()]
    ctx.profile(False)
    assert "batched_add_first" in names and "batched_add" in names and "accumulate_buckets" in names
    assert U.g1_from_array(out, inf) == U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
    same = [12345] * n                                   # every term lands in the same buckets: long runs of pairs
    out, inf = ctx.msm(pts.set, z.scalars_to_array(same))
    assert U.g1_from_array(out, inf) == U.expected_from_dlogs(O.G1_GEN, dlogs, same)
    # a set made of P, P, -P and AtInfinity only: inside the rounds nearly every pair is a tangent, a cancellation
    # or a copy (the rare-case paths of BatchedAddRound at full size)
    dl2 = [7, 7, O.R - 7, 0] * (n // 4)
    pts2 = z.G1Points.generator_multiples(dl2, precompute=True)
    assert pts2[3].is_zero() and to_o1(pts2[2]) == O.point_neg(O.scalar_mul(O.G1_GEN, 7))
    sc2 = U.rand_scalars(rnd, n)
    out, inf = ctx.msm(pts2.set, z.scalars_to_array(sc2))
    assert U.g1_from_array(out, inf) == U.expected_from_dlogs(O.G1_GEN, dl2, sc2)


def test_bucket_range_split_partials(z, ctx):
    """zkmsm_g1_msm_partial_range: `world` ranks each take the whole precomputed set and all scalars but 1/world of the
    bucket range; the partials combine to the MSM.  Host and stream-ordered device variants."""
    import torch
    n = 5000
    rnd = random.Random(808)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    sc[0], sc[1], sc[2] = 0, O.R - 1, (O.R - 1) // 2
    exp = U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
    pts = z.G1Points.generator_multiples(dlogs, precompute=True)
    arr = z.scalars_to_array(sc)
    for world in (1, 2, 8):
        parts = [ctx.msm_partial_range(pts.set, arr, r, world) for r in range(world)]
        assert U.g1_from_array(*ctx.combine(1, np.stack(parts))) == exp
    d_sc = torch.from_numpy(arr.view(np.int32)).cuda()
    d_parts = torch.zeros(4, 48, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for r in range(4):
        ctx.msm_partial_range_device(pts.set, d_sc.data_ptr(), n, r, 4, d_parts[r].data_ptr())
    assert U.g1_from_array(*ctx.combine_device(d_parts.data_ptr(), 4)) == exp
    plain = z.G1Points.generator_multiples(dlogs[:16])
    with pytest.raises(z.ZkmsmError):                      # needs the precomputed slabs
        ctx.msm_partial_range(plain.set, arr[:16], 0, 2)
    with pytest.raises(z.ZkmsmError):                      # world must be a power of two
        ctx.msm_partial_range(pts.set, arr, 0, 3)
    # G2 twin
    dl2 = dlogs[:300]
    pts2 = z.G2Points.generator_multiples(dl2, precompute=True)
    parts = [ctx.msm_partial_range(pts2.set, arr[:300], r, 2) for r in range(2)]
    assert U.g2_from_array(*ctx.combine(2, np.stack(parts))) == U.expected_from_dlogs(O.G2_GEN, dl2, sc[:300])


def test_partial_device_reports_bad_scalar_through_combine(z, ctx):
    """the stream-ordered partial cannot return ZKMSM_ERR_SCALAR_RANGE itself: the blob is poisoned and the combine
    call that meets it fails (previously the bad scalar was silently dropped)"""
    import torch
    n = 64
    dlogs = list(range(1, n + 1))
    pts = z.G1Points.generator_multiples(dlogs)
    good = z.scalars_to_array([3] * n)
    bad = good.copy()
    bad[5, 7] = 0x80000000                                  # bit 255 set
    d_good, d_bad = (torch.from_numpy(a.view(np.int32)).cuda() for a in (good, bad))
    d_parts = torch.zeros(2, 48, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.msm_partial_device(pts.set, d_good.data_ptr(), n, d_parts[0].data_ptr())
    ctx.msm_partial_device(pts.set, d_bad.data_ptr(), n, d_parts[1].data_ptr())
    with pytest.raises(z.ZkmsmError) as e:
        ctx.combine_device(d_parts.data_ptr(), 2)
    assert e.value.code == -3
    ctx.msm_partial_device(pts.set, d_good.data_ptr(), n, d_parts[1].data_ptr())
    assert U.g1_from_array(*ctx.combine_device(d_parts.data_ptr(), 2)) == O.scalar_mul(O.G1_GEN, 2 * 3 * n * (n + 1) // 2)


def test_graph_replay_matches_plain_launches(z, ctx):
    """the launch sequence is captured once into a CUDA graph and replayed while point set, size and buffers recur;
    no_graph launches kernel by kernel.  Same points; new scalars in the same buffer are picked up by the replay."""
    n = 3000
    rnd = random.Random(4242)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    pts = z.G1Points.generator_multiples(dlogs, precompute=True)
    for trial in range(3):
        sc = U.rand_scalars(rnd, n)
        exp = U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
        assert U.g1_from_array(*ctx.msm(pts.set, z.scalars_to_array(sc))) == exp      # capture, then replays
    ctx.set_option("no_graph", 1)
    try:
        assert U.g1_from_array(*ctx.msm(pts.set, z.scalars_to_array(sc))) == exp
    finally:
        ctx.set_option("no_graph", 0)
    # an out-of-range scalar is still reported on a replay
    bad = z.scalars_to_array(sc)
    bad[7, 7] |= 0x80000000
    with pytest.raises(z.ZkmsmError):
        ctx.msm(pts.set, bad)
    assert U.g1_from_array(*ctx.msm(pts.set, z.scalars_to_array(sc))) == exp


def test_msm_2p20_default_path_vs_oracle_closed_form(z, ctx):
    """BASELINE configs[2] size on the default path (precomputed CRS tables, batched rounds, bucket accumulation,
    CUDA-graph replay), compared with the ORACLE's scalar multiplication of the closed form (sum s_i k_i) g"""
    n = 1 << 20
    rnd = random.Random(2020)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    pts = z.G1Points.generator_multiples(dlogs, precompute=True)
    # the generated points themselves are spot-checked against the oracle's double-and-add
    for i in (0, 1, n // 2, n - 1):
        assert to_o1(pts[i]) == O.scalar_mul(O.G1_GEN, dlogs[i])
    want = O.scalar_mul(O.G1_GEN, sum(k * s for k, s in zip(dlogs, sc)) % O.R)
    arr = z.scalars_to_array(sc)
    assert U.g1_from_array(*ctx.msm(pts.set, arr)) == want
    assert U.g1_from_array(*ctx.msm(pts.set, arr)) == want             # replayed graph
    # the same through the two multi-GPU splits, all ranks on this device
    parts = [ctx.msm_partial_range(pts.set, arr, r, 8) for r in range(8)]
    assert U.g1_from_array(*ctx.combine(1, np.stack(parts))) == want


def test_scalar_longer_than_256_bits(z):
    """`&G1Point * n` takes any BigUint in the reference (macros.rs:10-21, no reduction): a 600-bit multiplier goes
    through the 256-bit device entry in chunks and equals the multiple by n mod r (the generator has order r)"""
    k = (1 << 599) + 0x123456789ABCDEF * (1 << 300) + 987654321
    g = z.G1Point.g()
    assert to_o1(g * k) == O.scalar_mul(O.G1_GEN, k % O.R)
    assert to_o1(g * (1 << 256)) == O.scalar_mul(O.G1_GEN, (1 << 256) % O.R)
    assert (z.G1Point.zero() * k).is_zero()


def test_g2_batched_affine_rounds(z, ctx):
    """G2 takes the batched-affine pre-reduction rounds as well (BatchedAddRound<G2>, Fq2 affine law of g2_point.rs /
    macros.rs:35-163): forced on a small set with every rare case inside the rounds, and the default path at 2^18"""
    g = z.G2Point.g()
    P, Q = g * 11, g * 13
    try:
        ctx.set_option("batch_rounds", 2)
        ctx.set_option("batch_T", 5)
        for pre in (False, True):
            dup = z.G2Points([g * 7] * 40, precompute=pre)
            assert U.g2_from_array(*ctx.msm(dup.set, z.scalars_to_array([5] * 40))) == O.scalar_mul(O.G2_GEN, 7 * 5 * 40)
            opp = z.G2Points([P, -P, Q, -Q, P, -P], precompute=pre)
            assert U.g2_from_array(*ctx.msm(opp.set, z.scalars_to_array([9, 9, 11, 11, 3, 3]))) == O.INF
            withinf = z.G2Points([P, z.G2Point.zero(), Q, z.G2Point.zero(), g * 17, P], precompute=pre)
            assert U.g2_from_array(*ctx.msm(withinf.set, z.scalars_to_array([5] * 6))) == O.scalar_mul(O.G2_GEN, 5 * (11 + 13 + 17 + 11))
        n = 3000
        rnd = random.Random(222)
        dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
        sc = U.rand_scalars(rnd, n)
        pts = z.G2Points.generator_multiples(dlogs, precompute=True)
        for rounds, T in ((1, 128), (3, 7)):
            ctx.set_option("batch_rounds", rounds)
            ctx.set_option("batch_T", T)
            assert U.g2_from_array(*ctx.msm(pts.set, z.scalars_to_array(sc))) == U.expected_from_dlogs(O.G2_GEN, dlogs, sc)
    finally:
        ctx.set_option("batch_rounds", -1)
        ctx.set_option("batch_T", 0)
    n = 1 << 18                            # three rounds by default from here on (fewer do not pay for G2)
    rnd = random.Random(217)
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    pts = z.G2Points.generator_multiples(dlogs, precompute=True)
    ctx.profile(True)
    out, inf = ctx.msm(pts.set, z.scalars_to_array(sc))
    names = [nm for nm, _, _ in ctx.profile_read()]
    ctx.profile(False)
    assert "batched_add_first" in names and "fixup_direct" in names
    assert U.g2_from_array(out, inf) == O.scalar_mul(O.G2_GEN, sum(k * s for k, s in zip(dlogs, sc)) % O.R)


def test_subgroup_check_at_load(z, ctx):
    """ZKMSM_CHECK_SUBGROUP: a ZKMSM_SUBGROUP set is the caller's claim that every point has order r; with the check
    the device verifies it and refuses a curve point outside the subgroup (for which the folded scalars would give a
    different point than the reference's raw multiple)"""
    from tests.test_host_emu_msm import _g1_point_outside_subgroup
    bad = _g1_point_outside_subgroup()
    good = [O.scalar_mul(O.G1_GEN, k) for k in (3, 5, 7)]
    xy, inf = U.g1_points_to_array(good)
    pts = ctx.load_points(1, xy, None, precompute=True, in_subgroup=True, check_subgroup=True)
    assert U.g1_from_array(*ctx.msm(pts, z.scalars_to_array([O.R - 1, 2, 3]))) == O.msm(good, [O.R - 1, 2, 3])
    xy, inf = U.g1_points_to_array(good + [bad])
    with pytest.raises(z.ZkmsmError) as e:
        ctx.load_points(1, xy, None, precompute=True, in_subgroup=True, check_subgroup=True)
    assert e.value.code == -7
    # without the subgroup claim nothing is assumed and the raw multiple comes out, as in the reference
    plain = ctx.load_points(1, xy, None)
    sc = [5, 6, 7, (1 << 200) + 12345]
    assert U.g1_from_array(*ctx.msm(plain, z.scalars_to_array(sc))) == O.msm(good + [bad], sc)
