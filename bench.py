#!/usr/bin/env python3
"""Benchmark of the hot path: BLS12-381 G1 MSM (Polynomial::eval_with_g1_hidings) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

A step = one MSM of n = 2^logn (point, scalar) pairs (default 2^20, BASELINE.json configs[2]).
With N GPUs the SAME n terms are sharded contiguously over the ranks (strong scaling, as the
config says "2^20 sharded across 1/2/4/8 B200"); each rank reduces its shard to one partial point,
one all-gather of 48 words per rank follows, rank order sum + affine conversion on every rank.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import ctypes
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
LP_PER_G1_POINT = 48000       # 16 windows x 10 modmul x 300 limb products
BYTES_PER_G1_POINT = 128      # 96 B affine point + 32 B scalar
ACC_STAGES = ("accumulate", "batched_add_first", "batched_add")   # the kernels that add points into buckets


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synth_scalars(n, seed, lo, hi):
    """uniform in [0, r), reproducible per global index range [lo, hi): (hi-lo, 8) uint32 + python ints"""
    import numpy as np
    rnd = random.Random(seed)
    # one generator stream per 4096-term block so that any shard can be produced independently
    out = []
    blk = 4096
    for b in range(lo // blk, (hi + blk - 1) // blk):
        r = random.Random(seed * 1000003 + b)
        vals = [r.randrange(R) for _ in range(blk)]
        s, e = max(lo, b * blk) - b * blk, min(hi, (b + 1) * blk) - b * blk
        out.extend(vals[s:e])
    arr = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in out), dtype=np.uint32).reshape(-1, 8).copy()
    return arr, out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) >= 7 and r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ CPU arm
def load_cpu_ref():
    """the C restatement of the reference algorithm (oracle/c/zkt_ref.c) -- the one place bench.py runs oracle/"""
    so = os.path.join(ROOT, "oracle", "_build", "libzkt_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "c")], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(so)


def cpu_msm(lib, xy, scalars, threads):
    import numpy as np
    out = np.zeros(24, dtype=np.uint32)
    inf = ctypes.c_int(0)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    t0 = time.perf_counter()
    lib.zkt_g1_msm_ref(P(xy), None, P(scalars), ctypes.c_size_t(xy.shape[0]), threads, P(out), ctypes.byref(inf))
    return time.perf_counter() - t0, out, bool(inf.value)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ oracle checks
def oracle_g1_multiple(k):
    """(k mod r) * g by the ORACLE's restatement of the reference's double-and-add (macros.rs:2-32): the checker of
    every result this script reports (one ~255-step affine scalar multiplication in Python, ~20 ms)"""
    from oracle import zkt_oracle as O
    p = O.scalar_mul(O.G1_GEN, k % R)
    return None if p is O.INF else O.g1_to_limbs(p)


def oracle_g2_multiple(k):
    from oracle import zkt_oracle as O
    p = O.scalar_mul(O.G2_GEN, k % R)
    return None if p is O.INF else O.g2_to_limbs(p)


def same_point(res, want):
    xy, inf = res
    return inf if want is None else ((not inf) and list(xy.tolist()) == list(want))


# ------------------------------------------------------------------------------------------------ main
def main():
    # stdout carries exactly ONE JSON line: anything a library prints there (e.g. NCCL's version banner)
    # is diverted to stderr
    global _real_stdout
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--logn", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0x5EED0002)
    ap.add_argument("--no-precompute", action="store_true", help="plain point sets (per-window buckets + Horner)")
    ap.add_argument("--split", default="range", choices=["range", "points"],
                    help="N > 1: 'range' = every GPU holds the CRS tables and all scalars, owns 1/N of the bucket range; "
                         "'points' = contiguous shards of the (point, scalar) vectors")
    ap.add_argument("--cpu-sample-per-core", type=int, default=64)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--g2-logn", type=int, default=18, help="size of the G2 MSM block (0 = skip)")
    ap.add_argument("--groth16-logn", type=int, default=18, help="constraints of the Groth16 block (0 = skip)")
    ap.add_argument("--big-logn", type=int, default=24, help="size of the large-MSM block that runs when --gpus is 8 "
                                                            "(BASELINE configs[3]: 2^24 across 8 B200; 0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_total = 1 << args.logn
    workload = f"BLS12-381 G1 MSM n=2^{args.logn}: points k_i*g (k_i uniform in [1,r)), scalars uniform in [0,r), seed {args.seed:#x}"
    W = max(args.warmup, 3)

    if args.impl == "reference":
        return reference_arm(args, rank, world, n_total, workload, W)

    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_toolkit_b200 as z
    from importlib import import_module
    sharding = import_module("zk-toolkit_b200.sharding")

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = z.Context(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    ctx.set_stream(stream.cuda_stream)
    split = args.split if (world > 1 and not args.no_precompute and world & (world - 1) == 0) else "points"

    lo, hi = (0, n_total) if split == "range" and world > 1 else sharding.shard_range(n_total, rank, world)
    n = hi - lo
    t_setup = time.time()
    dlog_arr, dlogs = synth_scalars(n_total, args.seed + 1, lo, hi)
    dlog_arr = dlog_arr.copy()
    for i, k in enumerate(dlogs):          # k_i in [1, r)
        if k == 0:
            dlogs[i] = 1
            dlog_arr[i] = 0
            dlog_arr[i, 0] = 1
    sc_arr, scalars = synth_scalars(n_total, args.seed + 2, lo, hi)
    pts = ctx.points_from_scalars(1, z.G1Point.g().limbs(), dlog_arr, precompute=not args.no_precompute,
                                  in_subgroup=True)   # multiples of g have order r; scalars are < r
    expected_k = sum(k * s for k, s in zip(dlogs, scalars)) % R     # this rank's terms of sum s_i k_i
    if split == "range" and world > 1:
        expected_k = expected_k if rank == 0 else 0                 # every rank holds all terms: count them once
    log(f"[rank {rank}] terms [{lo},{hi}) ({split} split) set up in {time.time() - t_setup:.1f}s")

    with torch.cuda.stream(stream):
        h_sc = torch.from_numpy(sc_arr.view(np.int32)).pin_memory()
        d_sc = h_sc.to("cuda", non_blocking=True)
        d_partial = torch.zeros(48, dtype=torch.int32, device="cuda")
        d_slice = torch.empty(max(1, n // world), 8, dtype=torch.int32, device="cuda") if world > 1 else None
        stream.synchronize()

        def enqueue_step():
            """one MSM with inputs resident in HBM, stream-ordered (no host synchronisation); the canonical affine
            result lands in the context's device result block and is fetched by fetch()"""
            if world == 1:
                ctx.msm_enqueue(pts, d_sc.data_ptr(), n)
                return
            if split == "range":
                ctx.msm_partial_range_device(pts, d_sc.data_ptr(), n, rank, world, d_partial.data_ptr())
            else:
                ctx.msm_partial_device(pts, d_sc.data_ptr(), n, d_partial.data_ptr())
            g = sharding.gather_partials(d_partial)
            ctx.combine_enqueue(g.data_ptr(), world)
            return g

        def fetch():
            return ctx.msm_result(1)

        def e2e_step():
            """through the public call with HOST scalars: H2D copy + MSM + D2H of the result"""
            if world == 1:
                return ctx.msm_host_ptr(pts, h_sc.data_ptr(), n)
            if split == "range":
                # every rank needs all scalars: each uploads 1/N of them and the slices are all-gathered over NVLink, so
                # the vector crosses PCIe once per step, not once per GPU
                per = n // world
                d_slice.copy_(h_sc[rank * per:(rank + 1) * per], non_blocking=True)
                dist.all_gather_into_tensor(d_sc.view(-1), d_slice.view(-1))
            else:
                d_sc.copy_(h_sc, non_blocking=True)
            enqueue_step()
            return fetch()

        # ---- warm-up (the clock sampler starts here so that nvidia-smi's start-up cost is not in the timed region)
        sampler = ClockSampler(local_rank) if rank == 0 else None
        for _ in range(W):
            enqueue_step()
            res = fetch()
        # ---- check the result: sum_i s_i (k_i g) == (sum_i s_i k_i mod r) g, the right-hand side by the ORACLE
        tot_k = expected_k
        if world > 1:
            ks = [None] * world
            dist.all_gather_object(ks, expected_k)
            tot_k = sum(ks) % R
        want_xy = oracle_g1_multiple(tot_k)
        if not same_point(res, want_xy):
            raise SystemExit(f"[rank {rank}] MSM result does not match the oracle's (sum s_i k_i) * g")
        # (the generated points themselves: a sample against the oracle's scalar multiplication)
        if rank == 0:
            for i in (0, n // 2, n - 1):
                xy, inf = pts.read(i, 1)
                if not same_point((xy[0], bool(inf[0])), oracle_g1_multiple(dlogs[i])):
                    raise SystemExit("generated point differs from the oracle's k_i * g")

        # ---- timed region: EXACTLY K steps back to back on the launching stream, one CUDA-event bracket, a barrier and
        # a device synchronisation on both sides, max over ranks.  No L2 flush: every step gathers from the resident
        # CRS tables (1.4 GB at 2^20, 11x the 126 MB L2) and rewrites > 300 MB of sorted pairs and partial sums.
        time.sleep(0.5)             # let the sampler's first queries finish
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        keep = []
        for _ in range(args.steps):
            keep.append(enqueue_step())
        e1.record(stream)
        e1.synchronize()
        res = fetch()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = e0.elapsed_time(e1)
        launches = args.steps * (ctx.last_launch_count() + (2 if world > 1 else 0))
        if not same_point(res, want_xy):
            raise SystemExit(f"[rank {rank}] result of the timed steps does not match the oracle's")
        del keep
        if world > 1:
            t = torch.tensor([total_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        ms_per_step = total_ms / args.steps
        value = n_total / (ms_per_step * 1e-3) / 1e6
        log(f"[rank {rank}] {args.steps} steps in {total_ms:.3f} ms")

        # ---- the same steps once more with a CUDA-event pair around every launch (zkmsm_profile; the launches then go
        # out one by one instead of as a graph replay): stage split, duration of the dominant kernel group
        ctx.profile(True)
        enqueue_step(); fetch()      # first use of the per-launch events
        acc_ms, batch_rounds, prof_step_ms, phase_ms = [], 0, [], {}
        for _ in range(args.steps):
            p0, pa, pb, p1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            p0.record(stream)
            if world > 1:
                if split == "range":
                    ctx.msm_partial_range_device(pts, d_sc.data_ptr(), n, rank, world, d_partial.data_ptr())
                else:
                    ctx.msm_partial_device(pts, d_sc.data_ptr(), n, d_partial.data_ptr())
                prof = None
                pa.record(stream)
                g = sharding.gather_partials(d_partial)
                pb.record(stream)
                ctx.combine_enqueue(g.data_ptr(), world)
            else:
                enqueue_step()
            p1.record(stream)
            p1.synchronize()
            prof_step_ms.append(p0.elapsed_time(p1))
            if world > 1:
                phase_ms = {"shard_msm": round(p0.elapsed_time(pa), 4), "all_gather": round(pa.elapsed_time(pb), 4),
                            "combine": round(pb.elapsed_time(p1), 4)}
            fetch()
            prof = ctx.profile_read()
            if world == 1:
                acc_ms.append(sum(ms for name, ms, _ in prof if name in ACC_STAGES))
                batch_rounds = sum(1 for name, _, _ in prof if name in ("batched_add_first", "batched_add"))
        stage_profile = {}
        if world == 1:
            for name, ms, _ in prof:
                stage_profile[name] = round(stage_profile.get(name, 0.0) + ms, 4)
        ctx.profile(False)
        if world > 1:
            # the combine's profile replaced the shard's: take the shard MSM's stages from one more run of it alone
            ctx.profile(True)
            for _ in range(2):
                if split == "range":
                    ctx.msm_partial_range_device(pts, d_sc.data_ptr(), n, rank, world, d_partial.data_ptr())
                else:
                    ctx.msm_partial_device(pts, d_sc.data_ptr(), n, d_partial.data_ptr())
                prof = ctx.profile_read()
                acc_ms.append(sum(ms for name, ms, _ in prof if name in ACC_STAGES))
            batch_rounds = sum(1 for name, _, _ in prof if name in ("batched_add_first", "batched_add"))
            for name, ms, _ in prof:
                stage_profile[name] = round(stage_profile.get(name, 0.0) + ms, 4)
            ctx.profile(False)

        # ---- end to end through the public API with host buffers
        for _ in range(2):
            e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_res = e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps
        if not same_point(e2e_res, want_xy):
            raise SystemExit("end-to-end result does not match the oracle's")
        if world > 1:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": round(n_total / e2e_s / 1e6, 3), "unit": "Mpoints/s", "ms_per_step": round(e2e_s * 1e3, 4),
               "h2d_bytes_per_step": (n // world if world > 1 and split == "range" else n) * 32, "d2h_bytes_per_step": 24 * 4 + 8,
               "api": "zkmsm_g1_msm(ctx, resident CRS points, host scalars) -> host affine point" if world == 1 else
                      "per rank: H2D of 1/N of the scalars + NVLink all-gather of the slices (range split; all of its own under the points split), "
                      "zkmsm_g1_msm_partial_*_device, NCCL all-gather of the partials, zkmsm_g1_combine_enqueue + result"}
        # clocks / throttle reasons: nvidia-smi samples every 50 ms from the start of the timed steps to the end of the
        # end-to-end steps (the same workload under load throughout; K steps alone last only ~60 ms)
        clocks = sampler.stop(t_wall0, time.time()) if sampler else None
        if world == 1:
            # the same steps issued through the asynchronous pair of calls on two contexts, so that the copy of step
            # k + 1 runs under the kernels of step k (what a prover with several MSMs per proof does);
            # every step still copies its scalars from pinned host memory and reads its result back
            ctx_b = z.Context(local_rank)
            h_sc_b = h_sc.clone().pin_memory()
            lanes = [(ctx, h_sc), (ctx_b, h_sc_b)]

            def pipelined(steps):
                lanes[0][0].msm_begin_ptr(pts, lanes[0][1].data_ptr(), n)
                for k in range(1, steps):
                    lanes[k % 2][0].msm_begin_ptr(pts, lanes[k % 2][1].data_ptr(), n)
                    out = lanes[(k - 1) % 2][0].msm_result(1)
                return lanes[(steps - 1) % 2][0].msm_result(1)

            out = pipelined(3)
            if not same_point(out, want_xy):
                raise SystemExit("pipelined end-to-end result differs")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipelined(args.steps)
            torch.cuda.synchronize()
            pip_s = (time.perf_counter() - t0) / args.steps
            e2e["pipelined"] = {"value": round(n_total / pip_s / 1e6, 3), "ms_per_step": round(pip_s * 1e3, 4),
                                "api": "zkmsm_g1_msm_begin / zkmsm_g1_msm_result alternating on two contexts"}
            ctx_b.close()

    # ---- roofline of the dominant kernel (bucket accumulation), measured live above
    # Denominator: a limb product (32x32->64 multiply-accumulate) is one IMAD.WIDE, which issues at HALF the
    # rate of a 32-bit IMAD on sm_100 (measured: a 300-LP modmul takes 1237 cycles per warp per SMSP, see
    # DESIGN.md); the carry-chain probe (the Montgomery inner loop's instruction) is measured beside it and the
    # LARGER of the two is the peak.
    lp_peak, probe = None, {}
    if rank == 0:
        for v, name in ((2, "imad32"), (1, "imad_wide_x_carry_chain"), (3, "imad_wide_x_dependent")):
            try:
                lp, ms = ctx.bench_imad(v, 8192)
                probe[name] = round(lp / 1e12, 3)
            except Exception:
                pass
        if "imad32" in probe:
            lp_peak = max(probe["imad32"] / 2, probe.get("imad_wide_x_carry_chain", 0.0)) * 1e12
    acc_avg_ms = sum(acc_ms) / len(acc_ms)
    n_share = n_total / world        # terms' worth of accumulation work on this rank
    achieved = n_share * LP_PER_G1_POINT / (acc_avg_ms * 1e-3) if acc_avg_ms > 0 else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # DRAM bytes per launch of the dominant kernel group: from the committed ncu --set full capture of the
    # same command (profiles/accumulate_traffic.json, written by tools/summarize_profiles.py), never guessed
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "accumulate_traffic.json")))
        if tr.get("n") == n_total and tr.get("n_gpus", 1) == world:
            traffic = tr
    except Exception:
        pass
    info = pts.info()
    roofline = {
        "bound": "int32 multiplier (IMAD.WIDE on the fma pipe); not hbm, not tensor",
        "kernel": "bucket accumulation = BatchedAddRound<G1> x rounds (affine adds, shared inversion) + AccumulateBuckets<G1> "
                  "(XYZZ mixed adds); one 'launch' is this group, durations summed",
        "achieved": round(achieved / 1e12, 3) if achieved else None,
        "peak": round(lp_peak / 1e12, 3) if lp_peak else None,
        "unit": "T limb-products/s (32x32->64 multiply-accumulate)",
        "frac": None,
        "canonical_frac": round(achieved / lp_peak, 4) if achieved and lp_peak else None,
        "peak_source": "measured live: max(zkmsm_bench_imad 32-bit IMAD rate / 2, carry-chained IMAD.WIDE.X rate); MEASURED_PEAKS.json has no integer figure",
        "frac_of_imad32_issue_rate": round(achieved / (2 * probe["imad32"] * 1e12), 4) if achieved and probe.get("imad32") else None,
        "probe_T_per_s": probe,
        "kernel_ms": round(acc_avg_ms, 4),
        "kernel_ms_source": "CUDA events around every launch of the group, second pass of the same K steps (zkmsm_profile)",
        "kernel_share_of_step": round(acc_avg_ms / (sum(prof_step_ms) / len(prof_step_ms)), 4),
        "step_frac_canonical": round(n_total * LP_PER_G1_POINT / (ms_per_step * 1e-3) / lp_peak / world, 4) if lp_peak else None,
        "traffic": traffic,
        "actual_lp_per_point": None,
        "hbm": {"achieved_GBps": round(n_total * BYTES_PER_G1_POINT / (ms_per_step * 1e-3) / 1e9, 2),
                "peak_GBps": peaks.get("hbm_gbs"), "note": "algorithmic 128 B/point; the path is multiplier-bound"},
        "stages_ms_profiled_step": stage_profile,
        "profiled_step_ms": round(sum(prof_step_ms) / len(prof_step_ms), 4),
    }
    if info["precomputed"]:
        # what the kernels really multiply (fewer windows than the canonical 16 with precomputed tables; a batched
        # affine addition is 6 modmul, an XYZZ mixed addition 10, 300 LP each): after R halving rounds a fraction
        # 2^-R of the entries is left for the mixed additions.
        left = 0.5 ** batch_rounds
        roofline["actual_lp_per_point"] = round(info["windows"] * ((1 - left) * 1800 + left * 3000))
        roofline["batched_affine_rounds"] = batch_rounds
        roofline["window_bits"] = info["c"]
        if achieved and lp_peak:
            # `frac` = limb products the group really ISSUES / its time / peak (the multiplier-pipe utilisation ncu
            # reports as sm__pipe_fmaheavy_cycles_active, profiles/); `canonical_frac` counts SURVEY.md 8(d)'s
            # 48 000 LP per point whatever the algorithm executes and so can exceed 1
            ex = achieved * roofline["actual_lp_per_point"] / LP_PER_G1_POINT
            roofline["achieved_executed"] = round(ex / 1e12, 3)
            roofline["frac"] = round(ex / lp_peak, 4)
            roofline["frac_note"] = ("frac = executed limb products / kernel time / peak; canonical_frac uses the fixed 48 000 LP per "
                                     "point of SURVEY.md 8(d) (precomputed tables + batched-affine additions execute fewer)")
    elif achieved and lp_peak:
        roofline["frac"] = roofline["canonical_frac"]

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference algorithm on a bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cores = host_cores()
        m = min(n, cores * args.cpu_sample_per_core)
        xy, _ = pts.read(0, m)
        lib = load_cpu_ref()
        dt, cpu_xy, cpu_inf = cpu_msm(lib, xy, sc_arr[:m].copy(), cores)
        gpu_xy, gpu_inf = ctx.msm(pts, sc_arr[:m], n=m)
        if cpu_inf != gpu_inf or (not cpu_inf and cpu_xy.tolist() != gpu_xy.tolist()):
            raise SystemExit("GPU and CPU-reference MSM differ on the baseline sample")
        cpu_baseline = {"value": round(m / dt / 1e6, 9), "unit": "Mpoints/s", "cores": cores, "kind": "port",
                        "sample": f"first {m} (point, scalar) pairs of the workload, {dt:.2f} s; result bit-identical to the GPU's; "
                                  "the reference's cost is linear in n, so the figure extrapolates to 2^20",
                        "points_per_s": round(m / dt, 2)}

    # ---- the metric's other halves: G2 MSM throughput and Groth16 prove ms (extra keys; `value` stays the G1 MSM)
    pts.free()
    del d_sc
    torch.cuda.empty_cache()
    g2_block = g2_bench(args, ctx, stream, rank, world, lp_peak) if args.g2_logn else None
    groth16_block = groth16_bench(args, ctx, rank, world) if args.groth16_logn else None
    big_block = None
    if args.big_logn and world == 8 and not args.no_precompute:
        try:
            big_block = big_msm_bench(args, ctx, stream, rank, world)
        except z.ZkmsmError as e:          # e.g. not enough device memory next to another tenant: reported, not fatal
            big_block = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "bls12_381_g1_msm_mpoints_per_s", "value": round(value, 3), "unit": "Mpoints/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x12 (381-bit Montgomery integers, exact)", "data": "synthetic",
            "config": {"workload": workload, "n": n_total, "per_gpu": n, "precomputed_crs_tables": not args.no_precompute,
                       "multi_gpu_split": ("bucket range: every GPU holds the CRS tables and all scalars, owns 1/N of the buckets"
                                           if split == "range" and world > 1 else "contiguous shards of the (point, scalar) vectors")
                                          if world > 1 else None,
                       "l2": "not flushed: inputs exceed the 126 MB L2 (each step gathers from 1.4 GB of resident CRS tables and "
                             "rewrites > 300 MB of sorted pairs / partial sums); steps run back to back on one stream",
                       "result_check": "sum s_i*(k_i g) == (sum s_i k_i mod r) g with the right-hand side computed by the ORACLE "
                                       "(oracle/zkt_oracle.py scalar_mul), before timing and again on the timed steps' result"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
            "g2": g2_block, "groth16": groth16_block, "msm_big": big_block,
        }
        if world > 1:
            line["multi_gpu"] = {"exchange": "one all-gather of 48 words per rank (NCCL), rank-order sum + affine on every rank",
                                 "phase_ms_profiled_step": phase_ms}
        print(json.dumps(line), file=_real_stdout, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def g2_bench(args, ctx, stream, rank, world, lp_peak):
    """G2 MSM (Polynomial::eval_with_g2_hidings, polynomial.rs:284-293) at 2^g2_logn, same protocol as the headline:
    K steps back to back, inputs resident, oracle-checked; N > 1 splits the bucket range."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_toolkit_b200 as z
    from importlib import import_module
    sharding = import_module("zk-toolkit_b200.sharding")
    n = 1 << args.g2_logn
    dlog_arr, dlogs = synth_scalars(n, args.seed + 11, 0, n)
    sc_arr, scalars = synth_scalars(n, args.seed + 12, 0, n)
    use_range = world > 1 and world & (world - 1) == 0
    if world > 1 and not use_range:
        return None
    with torch.cuda.stream(stream):
        pts = ctx.points_from_scalars(2, z.G2Point.g().limbs(), dlog_arr, precompute=True, in_subgroup=True)
        d_sc = torch.from_numpy(sc_arr.view(np.int32)).cuda()
        d_partial = torch.zeros(96, dtype=torch.int32, device="cuda")
        stream.synchronize()

        def enqueue():
            if world == 1:
                ctx.msm_enqueue(pts, d_sc.data_ptr(), n)
                return None
            ctx.msm_partial_range_device(pts, d_sc.data_ptr(), n, rank, world, d_partial.data_ptr())
            g = sharding.gather_partials(d_partial)
            ctx.combine_enqueue(g.data_ptr(), world, group=2)
            return g

        for _ in range(3):
            enqueue()
            res = ctx.msm_result(2)
        want = oracle_g2_multiple(sum(k * s for k, s in zip(dlogs, scalars)))
        if not same_point(res, want):
            raise SystemExit("G2 MSM result does not match the oracle's (sum s_i k_i) * g2")
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        keep = [enqueue() for _ in range(args.steps)]
        e1.record(stream)
        e1.synchronize()
        res = ctx.msm_result(2)
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if not same_point(res, want):
            raise SystemExit("G2 MSM timed result does not match the oracle's")
        info = pts.info()
        pts.free()
    lp = 134400   # 16 windows x 28 Fq modmul x 300 limb products per G2 point
    return {"metric": "bls12_381_g2_msm_mpoints_per_s", "n": n, "value": round(n / (ms * 1e-3) / 1e6, 3), "unit": "Mpoints/s",
            "ms_per_step": round(ms, 4), "window_bits": info["c"],
            "roofline": {"bound": "int32 multiplier", "achieved": round(n * lp / (ms * 1e-3) / 1e12, 3),
                         "peak": round(lp_peak / 1e12, 3) if lp_peak else None, "unit": "T limb-products/s",
                         "canonical_step_frac": round(n * lp / (ms * 1e-3) / lp_peak, 4) if lp_peak else None,
                         "note": "whole step over the canonical 134 400 LP per G2 point"},
            "result_check": "oracle closed form (sum s_i k_i) * g2"}


def big_msm_bench(args, ctx, stream, rank, world):
    """BASELINE configs[3]: G1 MSM of 2^big_logn terms across the 8 GPUs (bucket-range split: every rank holds the
    whole precomputed table -- 19 GiB at 2^24 -- and all scalars).  Same protocol as the headline; points k_i g and
    scalars from a counter-based generator (numpy) so that 2^24 of them take seconds, not minutes; checked against
    the ORACLE through the closed form (sum s_i k_i) g."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_toolkit_b200 as z
    from importlib import import_module
    sharding = import_module("zk-toolkit_b200.sharding")
    n = 1 << args.big_logn

    def limbs(seed):   # uniform 254-bit values (< r), reproducible on every rank
        a = np.random.default_rng(seed).integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        a[:, 7] &= 0x3FFFFFFF
        return a

    def to_int_sum_of_products(a, b):
        """sum a_i b_i mod r from limb arrays, exact: 16 x 16 dot products of the 16-bit half-limb columns (each term
        < 2^32, each sum of 2^24 terms < 2^56: exact in uint64), recombined with Python integers"""
        a16 = np.ascontiguousarray(a.view(np.uint16).astype(np.uint64).T)     # (16, n)
        b16 = np.ascontiguousarray(b.view(np.uint16).astype(np.uint64).T)
        tot = 0
        for j in range(16):
            for k in range(16):
                tot += int(np.dot(a16[j], b16[k])) << (16 * (j + k))
        return tot % R

    t0 = time.time()
    dl, sc = limbs(args.seed + 21), limbs(args.seed + 22)
    dl[:, 0] |= 1                                    # k_i != 0
    with torch.cuda.stream(stream):
        pts = ctx.points_from_scalars(1, z.G1Point.g().limbs(), dl, precompute=True, in_subgroup=True)
        d_sc = torch.from_numpy(sc.view(np.int32)).cuda()
        d_partial = torch.zeros(48, dtype=torch.int32, device="cuda")
        stream.synchronize()
        setup_s = time.time() - t0

        def enqueue():
            ctx.msm_partial_range_device(pts, d_sc.data_ptr(), n, rank, world, d_partial.data_ptr())
            g = sharding.gather_partials(d_partial)
            ctx.combine_enqueue(g.data_ptr(), world)
            return g

        for _ in range(3):
            enqueue()
            res = ctx.msm_result(1)
        want = oracle_g1_multiple(to_int_sum_of_products(dl, sc)) if rank == 0 else None
        if rank == 0 and not same_point(res, want):
            raise SystemExit("2^%d MSM result does not match the oracle's closed form" % args.big_logn)
        dist.barrier()
        torch.cuda.synchronize()
        steps = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        keep = [enqueue() for _ in range(steps)]
        e1.record(stream)
        e1.synchronize()
        res2 = ctx.msm_result(1)
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if res2[0].tolist() != res[0].tolist():
            raise SystemExit("2^%d MSM: timed result differs" % args.big_logn)
        info = pts.info()
        pts.free()
    return {"metric": "bls12_381_g1_msm_mpoints_per_s", "n": n, "n_gpus": world, "value": round(n / (ms * 1e-3) / 1e6, 3),
            "unit": "Mpoints/s", "ms_per_step": round(ms, 4), "steps": steps, "window_bits": info["c"], "windows": info["windows"],
            "split": "bucket range (striped), CRS tables replicated", "setup_s": round(setup_s, 1),
            "result_check": "oracle closed form (sum s_i k_i) * g on rank 0"}


def groth16_bench(args, ctx, rank, world):
    """Groth16 prove (Prover::prove, prover.rs:96-147) on the synthetic instance of BASELINE.json configs[4]
    (n constraints, n witness wires), end to end through the API: the coefficient vectors come
    from pinned HOST memory every proof and the proof is read back.  N > 1: ONE proof over the N GPUs."""
    import torch
    import torch.distributed as dist
    from importlib import import_module
    S = import_module("zk-toolkit_b200.synthetic")
    G = import_module("zk-toolkit_b200.groth16")
    n = 1 << args.groth16_logn
    use_dist = world > 1 and world & (world - 1) == 0
    if world > 1 and not use_dist:
        return None
    t0 = time.time()
    inst = S.build(n, n, ctx=ctx)
    setup_s = time.time() - t0
    r, s = 0x1234567 % R, 0x7654321 % R
    prove = (lambda: G.prove_distributed(inst["prover"], inst["crs"], r, s)) if use_dist else (lambda: inst["prover"].prove(inst["crs"], r, s))
    for _ in range(3):
        proof = prove()
    a, b, c = S.expected_dlogs(inst, r, s)
    ok = (same_point((proof.A.limbs(), proof.A.is_zero()), oracle_g1_multiple(a)) and
          same_point((proof.B.limbs(), proof.B.is_zero()), oracle_g2_multiple(b)) and
          same_point((proof.C.limbs(), proof.C.is_zero()), oracle_g1_multiple(c)))
    if not ok:
        raise SystemExit("Groth16 proof does not match its closed form (oracle scalar multiplications)")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    times, split = [], {}
    for _ in range(args.steps):
        t0 = time.perf_counter()
        p2 = prove()
        times.append((time.perf_counter() - t0) * 1e3)
    if use_dist:
        G.prove_distributed(inst["prover"], inst["crs"], r, s, timings=split)
        split = {k: round(v, 3) for k, v in split.items()}
    if (p2.A, p2.B, p2.C) != (proof.A, proof.B, proof.C):
        raise SystemExit("Groth16 proofs differ between runs")
    ms = sum(times) / len(times)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    inst["crs"].free()
    return {"metric": "groth16_prove_ms", "n_constraints": n, "n_witness": n, "prove_ms": round(ms, 3), "min_ms": round(min(times), 3),
            "unit": "ms", "higher_is_better": False, "n_gpus": world, "setup_s": round(setup_s, 1),
            "msms": "A: G1 n+2 | B: G2 n+2 | C: G1 3n+3 (s A + r B_g1 - r s delta folded into C's scalars), three streams",
            "h2d_bytes_per_proof": 4 * n * 32 // (world if use_dist else 1), "d2h_bytes_per_proof": 96 * 4, "phase_ms_one_proof": split or None,
            "api": "zkmsm_groth16_prove" if not use_dist else "zkmsm_groth16_prove_partial per rank + NCCL all-gather + zkmsm_groth16_combine",
            "result_check": "A, B, C equal their closed-form discrete logs times the generators, right-hand sides by the ORACLE"}


def reference_arm(args, rank, world, n_total, workload, W):
    """The reference's own algorithm (serial affine double-and-add with an extended-Euclid inversion per
    group operation) on the host cores.  The Rust crate cannot be built in this image, so this is the
    C restatement oracle/c/zkt_ref.c, validated against the reference's golden vectors."""
    if rank != 0:
        return
    import numpy as np
    cores = host_cores()
    m = min(n_total, cores * args.cpu_sample_per_core)
    lib = load_cpu_ref()
    # same workload definition: points k_i*g.  Built here with the reference algorithm itself (untimed).
    dl_arr, _ = synth_scalars(n_total, args.seed + 1, 0, m)
    sc_arr, _ = synth_scalars(n_total, args.seed + 2, 0, m)
    gen = np.array([(v >> (32 * i)) & 0xFFFFFFFF for v in (
        0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
        0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
        for i in range(12)], dtype=np.uint32)
    xy = np.zeros((m, 24), dtype=np.uint32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)

    def gen_points(lo, hi):
        inf = ctypes.c_int(0)
        for i in range(lo, hi):
            lib.zkt_g1_mul_ref(P(gen), 0, P(dl_arr[i]), P(xy[i]), ctypes.byref(inf))

    ths = [threading.Thread(target=gen_points, args=(m * t // cores, m * (t + 1) // cores)) for t in range(cores)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    times = []
    for it in range(W + args.steps):
        dt, out, inf = cpu_msm(lib, xy, sc_arr, cores)
        if it >= W:
            times.append(dt)
    s_per_step = sum(times) / len(times)
    value = m / s_per_step / 1e6
    line = {
        "impl": "reference", "metric": "bls12_381_g1_msm_mpoints_per_s", "value": round(value, 9), "unit": "Mpoints/s",
        "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": round(s_per_step * 1e3, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "arbitrary-precision integers (exact)",
        "data": "synthetic",
        "config": {"workload": workload, "n": n_total, "sample_per_step": m,
                   "note": "throughput is independent of n (the algorithm is a serial sum of per-term scalar multiplications)"},
        "cpu_baseline": {"value": round(value, 9), "unit": "Mpoints/s", "cores": cores, "kind": "port",
                         "sample": f"{m} (point, scalar) pairs per step; reference algorithm restated in C (no Rust toolchain in the image)"},
        "e2e": {"value": round(value, 9), "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_real_stdout, flush=True)


if __name__ == "__main__":
    main()
