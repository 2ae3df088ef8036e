#!/usr/bin/env python3
"""What ONE rank of an N-GPU strong-scaling run does, measured on a single GPU: per-stage breakdown (CUDA events
around every launch) and whole-MSM time (graph replay and plain launches) for
  full      the whole 2^logn MSM (N = 1),
  points/N  a contiguous shard of n / N points (zkmsm_g1_msm_partial_device),
  range/N   all n scalars, 1 / N of the bucket range (zkmsm_g1_msm_partial_range_device).
usage: shard_perf.py [logn] [N,N,...]      env: OPTS="name=value,..." (zkmsm_set_option), GROUP=1|2, DETAIL=1
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zk_toolkit_b200 as z


def rand_scalars(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x3FFFFFFF
    return a


GROUP = int(os.environ.get("GROUP", "1"))


def timed(stream, fn, reps=10):
    for _ in range(3):
        fn()
    stream.synchronize()
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    worlds = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8]
    n = 1 << logn
    ctx = z.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    for kv in filter(None, os.environ.get("OPTS", "").split(",")):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    if os.environ.get("WINDOW"):
        ctx.set_window(int(os.environ["WINDOW"]))
    gen = (z.G1Point if GROUP == 1 else z.G2Point).g().limbs()
    words = 48 if GROUP == 1 else 96
    with torch.cuda.stream(stream):
        full = ctx.points_from_scalars(GROUP, gen, rand_scalars(n, 1), precompute=True, in_subgroup=True)
        d_sc = torch.from_numpy(rand_scalars(n, 2).view(np.int32)).cuda()
        d_out = torch.zeros(words, dtype=torch.int32, device="cuda")
        stream.synchronize()
        for world in worlds:
            cases = [(f"range[rank {rk}]", full, n, (lambda rk=rk: ctx.msm_partial_range_device(full, d_sc.data_ptr(), n, rk, world, d_out.data_ptr())))
                     for rk in sorted({0, world - 1})]
            shard = None
            if world > 1:
                m = n // world
                shard = ctx.points_from_scalars(GROUP, gen, rand_scalars(m, 3), precompute=True, in_subgroup=True)
                cases.append(("points", shard, m, lambda: ctx.msm_partial_device(shard, d_sc.data_ptr(), m, d_out.data_ptr())))
            for name, pts, m, fn in cases:
                info = pts.info()
                ctx.set_option("no_graph", 0)
                g_med, g_min = timed(stream, fn)
                ctx.set_option("no_graph", 1)
                p_med, p_min = timed(stream, fn)
                ctx.profile(True)
                fn(); fn()
                rows = ctx.profile_read()
                ctx.profile(False)
                ctx.set_option("no_graph", 0)
                agg = collections.OrderedDict()
                for nm, ms, thr in rows:
                    a = agg.setdefault(nm, [0.0, 0, 0])
                    a[0] += ms; a[1] += 1; a[2] = max(a[2], thr)
                ssum = sum(v[0] for v in agg.values())
                print(f"== 2^{logn} G{GROUP} {name}/{world}: c={info['c']} W={info['windows']}  graph {g_med:.3f} ms (min {g_min:.3f})  "
                      f"plain {p_med:.3f} ms (min {p_min:.3f})  sum of kernels {ssum:.3f} ms, {len(rows)} launches  "
                      f"-> {n / g_med / 1e3:.1f} Mpts/s if every rank takes this long")
                for nm, (ms, cnt, thr) in agg.items():
                    print(f"   {nm:20s} {ms:8.3f} ms  {100 * ms / ssum:5.1f}%  x{cnt}  max_threads={thr}")
                if os.environ.get("DETAIL"):
                    print("   per launch:", " ".join(f"{nm[:7]}:{ms * 1e3:.0f}us/{thr}" for nm, ms, thr in rows))
                sys.stdout.flush()
            if shard is not None:
                shard.free()


main()
