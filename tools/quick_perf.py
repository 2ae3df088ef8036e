#!/usr/bin/env python3
"""Development probe (not the bench): IMAD throughput and a first MSM timing sweep."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zk_toolkit_b200 as z

ctx = z.Context(0)

R = z.R
def rand_scalars(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x3FFFFFFF   # < 2^254 < r
    return a

for logn in (16, 18, 20, 22):
    n = 1 << logn
    dl = rand_scalars(n, 1)
    t0 = time.time()
    for pre in (False, True):
        pts = ctx.points_from_scalars(1, z.G1Point.g().limbs(), dl, precompute=pre)
        torch.cuda.synchronize()
        tgen = time.time() - t0
        sc = rand_scalars(n, 2)
        d_sc = torch.from_numpy(sc.view(np.int32)).cuda()
        stream = torch.cuda.Stream()
        torch.cuda.synchronize()
        ctx.set_stream(stream.cuda_stream)
        for cc in ((0,) if pre else (0, 13, 14, 15, 16, 17)):
            if logn != 20 and cc: continue
            ctx.set_window(cc)
            if pre and cc: continue
            for _ in range(2):
                ctx.msm_enqueue(pts, d_sc.data_ptr(), n); ctx.msm_result(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            K = 5
            for _ in range(K):
                ctx.msm_enqueue(pts, d_sc.data_ptr(), n)
            e1.record(stream)
            out, inf = ctx.msm_result(1)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            print(f"n=2^{logn} precomp={pre} c={cc}: {ms:.3f} ms  {n/ms/1e3:.1f} Mpts/s  launches={ctx.last_launch_count()} gen={tgen:.2f}s", flush=True)
        ctx.set_window(0)
        ctx.set_stream(None)
        pts.free()
