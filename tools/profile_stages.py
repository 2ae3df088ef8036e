#!/usr/bin/env python3
"""Per-kernel time breakdown of one MSM (CUDA events around every launch, via zkmsm_profile)."""
import os, sys, collections
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

def main():
    logns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20]
    cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    ctx = z.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    for logn in logns:
        n = 1 << logn
        modes = [m == "1" for m in os.environ.get("MODES", "0,1").split(",")]
        for pre in modes:
          for c in cs:
            if True:
                ctx.set_window(c)
                pts = ctx.points_from_scalars(GROUP, (z.G1Point if GROUP == 1 else z.G2Point).g().limbs(), rand_scalars(n, 1), precompute=pre, in_subgroup=os.environ.get("HALF", "1") == "1")
                d_sc = torch.from_numpy(rand_scalars(n, 2).view(np.int32)).cuda()
                torch.cuda.synchronize()
                ctx.profile(False)
                for _ in range(3):
                    ctx.msm_enqueue(pts, d_sc.data_ptr(), n); ctx.msm_result(GROUP)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(5):
                    ctx.msm_enqueue(pts, d_sc.data_ptr(), n)
                e1.record(stream)
                ctx.msm_result(GROUP)
                total = e0.elapsed_time(e1) / 5
                ctx.profile(True)
                ctx.msm_enqueue(pts, d_sc.data_ptr(), n); ctx.msm_result(GROUP)
                rows = ctx.profile_read()
                agg = collections.OrderedDict()
                for name, ms, thr in rows:
                    a = agg.setdefault(name, [0.0, 0, 0])
                    a[0] += ms; a[1] += 1; a[2] = max(a[2], thr)
                ssum = sum(v[0] for v in agg.values())
                print(f"== n=2^{logn} precomp={pre} half={os.environ.get('HALF', '1')} c={c} L={os.environ.get('ZKMSM_L','-')} K={os.environ.get('ZKMSM_K','-')}: {total:.3f} ms/MSM ({n/total/1e3:.1f} Mpts/s), sum of kernels {ssum:.3f} ms, {len(rows)} launches")
                for name, (ms, cnt, thr) in agg.items():
                    print(f"   {name:18s} {ms:8.3f} ms  {100*ms/ssum:5.1f}%  x{cnt}  max_threads={thr}")
                if os.environ.get("DETAIL"):
                    print("   per launch:", " ".join(f"{nm[:6]}:{ms*1e3:.0f}us/{thr}" for nm, ms, thr in rows))
                sys.stdout.flush()
                ctx.set_window(0)
                ctx.profile(False)
                pts.free()

main()
