#!/usr/bin/env python3
"""Development probe: zkmsm_fr_quotient timings (first call builds the per-n tables)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_toolkit_b200 as z

ctx = z.default_context()
rng = np.random.default_rng(5)
for logn in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "10,14,16,18,20").split(",")]:
    n = 1 << logn
    arrs = []
    for _ in range(3):
        a = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        a[:, 7] &= 0x3FFFFFFF
        arrs.append(a)
    t0 = time.perf_counter(); ctx.fr_quotient(*arrs); t1 = time.perf_counter()
    ts = []
    for _ in range(3):
        s = time.perf_counter(); ctx.fr_quotient(*arrs); ts.append(time.perf_counter() - s)
    ctx.profile(True); ctx.fr_quotient(*arrs); prof = ctx.profile_read(); ctx.profile(False)
    agg = {}
    for name, ms, _ in prof:
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
    kern = sum(v[1] for v in agg.values())
    detail = ", ".join(f"{k} x{c} {ms:.3f}" for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:6])
    line = f"n=2^{logn}: first call (tables + proof) {1e3 * (t1 - t0):.2f} ms, then {1e3 * min(ts):.2f} ms per quotient (host buffers in, host out)"
    if logn <= 13:
        os.environ["ZKMSM_QUOTIENT_SCHOOLBOOK"] = "1"
        s = time.perf_counter(); ctx.fr_quotient(*arrs); line += f"; schoolbook path {1e3 * (time.perf_counter() - s):.1f} ms"
        del os.environ["ZKMSM_QUOTIENT_SCHOOLBOOK"]
    print(line + f"; kernels {kern:.3f} ms in {len(prof)} launches ({detail})", flush=True)
