#!/usr/bin/env python3
"""Groth16 prove time on the synthetic instance of BASELINE.json configs[4] (default n = 2^18), one GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
import zk_toolkit_b200 as z
S = import_module("zk-toolkit_b200.synthetic")

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << logn
t0 = time.time()
inst = S.build(n, n)
t_setup = time.time() - t0
r, s = 0x1234567 % z.R, 0x7654321 % z.R
ctx = inst["crs"].ctx
calls = []
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); out = fn(*a, **k); calls.append((name, a[0].group if hasattr(a[0], "group") else a[0], round((time.perf_counter() - t) * 1e3, 3))); return out
    return w
ctx.msm = timed("msm", ctx.msm); ctx.msm_oneshot = timed("oneshot", ctx.msm_oneshot)
times = []
for it in range(4):
    calls.clear()
    t0 = time.perf_counter()
    proof = inst["prover"].prove(inst["crs"], r, s)
    times.append(time.perf_counter() - t0)
a, b, c = S.expected_dlogs(inst, r, s)
ok = (proof.A == z.G1Point.g() * a) and (proof.B == z.G2Point.g() * b) and (proof.C == z.G1Point.g() * c)
print(json.dumps({"metric": "groth16_prove_ms", "n_constraints": n, "n_witness": n, "value": round(min(times[1:]) * 1e3, 2),
                  "unit": "ms", "all_ms": [round(t * 1e3, 2) for t in times], "setup_s": round(t_setup, 1),
                  "closed_form_check": bool(ok), "gpu_calls_ms_last": calls, "note": "5 MSMs (3 G1 of n+2, 1 G2 of n+2, 1 G1 of 2n) + host marshalling of Python ints"}))
