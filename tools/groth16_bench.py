#!/usr/bin/env python3
"""Groth16 prove time on the synthetic instance of BASELINE.json configs[4] (default n = 2^18), one GPU, with the
time of each of the three MSMs run alone beside it."""
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
times = []
for it in range(6):
    t0 = time.perf_counter()
    proof = inst["prover"].prove(inst["crs"], r, s)
    times.append(time.perf_counter() - t0)
a, b, c = S.expected_dlogs(inst, r, s)
ok = (proof.A == z.G1Point.g() * a) and (proof.B == z.G2Point.g() * b) and (proof.C == z.G1Point.g() * c)
parts = {}
for world in (2, 4, 8):
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        inst["prover"].prove_partial(inst["crs"], r, s, world - 1, world)
        ts.append(time.perf_counter() - t0)
    parts[f"share_1_of_{world}_ms"] = round(min(ts[1:]) * 1e3, 3)
print(json.dumps({"metric": "groth16_prove_ms", "n_constraints": n, "n_witness": n, "value": round(min(times[1:]) * 1e3, 3),
                  "unit": "ms", "all_ms": [round(t * 1e3, 2) for t in times], "setup_s": round(t_setup, 1),
                  "closed_form_check": bool(ok), **parts,
                  "note": "3 MSMs on three streams: A (G1, n+2), B (G2, n+2), C (G1, 3n+3); shares = one rank's part of an N-GPU proof, run here on one GPU"}))
