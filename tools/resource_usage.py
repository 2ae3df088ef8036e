#!/usr/bin/env python3
"""Registers, stack (spill) bytes and static shared memory of every kernel in libzkmsm.so, from
`cuobjdump --dump-resource-usage` (no GPU needed).  usage: resource_usage.py [path/to/libzkmsm.so] > profiles/...txt"""
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zk-toolkit_b200", "libzkmsm.so")
txt = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True, check=True).stdout
rows, tu = [], "?"
lines = txt.splitlines()
for i, ln in enumerate(lines):
    if ln.startswith("identifier ="):
        tu = ln.split("=")[1].strip()
    m = re.match(r"\s*Function (\S+):", ln)
    if m and i + 1 < len(lines):
        d = dict(kv.split(":") for kv in lines[i + 1].split() if ":" in kv)
        rows.append((m.group(1), tu, int(d.get("REG", 0)), int(d.get("STACK", 0)), int(d.get("SHARED", 0))))
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(so)}: {len(rows)} kernels; STACK > 0 means local-memory frames (spills or indexed local arrays)")
print(f"# {'regs':>4s} {'stack':>6s} {'smem':>6s}  translation unit     kernel")
for (name, tu, reg, stack, sh), d in sorted(zip(rows, dem), key=lambda t: (-t[0][2], t[1])):
    d = re.sub(r"zk::|\(.*", "", d)
    print(f"  {reg:4d} {stack:6d} {sh:6d}  {tu:20s} {d[:110]}")
