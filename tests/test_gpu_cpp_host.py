"""Builds and runs the C++ host-side mirror's own test (tests/cpp/test_seam.cpp over include/zkmsm.hpp),
which restates the reference's G1/G2/Polynomial tests in the reference's compiled-language style."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def write_kats(path):
    g1 = json.load(open(os.path.join(G, "ref_g1.json")))
    g2 = json.load(open(os.path.join(G, "ref_g2.json")))
    lines = [f"g1_add_same_point {g1['add_same_point']['x']} {g1['add_same_point']['y']}"]
    for n, p in enumerate(g1["g_multiples"]["points"], 1):
        lines.append(f"g1_mult_{n} {p['x']} {p['y']}")
    q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    for i, c in enumerate(g1["scalar_mul_gen_pubkey"]["cases"]):
        lines.append(f"g1_pubkey_{i} {int(c['multiple']) % q} {c['x']} {c['y']}")
    a = g2["add_same_point"]
    lines.append(f"g2_add_same_point {a['x1']} {a['x0']} {a['y1']} {a['y0']}")
    for n, p in enumerate(g2["g_multiples"]["points"], 1):
        lines.append(f"g2_mult_{n} {p['x1']} {p['x0']} {p['y1']} {p['y0']}")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def build(tmp, name="test_seam"):
    exe = os.path.join(tmp, name)
    libdir = os.path.join(ROOT, "zk-toolkit_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe,
                           "-L", libdir, "-lzkmsm", f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"])
    return exe


def test_cpp_host_mirror_compiles(tmp_path):
    """no GPU needed: the header and the tests link against the C ABI"""
    build(str(tmp_path))
    build(str(tmp_path), "test_multi_device")


@pytest.mark.gpu
def test_two_contexts_in_one_process_partials_and_combine(tmp_path):
    """single-process host with one context per device (two devices when the box has them): point-split and
    bucket-range-split partials, combined on either device, against the fixed-base closed form"""
    exe = build(str(tmp_path), "test_multi_device")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK: 0 failure(s)" in out.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_passes_reference_tests(tmp_path):
    exe = build(str(tmp_path))
    kats = os.path.join(str(tmp_path), "kats.txt")
    write_kats(kats)
    out = subprocess.run([exe, kats], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK: 0 failure(s)" in out.stdout
