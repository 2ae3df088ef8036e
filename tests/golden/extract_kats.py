#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into tests/golden/*.json.

Run in the build container only (it reads /root/reference, which does not exist on
the GPU box); the JSON it writes is committed.  Only literal test DATA (decimal
integers inside the reference's #[test] functions) is extracted, never code.

    python tests/golden/extract_kats.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
BLS = os.path.join(REF, "src/building_block/curves/bls12_381")
OUT = os.path.dirname(os.path.abspath(__file__))

DEC = r'b"(\d+)"'


def read(path):
    with open(path) as f:
        return f.read()


def fn_body(src, name):
    """Text of `fn name(...) { ... }` by brace matching."""
    m = re.search(r"fn\s+" + re.escape(name) + r"\s*(<[^>]*>)?\s*\(", src)
    assert m, name
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                return src[i:j + 1]
        j += 1


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def g1_kats():
    path = os.path.join(BLS, "g1_point.rs")
    src = read(path)
    out = {"source": "src/building_block/curves/bls12_381/g1_point.rs"}
    body = strip_comments(fn_body(src, "add_same_point"))
    xs = re.findall(DEC, body)
    out["add_same_point"] = {"x": xs[0], "y": xs[1], "lines": "223-237"}
    body = strip_comments(fn_body(src, "get_g_multiples"))
    pts = re.findall(r"Xy\s*\{\s*x:\s*" + DEC + r",\s*y:\s*" + DEC, body)
    assert len(pts) == 10
    out["g_multiples"] = {"points": [{"x": x, "y": y} for x, y in pts], "lines": "303-345"}
    body = strip_comments(fn_body(src, "scalar_mul_gen_pubkey"))
    cases = re.findall(r"Xy\s*\{\s*x:\s*" + DEC + r",\s*y:\s*" + DEC + r"\s*\},\s*multiple:\s*" + DEC, body)
    assert len(cases) == 4, len(cases)
    out["scalar_mul_gen_pubkey"] = {
        "cases": [{"x": x, "y": y, "multiple": k} for x, y, k in cases], "lines": "352-371"}
    body = strip_comments(fn_body(src, "add_different_points"))
    adds = re.findall(r"AddTestCase::new\((\d+),\s*(\d+),\s*(\d+)\)", body)
    out["add_different_points"] = {"cases": [[int(a), int(b), int(c)] for a, b, c in adds], "lines": "389-412"}
    return out


def g2_kats():
    path = os.path.join(BLS, "g2_point.rs")
    src = read(path)
    out = {"source": "src/building_block/curves/bls12_381/g2_point.rs"}
    body = strip_comments(fn_body(src, "add_same_point"))
    xs = re.findall(DEC, body)
    assert len(xs) == 4
    out["add_same_point"] = {"x1": xs[0], "x0": xs[1], "y1": xs[2], "y0": xs[3], "lines": "199-230"}
    pt_re = (r"Xy\s*\{\s*x1:\s*" + DEC + r",\s*x0:\s*" + DEC + r",\s*y1:\s*" + DEC + r",\s*y0:\s*" + DEC)
    body = strip_comments(fn_body(src, "get_g_multiples"))
    pts = re.findall(pt_re, body)
    assert len(pts) == 10
    out["g_multiples"] = {"points": [dict(zip(("x1", "x0", "y1", "y0"), p)) for p in pts], "lines": "308-350"}
    body = strip_comments(fn_body(src, "scalar_mul_gen_pubkey"))
    cases = re.findall(r"multiple:\s*" + DEC + r",\s*p:\s*&" + pt_re, body)
    assert len(cases) >= 1
    out["scalar_mul_gen_pubkey"] = {
        "cases": [dict(zip(("multiple", "x1", "x0", "y1", "y0"), c)) for c in cases], "lines": "357-403"}
    body = strip_comments(fn_body(src, "add_different_points"))
    adds = re.findall(r"AddTestCase::new\((\d+),\s*(\d+),\s*(\d+)\)", body)
    out["add_different_points"] = {"cases": [[int(a), int(b), int(c)] for a, b, c in adds], "lines": "421-444"}
    return out


def tower_kats(fname, tests):
    """Ordered assert_eq!(.., "<decimal>") strings per test function."""
    src = read(os.path.join(BLS, fname))
    out = {"source": "src/building_block/curves/bls12_381/" + fname}
    for t in tests:
        body = strip_comments(fn_body(src, t))
        out[t] = re.findall(r'assert_eq!\(\s*\w+\s*,\s*"(\d+)"\s*\)', body)
        assert out[t], (fname, t)
    return out


def params():
    src = read(os.path.join(BLS, "params.rs"))
    hexes = re.findall(r'parse_bytes\(b"([0-9a-f]+)",\s*16\)', src)
    g1 = read(os.path.join(BLS, "g1_point.rs"))
    g2 = read(os.path.join(BLS, "g2_point.rs"))
    g1h = re.findall(r'parse_bytes\(b"([0-9a-f]+)",\s*16\)', g1[: g1.index("impl G1Point")])
    g2h = re.findall(r'from_u8_slice\(b"([0-9a-f]+)"\)', g2[: g2.index("impl G2Point")])
    assert len(g1h) == 2 and len(g2h) == 4
    return {"source": "params.rs:9,14; g1_point.rs:38-47; g2_point.rs:36-46",
            "q": hexes[0], "r": hexes[1], "g1": g1h, "g2_x1_x0_y1_y0": g2h}


def main():
    data = {
        "g1": g1_kats(),
        "g2": g2_kats(),
        "fq2": tower_kats("fq2.rs", ["test_add", "test_sub", "test_mul", "test_inv", "test_reduce"]),
        "fq6": tower_kats("fq6.rs", ["test_add", "test_sub", "test_mul", "test_inv", "test_reduce"]),
        "fq12": tower_kats("fq12.rs", ["test_add", "test_sub", "test_mul", "test_inv"]),
        "params": params(),
    }
    for k, v in data.items():
        with open(os.path.join(OUT, f"ref_{k}.json"), "w") as f:
            json.dump(v, f, indent=1)
        print("wrote", f"ref_{k}.json")


if __name__ == "__main__":
    main()
