"""Memory-safety check of the MSM pipeline logic.  compute-sanitizer is closed on the B200 pool, so the identical
kernel bodies (tests/host_emu runs them on the CPU) are built with -fsanitize=address,undefined and driven over
awkward shapes (1 point, K = 1, L = 1, odd sizes, precomputed slabs, batched rounds): any out-of-bounds index,
misaligned access or signed overflow in the stage logic aborts the run."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "host_emu")


def _san(name):
    return subprocess.check_output(["gcc", f"-print-file-name={name}"], text=True).strip()


def test_pipeline_under_asan_ubsan(tmp_path):
    asan, ubsan = _san("libasan.so"), _san("libubsan.so")
    if not (os.path.isabs(asan) and os.path.exists(asan)):
        pytest.skip("libasan not available")
    so = os.path.join(str(tmp_path), "libemu_msm_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-shared", "-fPIC",
                           "-o", so, os.path.join(EMU, "emu_msm.cpp")])
    env = dict(os.environ, LD_PRELOAD=f"{asan}:{ubsan}", ASAN_OPTIONS="detect_leaks=0:abort_on_error=1",
               UBSAN_OPTIONS="halt_on_error=1")
    out = subprocess.run([sys.executable, os.path.join(EMU, "asan_run.py"), so], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "asan run ok: True" in out.stdout
    assert "runtime error" not in out.stderr and "AddressSanitizer" not in out.stderr
