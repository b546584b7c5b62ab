"""CPU check of the float-quotient arithmetic the fused kernels use (libsafecrypto_b200/csrc/fq_arith.cuh,
fq_host.h): tools/fq_model.cpp compiles the SAME headers with g++ and runs the kernels' stage structure with the
same twiddle entries against a schoolbook negacyclic product, for both schedules, and checks that every value
read as a float stays inside the bounds the host analysis proved.  No GPU, no oracle."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_float_quotient_model_matches_schoolbook_and_bounds(tmp_path):
    exe = str(tmp_path / "fq_model")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "libsafecrypto_b200", "csrc"),
                           os.path.join(ROOT, "tools", "fq_model.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "ALL OK" in out.stdout
    assert "MISMATCH" not in out.stdout and "BOUND VIOLATED" not in out.stdout
    # the headline parameter sets are served by this arithmetic
    for line in ("q=12289 n=512 ok=1", "q=12289 n=1024 ok=1", "q=7681 n=256 ok=1"):
        assert line in out.stdout
    # ... and their two-operand products by the degree-3 base multiplication (fq_arith.cuh: basemul4), without the
    # extra reduction pass (r0 = 0) -- three of the four served sets (q = 12289 at n = 256 too)
    assert out.stdout.count("warp-local schedule with base multiplication: r0=0") >= 4
