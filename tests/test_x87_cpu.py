"""CPU: the product's emulation of the reference's x87 extended-precision dot product / cosine
(crypto_recommendation_b200/csrc/x87.cuh, compiled for the host) against real `long double` arithmetic."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_x87_emulation_matches_long_double(tmp_path):
    exe = str(tmp_path / "x87_check")
    subprocess.run(["g++", "-O2", "-std=c++14", "-ffp-contract=off", "-I", os.path.join(ROOT, "crypto_recommendation_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "x87_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "300000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    fields = r.stdout.split()
    assert fields[3] == "0" and fields[5] == "0", r.stdout
    # the emulation matters: the plain "round the dot to double, then divide" value differs from the reference in a large
    # share of the cases (by one ulp)
    assert int(fields[7]) > 30000, r.stdout
