"""Build tests/cpp/dropin_driver.cpp against the REFERENCE headers (/root/reference) and run it on a seeded
synthetic input: writes tests/golden/dropin_input.bin and tests/golden/dropin_expected.txt.
Build container only (needs /root/reference and g++)."""
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crypto_recommendation_b200 import synth  # noqa: E402

REF = os.environ.get("CRX_REF_DIR", "/root/reference")
G = os.path.join(ROOT, "tests", "golden")


def write_input(path, n=260, d=100, seed=77):
    # continuous (tie-free) vectors with a random ~90% of the coins unknown
    X = synth.normal_points(n, d, seed=seed, dtype=np.float64)
    unk = (np.random.default_rng(seed).random((n, d)) < 0.9).astype(np.uint8)
    mean = np.where(unk == 0, X, 0).sum(1) / np.maximum(1, (unk == 0).sum(1))
    with open(path, "wb") as f:
        f.write(struct.pack("<qi", n, d))
        f.write(np.ascontiguousarray(X).tobytes()); f.write(unk.tobytes()); f.write(np.ascontiguousarray(mean).tobytes())


if __name__ == "__main__":
    inp = os.path.join(G, "dropin_input.bin")
    write_input(inp)
    exe = "/tmp/dropin_ref"
    subprocess.check_call(["g++", "-O2", "-w", "-std=c++14", "-ffp-contract=off", "-DCRX_REFERENCE_BUILD", "-I" + REF,
                           os.path.join(ROOT, "tests", "cpp", "dropin_driver.cpp"), os.path.join(REF, "lib", "utils.cpp"),
                           os.path.join(REF, "lib", "data_structures", "tweet.cpp"), "-o", exe])
    subprocess.check_call([exe, inp, os.path.join(G, "dropin_expected.txt")])
    print("wrote", os.path.join(G, "dropin_expected.txt"))
