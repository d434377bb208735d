#!/usr/bin/env python
"""Regenerates tests/golden/main_expected.txt: the output file of the reference's own main.cpp
(oracle/_ref/recommendation_ref, built by tools/build_main_dropin.sh) on the synthetic inputs of
tools/make_main_inputs.py (defaults) with the RNG seed pinned to CRX_FAKE_SEED=5.  `Execution Time:` lines are
dropped.  Run in the build container (needs /root/reference for the build step only)."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SEED = "5"


def run_main(binary, workdir, extra_env=None, validate=False):
    """Output-file lines without the `Execution Time:` ones; with validate=True also the stdout lines of the 10-fold
    validation (main.cpp:180-183, 393-437: running |error| sums and the final MAE)."""
    env = dict(os.environ, CRX_FAKE_SEED=SEED)
    env.update(extra_env or {})
    out = os.path.join(workdir, "out_%s.txt" % os.path.basename(binary))
    cmd = [binary, "-d", "./tweets.tsv", "-o", out] + (["-validate"] if validate else [])
    r = subprocess.run(cmd, cwd=workdir, env=env, check=True, stdout=subprocess.PIPE, timeout=1200)
    with open(out) as f:
        lines = [l for l in f.read().splitlines() if not l.startswith("Execution Time:")]
    if validate:
        return lines, r.stdout.decode().splitlines()
    return lines


def make_inputs(workdir):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_main_inputs.py"), workdir], check=True,
                   stdout=subprocess.DEVNULL)


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d)
        lines, val = run_main(os.path.join(ROOT, "oracle", "_ref", "recommendation_ref"), d, validate=True)
    with open(os.path.join(ROOT, "tests", "golden", "main_expected.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(ROOT, "tests", "golden", "main_validate_expected.txt"), "w") as f:
        f.write("\n".join(val) + "\n")
    print("wrote %d + %d lines" % (len(lines), len(val)))
