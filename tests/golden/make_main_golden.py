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


def run_main(binary, workdir, extra_env=None):
    env = dict(os.environ, CRX_FAKE_SEED=SEED)
    env.update(extra_env or {})
    out = os.path.join(workdir, "out_%s.txt" % os.path.basename(binary))
    subprocess.run([binary, "-d", "./tweets.tsv", "-o", out], cwd=workdir, env=env, check=True,
                   stdout=subprocess.DEVNULL, timeout=1200)
    with open(out) as f:
        return [l for l in f.read().splitlines() if not l.startswith("Execution Time:")]


def make_inputs(workdir):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_main_inputs.py"), workdir], check=True,
                   stdout=subprocess.DEVNULL)


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d)
        lines = run_main(os.path.join(ROOT, "oracle", "_ref", "recommendation_ref"), d)
    with open(os.path.join(ROOT, "tests", "golden", "main_expected.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote %d lines" % len(lines))
