"""CPU: the C-ABI shared library loads, exports every symbol include/crx.h declares (and nothing is
declared that is not exported), the host-only helpers agree with the oracle, and compute entry points
fail loudly without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from crypto_recommendation_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from crypto_recommendation_b200 import build
    return build.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "crx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crx_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_matches_binding(built):
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_symbol(built):
    lib = ctypes.CDLL(built)
    for s in header_symbols():
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (crx_[A-Za-z0-9_]+)", out))
    assert set(header_symbols()) <= exported


def test_library_is_sm100a_only(built):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_torch_types_in_signatures():
    src = open(os.path.join(ROOT, "include", "crx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in src and "at::" not in src and "std::" not in src


def test_hamming_host_helper(built, port, golden):
    i = 0
    while "kat_hamming_args_%d" % i in golden:
        a = golden["kat_hamming_args_%d" % i].tolist()
        assert capi.get_num_hamming_dist_from(*a) == golden["kat_hamming_out_%d" % i].tolist() == port.hamming(*a)
        i += 1
    for a in [(0, 1, 0, 16), (65535, 2, 0, 16), (77, 3, 2, 9), (5, 0, 0, 4), (5, 9, 0, 4)]:
        assert capi.get_num_hamming_dist_from(*a) == port.hamming(*a)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present; the failure path is exercised on CPU-only hosts")
    with pytest.raises(capi.CrxError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "crypto_recommendation_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                s = open(os.path.join(root, f)).read()
                assert "oracle" not in s.replace("no oracle", "").replace("or calls oracle/", "") or f in ("synth.py",), (root, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read()
