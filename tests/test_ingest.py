"""binary columnar ingest in place of VectorReader's getline + stod (vector_reader.hpp:55-85).
CPU: the file format (converter tool == the ABI's writer, byte for byte; ids and shape round-trip; damaged files are
refused).  GPU: the transposed rows are the converted values bit for bit, as host rows and as a point set, and the
reference's main.cpp over the drop-in reader with a sidecar next to its input writes the expected output."""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import csv_to_columnar  # noqa: E402

from crypto_recommendation_b200 import capi  # noqa: E402


def write_csv(path, ids, X, delimiter=",", crlf=False):
    with open(path, "w", newline="") as f:
        for i, row in zip(ids, X):
            f.write(delimiter.join([i] + [repr(float(v)) for v in row]) + ("\r\n" if crlf else "\n"))


def sample(n=257, d=37, seed=1):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)) * 10.0 ** rng.integers(-8, 8, size=(n, d))
    X[rng.random((n, d)) < 0.3] = 0.0
    ids = ["v%d" % (7 * i + 3) for i in range(n)]
    return ids, X


def test_converter_and_abi_writer_agree(tmp_path):
    ids, X = sample()
    csv = str(tmp_path / "in.csv")
    write_csv(csv, ids, X, crlf=True)
    a, b = str(tmp_path / "a.crxcol"), str(tmp_path / "b.crxcol")
    assert csv_to_columnar.convert(csv, a) == X.shape
    capi.Columnar.write(b, ids, X)
    assert open(a, "rb").read() == open(b, "rb").read()
    f = capi.Columnar(a)
    assert (f.n, f.d) == X.shape and f.ids() == ids
    f.close()


def test_converter_follows_the_readers_line_rules(tmp_path):
    # vector_reader.hpp:67-80: metadata lines in front are skipped, '\r' dropped, the id is the text before the first delimiter,
    # a trailing delimiter adds no coordinate (getline-based split), any one-character delimiter
    p = str(tmp_path / "t.tsv")
    with open(p, "w", newline="") as f:
        f.write("meta line one\nmeta line two\n")
        f.write("a\t1.5\t-2e-3\t7\t\r\n")
        f.write("id with spaces\t0.1\t0.2\t1e308\n")
        f.write("c\t4.9e-324\t-0.0\t3\t\n")
    out = str(tmp_path / "t.crxcol")
    assert csv_to_columnar.convert(p, out, delimiter="\t", skip_lines=2) == (3, 3)
    f = capi.Columnar(out)
    assert f.ids() == ["a", "id with spaces", "c"] and (f.n, f.d) == (3, 3)
    f.close()
    want = np.array([[1.5, -2e-3, 7.0], [0.1, 0.2, 1e308], [4.9e-324, -0.0, 3.0]])
    blob = open(out, "rb").read()
    cols = np.frombuffer(blob[-3 * 3 * 8:], np.float64).reshape(3, 3)
    assert np.array_equal(cols.T.view(np.uint64), want.view(np.uint64))     # bit for bit, -0.0 and the subnormal included
    ragged = str(tmp_path / "r.csv")
    open(ragged, "w").write("a,1,2\nb,1\n")
    with pytest.raises(SystemExit):
        csv_to_columnar.convert(ragged, str(tmp_path / "r.crxcol"))


def test_damaged_files_are_refused(tmp_path):
    ids, X = sample(40, 5)
    good = str(tmp_path / "g.crxcol")
    capi.Columnar.write(good, ids, X)
    blob = open(good, "rb").read()
    for name, data in (("short", blob[:-8]), ("magic", b"X" + blob[1:]), ("tiny", blob[:16]), ("long", blob + b"\0" * 8)):
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        with pytest.raises(capi.CrxError):
            capi.Columnar(p)
    with pytest.raises(capi.CrxError):
        capi.Columnar(str(tmp_path / "absent"))


@pytest.mark.gpu
def test_rows_and_points_are_the_converted_values(ctx, port, tmp_path):
    ids, X = sample(5003, 203, seed=2)
    csv = str(tmp_path / "in.csv")
    write_csv(csv, ids, X)
    col = csv + ".crxcol"
    csv_to_columnar.convert(csv, col)
    f = capi.Columnar(col)
    assert np.array_equal(f.rows(ctx), X)          # repr(float) -> float() round-trips: the rows are X itself
    P, Q = f.points(ctx), ctx.points(X)
    a = np.arange(0, 5003, 7, dtype=np.int32)
    b = (a * 31 + 5) % 5003
    for op in range(4):                            # inner product, euclidean, cosine distance, cosine similarity
        assert np.array_equal(capi.pair_op(ctx, P, a, P, b, op), capi.pair_op(ctx, Q, a, Q, b, op))
    cidx = capi.k_means_pp(ctx, P, 9, "euclidean", 3)
    assert np.array_equal(cidx, capi.k_means_pp(ctx, Q, 9, "euclidean", 3))
    la, da = capi.lloyds_assignment(ctx, P, X[cidx], cidx, "euclidean")
    lb, db = capi.lloyds_assignment(ctx, Q, X[cidx], cidx, "euclidean")
    assert np.array_equal(la, lb) and np.array_equal(da, db)
    f.close()


@pytest.mark.gpu
def test_main_cpp_reads_the_sidecar(tmp_path):
    spec = importlib.util.spec_from_file_location("make_main_golden", os.path.join(ROOT, "tests", "golden", "make_main_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    crx_bin = os.path.join(ROOT, "oracle", "_ref", "recommendation_crx")
    assert os.path.exists(crx_bin), "oracle/_ref/recommendation_crx missing: run tools/build_main_dropin.sh where /root/reference exists"
    mg.make_inputs(str(tmp_path))
    src = str(tmp_path / "proj2_input.csv")
    csv_to_columnar.convert(src, src + ".crxcol")
    env = {"CRX_SHIM_PROFILE": "1"}
    out = os.path.join(str(tmp_path), "out_recommendation_crx.txt")
    r = subprocess.run([crx_bin, "-d", "./tweets.tsv", "-o", out], cwd=str(tmp_path), env=dict(os.environ, CRX_FAKE_SEED=mg.SEED, **env),
                       check=True, capture_output=True, timeout=1200)
    assert "VectorReader::read (columnar)" in r.stderr.decode() and "VectorReader::read (text)" not in r.stderr.decode()
    lines = [l for l in open(out).read().splitlines() if not l.startswith("Execution Time:")]
    want = open(os.path.join(ROOT, "tests", "golden", "main_expected.txt")).read().splitlines()
    assert lines == want
