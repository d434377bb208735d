"""the reference's unmodified main.cpp, compiled against the drop-in headers
(include/crx/lib) and linked to libcrx.so, must write the same output file as main.cpp compiled against the
reference's own headers.  Both binaries are built by tools/build_main_dropin.sh (from /root/reference, in the build
container) into oracle/_ref/, which travels to the GPU box; the expected output is committed as
tests/golden/main_expected.txt (generator: tests/golden/make_main_golden.py)."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_main_golden", os.path.join(ROOT, "tests", "golden", "make_main_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "recommendation_ref")
CRX_BIN = os.path.join(ROOT, "oracle", "_ref", "recommendation_crx")


def expected():
    with open(os.path.join(ROOT, "tests", "golden", "main_expected.txt")) as f:
        return f.read().splitlines()


def sections(lines):
    """[(title, [lines])] of the four stages (main.cpp:151,192,241,312)."""
    out = []
    for l in lines:
        if l in ("Cosine LSH", "Clustering Recommendation"):
            out.append((l, []))
        else:
            out[-1][1].append(l)
    return out


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/recommendation_ref not built (needs /root/reference)")
def test_reference_main_reproduces_golden(tmp_path):
    mg.make_inputs(str(tmp_path))
    assert mg.run_main(REF_BIN, str(tmp_path)) == expected()


def expected_validation():
    with open(os.path.join(ROOT, "tests", "golden", "main_validate_expected.txt")) as f:
        return f.read().splitlines()


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/recommendation_ref not built (needs /root/reference)")
def test_reference_validation_reproduces_golden(tmp_path):
    mg.make_inputs(str(tmp_path))
    lines, val = mg.run_main(REF_BIN, str(tmp_path), validate=True)
    assert lines == expected() and val == expected_validation()


@pytest.mark.gpu
def test_ten_fold_validation_over_dropin_headers_matches_reference(tmp_path):
    # f-3: main.cpp -validate = split_to_10 / merge_except_for / hide_one_score (drop-in mirrors) around ten rebuilds of
    # the LSH tables + get_P_closest + get_predicted_user_sim on the engine; every running error sum and the MAE printed
    # by main.cpp:393-437 must equal the reference build's
    assert os.path.exists(CRX_BIN), "oracle/_ref/recommendation_crx missing: run tools/build_main_dropin.sh where /root/reference exists"
    mg.make_inputs(str(tmp_path))
    lines, val = mg.run_main(CRX_BIN, str(tmp_path), validate=True)
    want = expected_validation()
    assert len(val) == len(want) and len(want) > 100
    bad = [(i, a, b) for i, (a, b) in enumerate(zip(val, want)) if a != b]
    assert not bad, "%d of %d validation lines differ, first: %r" % (len(bad), len(want), bad[:3])
    assert [l for l in lines] == expected()


@pytest.mark.gpu
def test_main_cpp_over_dropin_headers_matches_reference(tmp_path):
    assert os.path.exists(CRX_BIN), "oracle/_ref/recommendation_crx missing: run tools/build_main_dropin.sh where /root/reference exists"
    mg.make_inputs(str(tmp_path))
    got = sections(mg.run_main(CRX_BIN, str(tmp_path)))
    want = sections(expected())
    assert [t for t, _ in got] == [t for t, _ in want]
    for (title, g), (_, w) in zip(got, want):
        assert len(g) == len(w), title
        bad = [(a, b) for a, b in zip(g, w) if a != b]
        assert not bad, "%s: %d of %d users differ, first: %r" % (title, len(bad), len(w), bad[:3])
