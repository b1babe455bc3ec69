"""bamqc_b200/summary.py (the step after the statistics pass) against golden vectors written by the
reference's own bamqc_summary.py (tests/golden/make_summary_golden.py).  Floats within 1e-9 relative (the north-star
tolerance for derived floating summaries); strings, integers and flag sets exact."""
import json
import math
import os

import pytest

from bamqc_b200 import summary as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = json.load(open(os.path.join(GOLDEN, "summary", "summary_golden.json")))
FILES = sorted(k for k in REF if not k.startswith("__"))
REL_TOL = 1e-9


def _same(a, b):
    if isinstance(a, float) or isinstance(b, float):
        if isinstance(a, str) or isinstance(b, str):
            return a == b
        return math.isclose(float(a), float(b), rel_tol=REL_TOL, abs_tol=1e-300)
    return a == b


@pytest.mark.parametrize("name", FILES)
def test_fields_flags_and_layouts(name):
    ref = REF[name]["lanes"]
    lanes = S.read_bamqc(os.path.join(GOLDEN, name))
    assert len(lanes) == len(ref)
    for lane, r in zip(lanes, ref):
        if "summarize_error" in r:   # the reference fails on this input (e.g. no reads at all): so does the equivalent
            with pytest.raises(Exception) as ei:
                S.summarize(lane)
            assert type(ei.value).__name__ == r["summarize_error"]
            continue
        s = S.summarize(lane)
        for k, v in r["fields"].items():
            assert k in s, k
            assert _same(s[k], v), (k, s[k], v)
        for st in (3, 2, 1):
            s["flags%d" % st] = S.flags(s, st)
            assert sorted(s["flags%d" % st]) == r["flags%d" % st], st
        # dense line: same columns; the flag columns are sets (Python's set order), the numbers are compared as numbers
        mine, theirs = S.dense_line(s) + "\n", r["dense"]
        cm, ct = mine.split("\t"), theirs.split("\t")
        assert len(cm) == len(ct)
        for i, (x, y) in enumerate(zip(cm, ct)):
            if i in (2, 3, 4):
                assert set(x.strip()) == set(y.strip())
                continue
            assert x[len(x.rstrip(" \n")):] == y[len(y.rstrip(" \n")):]      # the blanks that separate the groups
            xs, ys = x.strip(), y.strip()
            try:
                assert math.isclose(float(xs), float(ys), rel_tol=REL_TOL, abs_tol=1e-300), (i, xs, ys)
            except ValueError:
                assert xs == ys, (i, xs, ys)
        # long layout: identical text; where the reference stops with a ValueError on an 'NA' value, identical up to there
        text = S.long_text(s)
        if r["long_error"] is None:
            assert text == r["long"]
        else:
            assert text.startswith(r["long"]) and len(text) > len(r["long"])


def test_header_line():
    assert S.dense_header() + "\n" == REF["__header__"]


def test_kmer_error_rate_known_points():
    assert S.kmer_error_rate(100, 10, 0, 32) == (0, "NA")
    x, e = S.kmer_error_rate(1000000, 400000, 30000000, 32)
    assert 0 < e < 1 and x > 0


def test_cli_entry(tmp_path, capsys):
    assert S.main(["-t", os.path.join(GOLDEN, "standard.bamqc")]) == 0
    out = capsys.readouterr().out.split("\n")
    assert out[0].startswith("SAMPLE_ID\tLANE") and out[1].startswith("S1\tL1\t")
