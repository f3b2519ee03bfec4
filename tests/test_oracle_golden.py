"""Pins the oracle (oracle/opal_oracle.c) to the reference's known answers.

Golden vectors were produced by the unmodified reference (tests/golden/make_golden.py); the
survey's hand-recorded known answers are asserted literally as well.
CPU only.
"""
import json
import os

import numpy as np
import pytest

from _util import (GOLDEN_DIR, MODES, OPAL_OVERFLOW_BUCKETS, OPAL_OVERFLOW_SIMPLE, README_DB, README_MATRIX,
                   README_QUERY, SequenceDB, dump_results, free_alignments, glibc_testcpp_data, new_results,
                   search_dump)
from opal_b200 import matrices


def golden(name):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)


def fix_sw_zero(records, mode, search_type):
    """Reference quirk Q2: SW score 0 under SCORE_END leaves garbage end locations
    (src/opal.cpp:221-225, 392-396); the defined answer is (-1,-1)."""
    if mode == "SW" and search_type == 1:
        for r in records:
            if r[1] == 0:
                r[2] = r[3] = -1
    return records


@pytest.mark.parametrize("mode", ["NW", "HW", "OV", "SW"])
@pytest.mark.parametrize("search_type", [0, 1, 2])
def test_readme_example(oracle, mode, search_type):
    g = golden("readme.json")
    db = SequenceDB.from_sequences(README_DB)
    for ovf in (OPAL_OVERFLOW_SIMPLE, OPAL_OVERFLOW_BUCKETS):
        rc, d = search_dump(oracle, README_QUERY, db, 3, 1, README_MATRIX, 4, search_type, MODES[mode], ovf)
        want = g[f"{mode}/{search_type}/{ovf}"]
        assert rc == want["rc"] == 0
        assert d == fix_sw_zero(want["results"], mode, search_type)


def test_readme_known_answers_from_survey(oracle):
    """SURVEY.md section 4, README example under OPAL_SEARCH_ALIGNMENT (score s(q,t) e(q,t) ops)."""
    want = {
        "NW": [(4, 0, 0, 9, 13, "1000301002220222"), (1, 0, 0, 9, 11, "2300010322030"),
               (9, 0, 0, 9, 12, "01300003022202"), (0, 0, 0, 9, 8, "1110130300222")],
        "HW": [(11, 0, 0, 9, 6, "1000110030"), (4, 0, 1, 9, 11, "300010322030"),
               (15, 0, 0, 9, 7, "0130000100"), (7, 0, 1, 9, 5, "1101110300")],
        "OV": [(14, 1, 0, 9, 6, "000110030"), (6, 0, 1, 8, 11, "300010322020"),
               (15, 1, 0, 9, 7, "330000100"), (10, 3, 0, 9, 5, "0130300")],
        "SW": [(14, 1, 0, 9, 6, "000110030"), (10, 1, 2, 5, 6, "00030"),
               (16, 3, 2, 9, 7, "0000100"), (12, 6, 2, 9, 5, "0300")],
    }
    db = SequenceDB.from_sequences(README_DB)
    for mode, recs in want.items():
        rc, d = search_dump(oracle, README_QUERY, db, 3, 1, README_MATRIX, 4, 2, MODES[mode])
        assert rc == 0
        got = [(r[1], r[4], r[5], r[2], r[3], r[7]) for r in d]
        assert got == recs, mode


def test_config1_known_answers_from_survey(oracle):
    """BASELINE configs[0] (opal_aligner -x 1 O74807 vs uniprot_sprot15, BLOSUM50 3/1, SW)."""
    g = golden("config1.json")
    b50 = matrices.blosum50()
    q = np.array(g["query"], dtype=np.uint8)
    db = SequenceDB.from_sequences(g["db"])
    assert len(q) == 110 and len(db) == 15 and db.total_residues == 4491
    rc, d = search_dump(oracle, q, db, 3, 1, b50.flat(), 24, 1, MODES["SW"], OPAL_OVERFLOW_BUCKETS)
    assert rc == 0
    survey = [(155, 106, 161), (178, 109, 229), (169, 107, 314), (119, 108, 145), (152, 109, 437), (68, 102, 58),
              (143, 106, 214), (151, 105, 181), (182, 109, 304), (94, 84, 74), (126, 87, 126), (169, 108, 398),
              (138, 99, 239), (181, 106, 927), (108, 97, 84)]
    assert [(r[1], r[2], r[3]) for r in d] == survey


@pytest.mark.parametrize("key", ["SW/0", "SW/1", "SW/2", "NW/1", "HW/1", "OV/1"])
def test_config1_golden(oracle, key):
    g = golden("config1.json")
    b50 = matrices.blosum50()
    mode, st = key.split("/")
    rc, d = search_dump(oracle, np.array(g["query"], dtype=np.uint8), SequenceDB.from_sequences(g["db"]),
                        3, 1, b50.flat(), 24, int(st), MODES[mode], OPAL_OVERFLOW_BUCKETS)
    assert rc == g[key]["rc"] == 0
    assert d == fix_sw_zero(g[key]["results"], mode, int(st))


@pytest.mark.parametrize("mode,maximum", [("SW", 573), ("NW", 460), ("HW", 567), ("OV", 567)])
def test_reference_selftest_inputs(oracle, mode, maximum):
    """The inputs of the reference's own ./test (src/test.cpp:35-99); `Maximum` per SURVEY.md section 4.
    Score + end are compared for all 200 targets in every mode; start + alignment bit-for-bit for SW.
    (For NW/HW/OV the reference's alignment stage is only validated semantically: SURVEY.md 8c Q9-Q10.)"""
    g = golden("testcpp.json")[mode]
    assert g["maximum"] == maximum
    q, db = glibc_testcpp_data()
    m = matrices.simple(4, 3, -1).flat()
    st = 2 if mode == "SW" else 1
    rc, d = search_dump(oracle, q, db, 11, 1, m, 4, st, MODES[mode], OPAL_OVERFLOW_SIMPLE, digest=True)
    assert rc == 0
    assert max(r[1] for r in d) == maximum
    if mode == "SW":
        assert d == g["results"]
    else:
        assert [r[:4] for r in d] == [r[:4] for r in g["results"]]


@pytest.mark.parametrize("key", ["NW/0", "NW/1", "HW/0", "HW/1", "OV/0", "OV/1", "SW/0", "SW/1", "SW/2"])
def test_protein_golden(oracle, key):
    g = golden("protein.json")
    b62 = matrices.blosum62()
    mode, st = key.split("/")
    rc, d = search_dump(oracle, np.array(g["query"], dtype=np.uint8), SequenceDB.from_sequences(g["db"]),
                        11, 1, b62.flat(), 23, int(st), MODES[mode], OPAL_OVERFLOW_BUCKETS)
    assert rc == g[key]["rc"] == 0
    assert d == fix_sw_zero(g[key]["results"], mode, int(st))


def test_api_semantics(oracle):
    """Reuse rule (src/opal.cpp:1446-1451), CharSW (:1522-1546), invalid mode (:1469-1473)."""
    g = golden("api.json")
    db = SequenceDB.from_sequences(README_DB)
    res = new_results(4)
    args = (README_QUERY, db, 3, 1, README_MATRIX, 4, res)
    oracle.search_database(*args, 0, MODES["SW"], OPAL_OVERFLOW_SIMPLE)
    res["score"][1] = 9999
    oracle.search_database(*args, 0, MODES["SW"], OPAL_OVERFLOW_SIMPLE)
    assert dump_results(res) == g["reuse_score_then_score"]
    oracle.search_database(*args, 1, MODES["SW"], OPAL_OVERFLOW_SIMPLE)
    assert dump_results(res) == g["reuse_then_score_end"]
    oracle.search_database(*args, 2, MODES["SW"], OPAL_OVERFLOW_SIMPLE, entry="opalSearchDatabaseRescore")
    assert dump_results(res) == g["reuse_then_alignment"]
    free_alignments(res)

    c = g["char_sw"]
    rc, res = oracle.search_database_char_sw(np.array(c["query"], dtype=np.uint8), SequenceDB.from_sequences(c["db"]),
                                             3, 1, np.array(c["matrix"], dtype=np.int32), 4)
    assert rc == c["rc"] == 1
    assert dump_results(res, with_alignment=False) == c["results"]

    res = new_results(4)
    rc, res = oracle.search_database(README_QUERY, db, 3, 1, README_MATRIX, 4, res, 0, 7, OPAL_OVERFLOW_SIMPLE)
    assert rc == g["invalid_mode"]["rc"] == 3
    assert dump_results(res) == g["invalid_mode"]["results"]
