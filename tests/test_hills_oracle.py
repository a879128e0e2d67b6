"""CPU: pins the hills-scan oracle (oracle/hills_oracle.c, restating motifs/Discrim.java:328-444)
against hand-derived known answers and the independent pure-Python fixtures in
tests/golden/hills_golden.json (generator: tests/golden/make_hills_golden.py)."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hills_golden.json")


def load_cases():
    with open(GOLDEN) as fh:
        return json.load(fh)


def expected(case):
    return [(s, l, float.fromhex(h)) for s, l, h in case["hills"]]


def run_case(scan, pack, case, cap=64):
    w = np.array(case["wi"], dtype=np.float64) / case["wdiv"]
    st, ln, sc, cnt, hn = scan(*pack([case["seq"]]), case["kmin"], case["kmax"], w, case["min_m"],
                               case["max_m"], case["thresh"], case["mode"], cap)
    return [(int(st[0, j]), int(ln[0, j]), float(sc[0, j])) for j in range(int(cnt[0]))], int(hn[0])


def test_golden_fixtures(oracle):
    cases = load_cases()
    assert len(cases) >= 30
    for case in cases:
        got, hn = run_case(oracle.hills_scan, oracle.pack_sequences, case)
        assert hn == int(case["has_n"])
        assert got == expected(case)       # positions, lengths and scores bit-exact, list order


def test_known_answers(oracle):
    O = oracle
    numk = 256                              # k = 4 only
    w = np.zeros(numk)
    w[27] = 1.0                             # ACGT (its own reverse complement)
    seq = "TTTTACGTTTTT"
    # windows of length 4..5: only those containing ACGT score 1.0 > 0.5. List order: l=4 first
    # (start 4), then l=5 candidates (starts 3 and 4) tie with it: findSeqlets keeps all three,
    # findHills lets each tie replace what it overlaps.
    st, ln, sc, cnt, hn = O.hills_scan(*O.pack_sequences([seq]), 4, 4, w, 4, 5, 0.5, 0, 8)
    assert hn[0] == 0 and cnt[0] == 3
    assert [(st[0, j], ln[0, j]) for j in range(3)] == [(4, 4), (3, 5), (4, 5)]
    assert sc[0, :3].tolist() == [1.0, 1.0, 1.0]
    st, ln, sc, cnt, hn = O.hills_scan(*O.pack_sequences([seq]), 4, 4, w, 4, 5, 0.5, 1, 8)
    assert cnt[0] == 1 and (st[0, 0], ln[0, 0], sc[0, 0]) == (4, 5, 1.0)
    # threshold is strict (Discrim.java:425): score == thresh is not a hill
    st, ln, sc, cnt, hn = O.hills_scan(*O.pack_sequences([seq]), 4, 4, w, 4, 5, 1.0, 0, 8)
    assert cnt[0] == 0
    # reverse complement collapses: AAAA and TTTT share column 0
    w2 = np.zeros(numk)
    w2[0] = 0.25
    st, ln, sc, cnt, hn = O.hills_scan(*O.pack_sequences(["AAAAAA"]), 4, 4, w2, 6, 6, 0.1, 0, 8)
    assert cnt[0] == 1 and sc[0, 0] == 0.75 and (st[0, 0], ln[0, 0]) == (0, 6)
    st, ln, sc, cnt, hn = O.hills_scan(*O.pack_sequences(["TTTTTT"]), 4, 4, w2, 6, 6, 0.1, 0, 8)
    assert cnt[0] == 1 and sc[0, 0] == 0.75


def canon2(s):
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    val = lambda t: int("".join(str("ACGT".index(c)) for c in t), 4)
    return min(val(s), val("".join(comp[c] for c in reversed(s))))


def quirk_case():
    """k = 2, sequence ACTTTTGG: hills AC@0 (1.0) and GG@6 (3.0) come from l = 2; every longer
    window holding only one of them scores less than that hill and is vetoed by it; the full
    window (l = 8) scores 1 - 0.25 - 3*0.5 - 0.25 + 3 = 2.0: the walk removes AC@0 (lower), then
    meets GG@6 (higher), which vetoes the candidate. Only GG@6 is left."""
    w = np.zeros(16)
    w[canon2("AC")] = 1.0
    w[canon2("GG")] = 3.0
    w[canon2("CT")] = -0.25
    w[canon2("TG")] = -0.25
    w[canon2("TT")] = -0.5
    return "ACTTTTGG", w


def test_removed_hills_stay_removed_when_the_candidate_is_vetoed(oracle):
    """Discrim.java:411-424: the walk removes lower overlapping hills until a higher one ends it;
    the candidate is then not added, but the removals stand."""
    O = oracle
    seq, w = quirk_case()
    for mode in (0, 1):
        st, ln, sc, cnt, hn = O.hills_scan(*O.pack_sequences([seq]), 2, 2, w, 2, 8, 0.5, mode, 16)
        assert cnt[0] == 1 and (st[0, 0], ln[0, 0], sc[0, 0]) == (6, 2, 3.0)
    # without the l = 8 window both hills survive
    st, ln, sc, cnt, hn = O.hills_scan(*O.pack_sequences([seq]), 2, 2, w, 2, 7, 0.5, 0, 16)
    assert cnt[0] == 2 and [(st[0, j], ln[0, j], sc[0, j]) for j in range(2)] == [(0, 2, 1.0), (6, 2, 3.0)]


def test_n_sequences_short_sequences_and_overflow(oracle):
    O = oracle
    rng = np.random.default_rng(5)
    w = rng.normal(0, 0.2, 1280)
    seqs = ["ACGTACGTACGTAAACCCGGGTTT", "ACGTNACGTACGT", "ACG", "", "GGGGGGGGGGGGGGGGGGGG"]
    st, ln, sc, cnt, hn = O.hills_scan(*O.pack_sequences(seqs), 4, 5, w, 6, 14, 0.1, 0, 32)
    assert hn.tolist() == [0, 1, 0, 0, 0]
    assert cnt[1] == 0 and cnt[2] == 0 and cnt[3] == 0
    with pytest.raises(ValueError):
        O.hills_scan(*O.pack_sequences(["ACGTXACGTAC"]), 4, 5, w, 6, 14, 0.1, 0, 32)
    wq = np.full(1280, 0.125)               # every window ties within its length class
    with pytest.raises(OverflowError):
        O.hills_scan(*O.pack_sequences(["ACGT" * 20]), 4, 5, wq, 6, 6, 0.1, 0, 2)
