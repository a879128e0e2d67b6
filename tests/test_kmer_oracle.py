"""CPU: pins the k-mer oracle (oracle/kmer_oracle.c, restating loaddata/MakeArff.java:323-377)
against hand-derived known answers and the independent pure-Python fixtures in
tests/golden/kmer_golden.json."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kmer_golden.json")


def dense(case, numk):
    row = np.zeros(numk, dtype=np.int32)
    for col, cnt in case["counts"]:
        row[col] = cnt
    return row


def test_known_answers(oracle):
    O = oracle
    one = lambda s, a, b: O.kmer_count(*O.pack_sequences([s]), a, b)
    c, n = one("ACGT", 4, 4)
    assert c.shape == (1, 256) and c[0, 27] == 1 and c.sum() == 1 and n[0] == 0
    c, _ = one("AAAA", 4, 4)
    assert c[0, 0] == 1 and c.sum() == 1
    c, _ = one("TTTT", 4, 4)          # revcomp AAAA -> column 0
    assert c[0, 0] == 1 and c.sum() == 1
    c, _ = one("ACGTA", 4, 4)         # ACGT -> 27, CGTA(108) vs TACG(198) -> 108
    assert c[0, 27] == 1 and c[0, 108] == 1 and c.sum() == 2
    c, _ = one("ACGTA", 4, 5)         # the k=5 block starts at column 256 (MakeArff.java:367)
    assert c.shape == (1, 1280) and c[0, :256].sum() == 2 and c[0, 256:].sum() == 1
    fwd = int("".join(str("ACGT".index(ch)) for ch in "ACGTA"), 4)
    rc = int("".join(str("ACGT".index(ch)) for ch in "TACGT"), 4)
    assert c[0, 256 + min(fwd, rc)] == 1


def test_n_rows_flagged_and_zero(oracle):
    c, n = oracle.kmer_count(*oracle.pack_sequences(["ACGTACGT", "ACGTNACGT", "NNNN"]), 4, 5)
    assert n.tolist() == [0, 1, 1]
    assert c[1].sum() == 0 and c[2].sum() == 0 and c[0].sum() == 5 + 4


def test_bad_character_raises(oracle):
    with pytest.raises(ValueError):
        oracle.kmer_count(*oracle.pack_sequences(["ACGTRCGT"]), 4, 5)
    # an N anywhere wins over the bad character: the row is skipped before seq2int runs
    c, n = oracle.kmer_count(*oracle.pack_sequences(["ACGTRCGTN"]), 4, 5)
    assert n[0] == 1


def test_row_invariants(oracle):
    rng = np.random.default_rng(7)
    seqs = ["".join(rng.choice(list("ACGT"), size=150)) for _ in range(50)]
    c, n = oracle.kmer_count(*oracle.pack_sequences(seqs), 4, 5)
    assert (c.sum(1) == 293).all()                       # sum_k (L-k+1)
    # columns whose index exceeds the index of their reverse complement stay zero
    comp = lambda v, k: int("".join(str(3 - int(d)) for d in reversed(np.base_repr(v, 4).zfill(k))), 4)
    base, dead = 0, 0
    for k in (4, 5):
        for v in range(4 ** k):
            if v > comp(v, k):
                assert c[:, base + v].sum() == 0
                dead += 1
        base += 4 ** k
    assert dead == 120 + 512


def test_against_pure_python_golden(oracle):
    cases = json.load(open(GOLDEN))
    assert len(cases) >= 40
    for case in cases:
        numk = oracle.num_kmers(case["kmin"], case["kmax"])
        c, n = oracle.kmer_count(*oracle.pack_sequences([case["seq"]]), case["kmin"], case["kmax"])
        assert bool(n[0]) == case["has_n"]
        assert np.array_equal(c[0], dense(case, numk)), case["seq"][:20]


def test_batched_equals_single(oracle):
    cases = [c for c in json.load(open(GOLDEN)) if (c["kmin"], c["kmax"]) == (4, 5)]
    seqs = [c["seq"] for c in cases]
    got, n = oracle.kmer_count(*oracle.pack_sequences(seqs), 4, 5)
    for i, case in enumerate(cases):
        assert np.array_equal(got[i], dense(case, 1280))
