"""Generates tests/golden/kmer_golden.json with an implementation that shares no code with the
oracle or the CUDA kernel: plain Python string operations following the semantics of
loaddata/MakeArff.java:357-366 (substring, reverse complement, base-4 value, min).
Run: python tests/golden/make_kmer_golden.py"""
import json
import os
import random

COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
VAL = {"A": 0, "C": 1, "G": 2, "T": 3}


def seq2int(s):
    v = 0
    for ch in s:
        v = v * 4 + VAL[ch]
    return v


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def count(seq, kmin, kmax):
    """Sparse {column: count} for one sequence; None when it holds an N."""
    if "N" in seq:
        return None
    out, base = {}, 0
    for k in range(kmin, kmax + 1):
        for i in range(len(seq) - k + 1):
            w = seq[i:i + k]
            col = base + min(seq2int(w), seq2int(revcomp(w)))
            out[col] = out.get(col, 0) + 1
        base += 4 ** k
    return out


def main():
    rng = random.Random(20260119)
    cases = []
    hand = [("ACGT", 4, 4), ("AAAA", 4, 4), ("TTTT", 4, 4), ("ACGTA", 4, 4), ("ACGTA", 4, 5),
            ("ACG", 4, 5), ("", 4, 5), ("ACGTNACGT", 4, 5), ("GATTACA", 1, 3), ("CCCCCCCCCC", 2, 6)]
    for s, a, b in hand:
        cases.append((s, a, b))
    for L, a, b in [(150, 4, 5), (150, 4, 6), (150, 4, 7), (33, 3, 5), (64, 5, 5), (65, 8, 8),
                    (200, 1, 2), (31, 4, 5), (32, 4, 5), (97, 6, 9)]:
        for _ in range(3):
            cases.append(("".join(rng.choice("ACGT") for _ in range(L)), a, b))
    out = []
    for s, a, b in cases:
        c = count(s, a, b)
        out.append({"seq": s, "kmin": a, "kmax": b, "has_n": c is None,
                    "counts": sorted((c or {}).items())})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kmer_golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()
