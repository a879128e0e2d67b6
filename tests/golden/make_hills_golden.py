"""Generates tests/golden/hills_golden.json with an implementation that shares no code with the
oracle or the CUDA kernel: plain Python strings and list operations following the statements of
motifs/Discrim.java:382-444 (findSeqlets) and :328-380 (findHills) one by one — substring,
reverse complement, base-4 value, min, score accumulated from 0.0 with `score = score + w`,
an explicit iterator over the mountains list with remove(). Weights are multiples of 1/8 in half
of the cases so that exactly tied scores occur (the two functions differ only on ties).
Run: python tests/golden/make_hills_golden.py"""
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


def base_ind(k, kmin):                      # SeqUnwinderConfig.getKmerBaseInd
    return sum(4 ** kk for kk in range(kmin, k))


def scan(seq, kmin, kmax, weights, min_m, max_m, thresh, mode):
    """[(start, length, score)] in list order, or None for a sequence with N."""
    if "N" in seq:
        return None
    mountains = []                          # [start, end, score]
    for l in range(min_m, max_m + 1):
        for i in range(0, len(seq) - l + 1):
            motif = seq[i:i + l]
            score = 0.0
            for k in range(kmin, kmax + 1):
                for j in range(0, len(motif) - k + 1):
                    currk = motif[j:j + k]
                    kmer = min(seq2int(currk), seq2int(revcomp(currk)))
                    score = score + weights[base_ind(k, kmin) + kmer]
            hs, he = i, i + l - 1
            add = True
            idx = 0                          # Java iterator: next() returns mountains[idx]
            while idx < len(mountains) and add:
                cs, ce, cscore = mountains[idx]
                overlaps = cs <= he and ce >= hs
                lower = (cscore <= score) if mode == 1 else (cscore < score)
                if overlaps and lower:
                    del mountains[idx]       # it.remove()
                    add = True
                elif overlaps and cscore > score:
                    add = False
                else:
                    idx += 1
            if add and score > thresh:
                mountains.append([hs, he, score])
    return [(s, e - s + 1, sc) for s, e, sc in mountains]


def main():
    rng = random.Random(20260119 + 4)
    cases = []
    specs = [(60, 4, 5, 6, 14, 0.1), (150, 4, 5, 6, 14, 0.1), (40, 3, 4, 5, 9, 0.0), (25, 4, 5, 6, 14, 0.1),
             (5, 4, 5, 6, 14, 0.1), (90, 3, 5, 6, 10, 0.5), (33, 2, 3, 3, 6, -0.25), (64, 4, 4, 4, 8, 0.1)]
    for L, kmin, kmax, mn, mx, th in specs:
        numk = sum(4 ** k for k in range(kmin, kmax + 1))
        for quant in (False, True):
            for mode in (0, 1):
                # weights are stored as integers wi with w = wi / wdiv (exact in both places)
                wdiv = 8 if quant else 10000
                wi = ([rng.randint(-3, 4) for _ in range(numk)] if quant
                      else [int(round(rng.gauss(0.0, 0.15) * 10000)) for _ in range(numk)])
                w = [v / wdiv for v in wi]
                seq = "".join(rng.choice("ACGT") for _ in range(L))
                if L == 25 and quant:
                    seq = seq[:10] + "N" + seq[11:]
                if L == 40:
                    seq = (seq[:8] * 5)[:L]          # repeats: identical windows, exact ties
                res = scan(seq, kmin, kmax, w, mn, mx, th, mode)
                cases.append({"seq": seq, "kmin": kmin, "kmax": kmax, "min_m": mn, "max_m": mx,
                              "thresh": th, "mode": mode, "wi": wi, "wdiv": wdiv, "has_n": res is None,
                              "hills": [[s, l, sc.hex()] for s, l, sc in (res or [])]})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hills_golden.json")
    with open(path, "w") as fh:
        json.dump(cases, fh)
    print("wrote", path, len(cases), "cases;", sum(len(c["hills"]) for c in cases), "hills")


if __name__ == "__main__":
    main()
