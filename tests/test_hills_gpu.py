"""GPU parity, SURVEY.md 8 f-4: sqw_hills_scan / sqw_hills_scan_dev (through the C ABI) vs the CPU
oracle (oracle/hills_oracle.c) and the independent golden fixtures. Positions, lengths, list order
and fp64 scores are bit-exact: the kernel accumulates every window score in the reference's order
(motifs/Discrim.java:394-405)."""
import json
import os

import numpy as np
import pytest

from sequnwinder_b200 import synth
from sequnwinder_b200.hills import HillsScanner
from sequnwinder_b200.kmer import num_kmers, pack_sequences

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hills_golden.json")


def same(res, want):
    st, ln, sc, cnt, hn = want
    assert np.array_equal(res.has_n, hn)
    assert np.array_equal(res.counts, cnt)
    cap = min(res.start.shape[1], st.shape[1])
    mask = np.arange(cap)[None, :] < cnt[:, None]
    assert np.array_equal(res.start[:, :cap][mask], st[:, :cap][mask])
    assert np.array_equal(res.length[:, :cap][mask], ln[:, :cap][mask])
    assert np.array_equal(res.score[:, :cap][mask].view(np.int64), sc[:, :cap][mask].view(np.int64))


def test_golden_fixtures():
    for case in json.load(open(GOLDEN)):
        w = np.array(case["wi"], dtype=np.float64) / case["wdiv"]
        hs = HillsScanner(case["kmin"], case["kmax"], case["min_m"], case["max_m"], case["thresh"])
        res = hs.scan(*pack_sequences([case["seq"]]), w, mode=case["mode"], cap=64)
        want = [(s, l, float.fromhex(h)) for s, l, h in case["hills"]]
        got = [(int(res.start[0, j]), int(res.length[0, j]), float(res.score[0, j]))
               for j in range(int(res.counts[0]))]
        assert bool(res.has_n[0]) == case["has_n"]
        assert got == want


@pytest.mark.parametrize("cfg,n,quant", [("cfg1", None, False), ("cfg1", None, True),
                                         ("cfg3", 1500, False), ("cfg4", 3000, True)])
@pytest.mark.parametrize("mode", [0, 1])
def test_configs_bit_exact(oracle, cfg, n, quant, mode):
    c = synth.CONFIGS[cfg]
    d = synth.generate_config(cfg, n, n_frac=0.005)
    rng = np.random.default_rng(11)
    numk = num_kmers(c["kmin"], c["kmax"])
    # quantised weights make exactly tied scores common (the two modes differ only on ties)
    w = rng.integers(-3, 5, numk) / 8.0 if quant else rng.normal(0.0, 0.08, numk)
    hs = HillsScanner(c["kmin"], c["kmax"], 6, 14, 0.1)
    cap = 96
    want = oracle.hills_scan(d.bases, d.offsets, c["kmin"], c["kmax"], w, 6, 14, 0.1, mode, cap)
    res = hs.scan(d.bases, d.offsets, w, mode=mode, cap=cap)
    assert want[3].sum() > 0 and want[4].sum() > 0
    same(res, want)


def test_ragged_empty_lowercase_and_quirk(oracle):
    from test_hills_oracle import quirk_case
    rng = np.random.default_rng(3)
    seqs = ["", "A", "ACGTA", "ACGTAC", "acgtacgtacgtaaccggtt", "ACGTNACGTACGTACGT", "n"]
    seqs += ["".join(rng.choice(list("ACGT"), size=int(L))) for L in rng.integers(0, 500, 300)]
    up = [s.upper() for s in seqs]
    for kmin, kmax, mn, mx, th in [(4, 5, 6, 14, 0.1), (1, 3, 2, 7, 0.0), (4, 7, 6, 10, 0.3), (3, 3, 3, 40, -1.0)]:
        w = rng.normal(0.0, 0.1, num_kmers(kmin, kmax))
        for mode in (0, 1):
            want = oracle.hills_scan(*oracle.pack_sequences(up), kmin, kmax, w, mn, mx, th, mode, 512)
            res = HillsScanner(kmin, kmax, mn, mx, th).scan(*pack_sequences(seqs), w, mode=mode, cap=512)
            same(res, want)
    seq, w = quirk_case()
    res = HillsScanner(2, 2, 2, 8, 0.5).scan(*pack_sequences([seq]), w, mode=0, cap=16)
    assert res.counts[0] == 1 and (res.start[0, 0], res.length[0, 0], res.score[0, 0]) == (6, 2, 3.0)


def test_errors():
    from sequnwinder_b200 import _lib
    w = np.full(1280, 0.125)
    hs = HillsScanner(4, 5, 6, 6, 0.1)
    with pytest.raises(ValueError):
        hs.scan(*pack_sequences(["ACGTACGTAC", "ACGTRYGTAC"]), w)
    # N wins over another bad character (the reference skips the sequence first)
    res = hs.scan(*pack_sequences(["ACGTRYGTN"]), w)
    assert res.has_n[0] == 1 and res.counts[0] == 0
    with pytest.raises(_lib.SqwError) as e:     # every window ties: more hills than cap
        hs.scan(*pack_sequences(["ACGT" * 20]), w, cap=2)
    assert e.value.code == _lib.SQW_ERR_UNSUPPORTED
    with pytest.raises(_lib.SqwError):
        HillsScanner(4, 5, 14, 6, 0.1).scan(*pack_sequences(["ACGTACGT"]), w)
    res = hs.scan(np.zeros(0, np.uint8), np.zeros(1, np.int64), w)
    assert res.counts.shape == (0,)


def test_reference_entry_points(oracle):
    rng = np.random.default_rng(9)
    seqs = ["".join(rng.choice(list("ACGT"), size=150)) for _ in range(50)]
    w = rng.normal(0.0, 0.1, 1280)
    hs = HillsScanner(4, 5, 6, 14, 0.1)
    seqlets = hs.findSeqlets(seqs, w, sort=True)
    st, ln, sc, cnt, _ = oracle.hills_scan(*oracle.pack_sequences(seqs), 4, 5, w, 6, 14, 0.1, 0, 64)
    flat = [(seqs[r][st[r, j]:st[r, j] + ln[r, j]], float(sc[r, j])) for r in range(50) for j in range(cnt[r])]
    assert seqlets == sorted(flat, key=lambda pr: -pr[1])      # stable, descending (Discrim.java:241-249)
    assert all(6 <= len(s) <= 14 and v > 0.1 for s, v in seqlets)
    hills = hs.findHills([s.lower() for s in seqs], [1000 * i for i in range(50)], w)
    st, ln, sc, cnt, _ = oracle.hills_scan(*oracle.pack_sequences(seqs), 4, 5, w, 6, 14, 0.1, 1, 64)
    flat = [(r, 1000 * r + int(st[r, j]), 1000 * r + int(st[r, j]) + int(ln[r, j]) - 1, float(sc[r, j]))
            for r in range(50) for j in range(cnt[r])]
    assert hills == flat


def test_device_resident_full_size_properties():
    """cfg-2 at full size through the _dev entry point: size-independent properties."""
    import torch
    c = synth.CONFIGS["cfg2"]
    d = synth.generate_config("cfg2")
    rng = np.random.default_rng(2)
    w = rng.normal(0.0, 0.08, num_kmers(c["kmin"], c["kmax"]))
    hs = HillsScanner(c["kmin"], c["kmax"], 6, 14, 0.1)
    dev = torch.device("cuda:0")
    bases = torch.from_numpy(d.bases).to(dev)
    offsets = torch.from_numpy(d.offsets).to(dev)
    wd = torch.from_numpy(w).to(dev)
    L = int(np.diff(d.offsets).max())
    for mode in (0, 1):
        st, ln, sc, cnt, hn, status = hs.scan_device(bases, offsets, L, wd, mode=mode, cap=64)
        torch.cuda.synchronize()
        assert int(status.item()) == 0 and int(hn.sum().item()) == 0
        st, ln, sc, cnt = st.cpu().numpy(), ln.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()
        mask = np.arange(64)[None, :] < cnt[:, None]
        assert cnt.sum() > 20000
        assert (sc[mask] > 0.1).all() and (ln[mask] >= 6).all() and (ln[mask] <= 14).all()
        assert (st[mask] >= 0).all() and ((st + ln)[mask] <= L).all()
        # every reported score is the window's score recomputed independently (numpy, same order)
        from sequnwinder_b200 import KmerCounter
        r_idx, j_idx = np.nonzero(mask)
        pick = rng.choice(len(r_idx), 300, replace=False)
        for t in pick:
            r, j = r_idx[t], j_idx[t]
            s0, l0 = int(st[r, j]), int(ln[r, j])
            sub = d.bases[d.offsets[r] + s0: d.offsets[r] + s0 + l0].tobytes().decode()
            exp = 0.0
            base = 0
            for k in range(c["kmin"], c["kmax"] + 1):
                for q in range(l0 - k + 1):
                    km = sub[q:q + k]
                    val = lambda t_: int("".join(str("ACGT".index(ch)) for ch in t_), 4)
                    rc = "".join({"A": "T", "C": "G", "G": "C", "T": "A"}[ch] for ch in reversed(km))
                    exp = exp + w[base + min(val(km), val(rc))]
                base += 4 ** k
            assert exp == sc[r, j]
        # hills of one sequence never overlap unless their scores are exactly equal (mode 0);
        # in mode 1 they never overlap at all
        for r in rng.choice(len(cnt), 400, replace=False):
            k_ = int(cnt[r])
            for a in range(k_):
                for b in range(a + 1, k_):
                    ov = st[r, a] <= st[r, b] + ln[r, b] - 1 and st[r, a] + ln[r, a] - 1 >= st[r, b]
                    if ov:
                        assert mode == 0 and sc[r, a] == sc[r, b]
