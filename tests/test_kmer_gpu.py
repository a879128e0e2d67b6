"""GPU parity, path 1: sqw_kmer_count / sqw_kmer_count_dev (through the C ABI) vs the CPU oracle
and the golden fixtures. Bit-exact (integer work)."""
import json
import os

import numpy as np
import pytest

from sequnwinder_b200 import KmerCounter, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kmer_golden.json")


def test_golden_fixtures():
    cases = json.load(open(GOLDEN))
    by_k = {}
    for c in cases:
        by_k.setdefault((c["kmin"], c["kmax"]), []).append(c)
    for (kmin, kmax), group in by_k.items():
        kc = KmerCounter(kmin, kmax)
        got, has_n = kc.count_sequences([c["seq"] for c in group])
        for i, c in enumerate(group):
            want = np.zeros(kc.numK, dtype=np.int32)
            for col, cnt in c["counts"]:
                want[col] = cnt
            assert bool(has_n[i]) == c["has_n"]
            assert np.array_equal(got[i], want), (kmin, kmax, c["seq"][:16])


@pytest.mark.parametrize("cfg,n", [("cfg1", None), ("cfg2", None), ("cfg3", 3000), ("cfg4", 20000),
                                    ("cfg3", 7000)])   # 7000 x k4..7 rows: two chunks of the host path
def test_configs_bit_exact(oracle, cfg, n):
    c = synth.CONFIGS[cfg]
    d = synth.generate_config(cfg, n, n_frac=0.005)
    want, want_n = oracle.kmer_count(d.bases, d.offsets, c["kmin"], c["kmax"])
    got, got_n = KmerCounter(c["kmin"], c["kmax"]).count(d.bases, d.offsets)
    assert np.array_equal(want_n, got_n) and want_n.sum() > 0
    assert np.array_equal(want, got)


def test_ragged_empty_and_lowercase(oracle):
    rng = np.random.default_rng(3)
    seqs = ["", "A", "ACG", "ACGT", "acgtacgtac", "ACGTNACGT", "n"]
    seqs += ["".join(rng.choice(list("ACGT"), size=int(L))) for L in rng.integers(0, 400, 200)]
    up = [s.upper() for s in seqs]
    for kmin, kmax in [(1, 3), (4, 5), (4, 7), (8, 8), (2, 9)]:   # (8,8),(2,9): global-atomic path
        want, want_n = oracle.kmer_count(*oracle.pack_sequences(up), kmin, kmax)
        got, got_n = KmerCounter(kmin, kmax).count_sequences(seqs)
        assert np.array_equal(want_n, got_n)
        assert np.array_equal(want, got), (kmin, kmax)


def test_zero_rows():
    got, has_n = KmerCounter(4, 5).count(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert got.shape == (0, 1280) and has_n.shape == (0,)


def test_bad_base_is_an_error():
    with pytest.raises(ValueError):
        KmerCounter(4, 5).count_sequences(["ACGTACGT", "ACGTRYGT"])
    # N wins: the reference skips the row before seq2int can throw (MakeArff.java:354-355)
    got, has_n = KmerCounter(4, 5).count_sequences(["ACGTRYGTN"])
    assert has_n[0] == 1 and got.sum() == 0


def test_device_resident_full_size_properties():
    """cfg-2 at full size through the _dev entry point: size-independent properties."""
    import torch
    c = synth.CONFIGS["cfg2"]
    d = synth.generate_config("cfg2")
    kc = KmerCounter(c["kmin"], c["kmax"])
    b = torch.from_numpy(d.bases).cuda()
    o = torch.from_numpy(d.offsets).cuda()
    out, has_n, status = kc.count_device(b, o, 150)
    torch.cuda.synchronize()
    assert int(status.item()) == 0 and int(has_n.sum().item()) == 0
    assert bool((out.sum(1) == 293).all())               # sum_k (L-k+1)
    # reverse-complementing every sequence leaves the collapsed counts unchanged
    comp = torch.zeros(256, dtype=torch.uint8)
    for a, t in zip(b"ACGT", b"TGCA"):
        comp[a] = t
    rc = comp.cuda()[b.long()].view(d.n_sites, 150).flip(1).contiguous().view(-1)
    out2, _, _ = kc.count_device(rc, o, 150)
    assert torch.equal(out, out2)
    # host entry point agrees with the device-resident one
    host, _ = kc.count(d.bases, d.offsets)
    assert np.array_equal(host, out.cpu().numpy())


def test_row_longer_than_max_len_is_reported_not_overrun():
    """_dev entry point with a max_len smaller than the longest row: the row is truncated and the
    status word says so (bit value 4); the neighbouring rows are unaffected."""
    import torch
    seqs = ["ACGT" * 10, "ACGTTGCA" * 40, "TTGACA" * 8]
    kc = KmerCounter(4, 5)
    want, _ = kc.count_sequences(seqs)
    bases = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()
    offs = np.cumsum([0] + [len(s) for s in seqs]).astype(np.int64)
    out, has_n, status = kc.count_device(torch.from_numpy(bases).cuda(), torch.from_numpy(offs).cuda(), 64)
    torch.cuda.synchronize()
    assert int(status.item()) & 4
    got = out.cpu().numpy()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2])
    assert got[1].sum() == (64 - 3) + (64 - 4)                    # the truncated row: 64 bases
