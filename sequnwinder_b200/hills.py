"""Host-side mirror of the hills scan of ``motifs/Discrim.java`` (inner class ``KmerModelScanner``,
:328-444, relative to /root/reference/src/org/seqcode/projects/sequnwinder/): the sliding-window
scan of one label's k-mer weights over the training sequences that yields the "hills" (regions
mode, ``findHills``) or "seqlets" (sequence mode, ``findSeqlets``) that are clustered and handed
to MEME. SURVEY.md 8 f-4.

The scan itself is the CUDA kernel behind ``sqw_hills_scan`` / ``sqw_hills_scan_dev``
(csrc/hills_scan.cu); there is no CPU path here. This module only flattens the sequences,
turns the [n, cap] result into the reference's lists and applies ``Collections.sort``'s stable
descending order (Discrim.java:231-249).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from .kmer import num_kmers, pack_sequences


@dataclass
class HillsResult:
    """Hills of sequence r, in the reference's list order: start/length/score[r, :counts[r]]."""
    start: np.ndarray    # [n, cap] int32, 0-based offset in the sequence
    length: np.ndarray   # [n, cap] int32
    score: np.ndarray    # [n, cap] float64
    counts: np.ndarray   # [n] int32
    has_n: np.ndarray    # [n] uint8: sequence skipped (contains N)

    def flat(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """(sequence index, start, length, score) of every hill in the order the reference's
        ``ret`` list has: sequences in input order, hills in list order."""
        n, cap = self.start.shape
        mask = np.arange(cap)[None, :] < self.counts[:, None]
        seq = np.broadcast_to(np.arange(n)[:, None], (n, cap))[mask]
        return seq, self.start[mask], self.length[mask], self.score[mask]


class HillsScanner:
    """Scan parameters of SeqUnwinderConfig: k range, minscanlen / maxscanlen (:429-430),
    hillsthresh (:431)."""

    def __init__(self, kmin: int = 4, kmax: int = 5, min_m: int = 6, max_m: int = 14,
                 hills_thresh: float = 0.1):
        self.kmin, self.kmax = int(kmin), int(kmax)
        self.min_m, self.max_m = int(min_m), int(max_m)
        self.hills_thresh = float(hills_thresh)
        self.numK = num_kmers(self.kmin, self.kmax)

    def default_cap(self, max_len: int) -> int:
        """Non-overlapping hills of length >= min_m in max_len bases, plus room for tied hills."""
        return max(8, max_len // self.min_m + 16)

    def scan(self, bases: np.ndarray, offsets: np.ndarray, kmer_weights: np.ndarray, mode: int = 0,
             cap: int | None = None) -> HillsResult:
        """mode 0 = findSeqlets, 1 = findHills. Host buffers in and out."""
        lib = _lib.load()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        w = np.ascontiguousarray(kmer_weights, dtype=np.float64)
        if w.size != self.numK:
            raise ValueError(f"kmer_weights must hold numK = {self.numK} values")
        n = len(offsets) - 1
        max_len = int(np.max(np.diff(offsets))) if n > 0 else 0
        cap = int(cap) if cap is not None else self.default_cap(max_len)
        start = np.zeros((n, cap), dtype=np.int32)
        length = np.zeros((n, cap), dtype=np.int32)
        score = np.zeros((n, cap), dtype=np.float64)
        counts = np.zeros(n, dtype=np.int32)
        has_n = np.zeros(n, dtype=np.uint8)
        buf = bases if bases.size else np.zeros(1, dtype=np.uint8)
        rc = lib.sqw_hills_scan(buf.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(_lib.i64p), n,
                                self.kmin, self.kmax, w.ctypes.data_as(_lib.f64p), self.min_m,
                                self.max_m, self.hills_thresh, int(mode), cap,
                                start.ctypes.data_as(_lib.i32p), length.ctypes.data_as(_lib.i32p),
                                score.ctypes.data_as(_lib.f64p), counts.ctypes.data_as(_lib.i32p),
                                has_n.ctypes.data_as(_lib.u8p))
        if rc == _lib.SQW_ERR_BAD_BASE:
            raise ValueError("Nonstandard nucleotide character encountered")
        _lib.check(rc)
        return HillsResult(start, length, score, counts, has_n)

    def scan_device(self, bases, offsets, max_len: int, kmer_weights, mode: int = 0,
                    cap: int | None = None):
        """Device-resident form: cuda tensors in (uint8 bases, int64 offsets, float64 weights),
        cuda tensors out (start, length, score [n, cap]; counts, has_n [n]; status [1])."""
        import torch

        lib = _lib.load()
        dev = bases.device
        n = int(offsets.numel()) - 1
        cap = int(cap) if cap is not None else self.default_cap(int(max_len))
        start = torch.zeros((n, cap), dtype=torch.int32, device=dev)
        length = torch.zeros((n, cap), dtype=torch.int32, device=dev)
        score = torch.zeros((n, cap), dtype=torch.float64, device=dev)
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        has_n = torch.zeros(n, dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.sqw_hills_scan_dev(bases.data_ptr(), offsets.data_ptr(), n, int(max_len),
                                          self.kmin, self.kmax, kmer_weights.data_ptr(), self.min_m,
                                          self.max_m, self.hills_thresh, int(mode), cap,
                                          start.data_ptr(), length.data_ptr(), score.data_ptr(),
                                          counts.data_ptr(), has_n.data_ptr(), status.data_ptr(), stream))
        return start, length, score, counts, has_n, status

    # ---- the reference's two entry points -------------------------------------------------
    def findSeqlets(self, seqs: Sequence[str], kmer_weights: np.ndarray,
                    sort: bool = False) -> List[Tuple[str, float]]:
        """Discrim.java:382-444: [(substring, score)] in list order; ``sort`` applies
        sortseqlets (:241-249, stable, descending score)."""
        bases, offsets = pack_sequences(seqs)
        res = self.scan(bases, offsets, kmer_weights, mode=0)
        seq, st, ln, sc = res.flat()
        out = [(seqs[int(r)][int(s):int(s) + int(l)], float(v)) for r, s, l, v in zip(seq, st, ln, sc)]
        return sorted(out, key=lambda pr: -pr[1]) if sort else out

    def findHills(self, seqs: Sequence[str], region_starts: Sequence[int], kmer_weights: np.ndarray,
                  sort: bool = False) -> List[Tuple[int, int, int, float]]:
        """Discrim.java:328-380 on regions whose sequences are ``seqs`` (upper-cased like :336):
        [(region index, hill start, hill end (inclusive), score)] with genome coordinates
        r.getStart()+i .. r.getStart()+i+l-1 (:355)."""
        up = [s.upper() for s in seqs]
        bases, offsets = pack_sequences(up)
        res = self.scan(bases, offsets, kmer_weights, mode=1)
        seq, st, ln, sc = res.flat()
        out = [(int(r), int(region_starts[int(r)]) + int(s), int(region_starts[int(r)]) + int(s) + int(l) - 1,
                float(v)) for r, s, l, v in zip(seq, st, ln, sc)]
        return sorted(out, key=lambda t: -t[3]) if sort else out
