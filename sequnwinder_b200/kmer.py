"""Host-side mirror of the reference's k-mer featuriser boundary.

``KmerCounter`` plays the role of the inner class ``MakeArff.KmerCounter``
(loaddata/MakeArff.java:314-377, relative to
/root/reference/src/org/seqcode/projects/sequnwinder/): same inputs (upper-case sequences, k
range), same outputs (dense numK-wide int rows in ``int2seq`` column order, rows with ``N``
omitted, the ``tmpCounts.mat`` text when asked for). The counting itself is the CUDA kernel
behind ``sqw_kmer_count`` / ``sqw_kmer_count_dev``; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Sequence, Tuple

import numpy as np

from . import _lib


def num_kmers(kmin: int, kmax: int) -> int:
    """numK = sum 4^k (MakeArff.java:325-328)."""
    v = int(_lib.load().sqw_num_kmers(kmin, kmax))
    if v < 0:
        raise ValueError(f"invalid k range [{kmin},{kmax}]")
    return v


def int2seq(value: int, k: int) -> str:
    """seqcode-core RegionFileUtilities.int2seq: A0 C1 G2 T3, first base most significant."""
    return "".join("ACGT"[(value >> (2 * (k - 1 - i))) & 3] for i in range(k))


def kmer_names(kmin: int, kmax: int) -> List[str]:
    """Header of tmpCounts.mat (MakeArff.java:336-340)."""
    return [int2seq(i, k) for k in range(kmin, kmax + 1) for i in range(4 ** k)]


def pack_sequences(seqs: Iterable) -> Tuple[np.ndarray, np.ndarray]:
    """list of str/bytes -> (concatenated uint8 bases, int64 offsets[n+1])."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs), dtype=np.uint8).copy(), offsets


class KmerCounter:
    """Drop-in for ``new KmerCounter().printPeakSeqKmerRange(kmin, kmax)``."""

    def __init__(self, kmin: int = 4, kmax: int = 5):
        self.kmin, self.kmax = int(kmin), int(kmax)
        self.numK = num_kmers(self.kmin, self.kmax)

    # ---- host buffers in, host buffers out: what the Java host calls through JNI ----
    def count(self, bases: np.ndarray, offsets: np.ndarray, out: np.ndarray | None = None):
        """Returns (counts[n, numK] int32, has_n[n] uint8); rows with has_n==1 are zero and must
        be omitted by the caller (MakeArff.java:354-355)."""
        lib = _lib.load()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        counts = out if out is not None else np.empty((n, self.numK), dtype=np.int32)
        assert counts.dtype == np.int32 and counts.flags.c_contiguous
        has_n = np.empty(n, dtype=np.uint8)
        buf = bases if bases.size else np.zeros(1, dtype=np.uint8)
        rc = lib.sqw_kmer_count(buf.ctypes.data_as(C.c_void_p),
                                offsets.ctypes.data_as(_lib.i64p), n, self.kmin, self.kmax,
                                counts.ctypes.data_as(_lib.i32p), has_n.ctypes.data_as(_lib.u8p))
        if rc == _lib.SQW_ERR_BAD_BASE:
            raise ValueError("Nonstandard nucleotide character encountered")
        _lib.check(rc)
        return counts, has_n

    def count_sequences(self, seqs: Sequence):
        """Sequences as Python strings (upper-cased like SeqUnwinderConfig.java:296)."""
        bases, offsets = pack_sequences([s.upper() if isinstance(s, str) else s for s in seqs])
        return self.count(bases, offsets)

    # ---- device resident: torch tensors on cuda, asynchronous on the current stream ----
    def count_device(self, bases, offsets, max_len: int, out=None, has_n=None, status=None):
        """bases: cuda uint8 tensor, offsets: cuda int64 tensor [n+1]. Returns
        (counts cuda int32 [n, numK], has_n cuda uint8 [n], status cuda int32 [1]); status != 0
        means a character outside ACGTN was met."""
        import torch

        lib = _lib.load()
        n = offsets.numel() - 1
        dev = bases.device
        if out is None:
            out = torch.empty((n, self.numK), dtype=torch.int32, device=dev)
        if has_n is None:
            has_n = torch.empty(n, dtype=torch.uint8, device=dev)
        if status is None:
            status = torch.zeros(1, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.sqw_kmer_count_dev(bases.data_ptr(), offsets.data_ptr(), n, int(max_len),
                                    self.kmin, self.kmax, out.data_ptr(), has_n.data_ptr(),
                                    status.data_ptr(), C.c_void_p(stream))
        _lib.check(rc)
        return out, has_n, status

    # ---- the reference's on-disk artefact ----
    def printPeakSeqKmerRange(self, seqs: Sequence[str], weights: Sequence[float],
                              labels: Sequence[str], path: str) -> int:
        """Writes ``tmpCounts.mat`` exactly as MakeArff.java:331-376 does (tab-separated header
        of k-mer names + Weight + Label, one row per N-free sequence). Returns rows written."""
        counts, has_n = self.count_sequences(seqs)
        written = 0
        with open(path, "w") as fh:
            fh.write("\t".join(kmer_names(self.kmin, self.kmax)) + "\tWeight\tLabel\n")
            for r in range(len(seqs)):
                if has_n[r]:
                    continue
                fh.write("\t".join(map(str, counts[r].tolist())))
                fh.write(f"\t{_java_double(weights[r])}\t{labels[r]}\n")
                written += 1
        return written


def _java_double(v: float) -> str:
    """Java's Double.toString for the values that occur as subgroup weights."""
    v = float(v)
    if v == int(v) and abs(v) < 1e7:
        return f"{int(v)}.0"
    return repr(v)
