"""sequnwinder_b200 — B200-native (sm_100a CUDA) drop-in for SeqUnwinder's two data-parallel
hot paths: k-mer featurisation (``KmerCounter``) and the consensus-ADMM trainer (``LOneCuda``,
an ``Optimizer`` plug-in). The compute lives in ``libsequnwinder_b200.so`` behind the C ABI in
``include/sequnwinder_b200.h``; this package is the host-side mirror of the reference's own
interfaces. There is no CPU fallback.
"""
from .kmer import KmerCounter, kmer_names, num_kmers, pack_sequences  # noqa: F401
from .structure import ClassRelationStructure, make_design  # noqa: F401
from .hills import HillsScanner  # noqa: F401

__all__ = ["KmerCounter", "kmer_names", "num_kmers", "pack_sequences",
           "ClassRelationStructure", "make_design", "HillsScanner"]
