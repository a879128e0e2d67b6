"""Deterministic synthetic datasets of the shapes BASELINE.json names.

Follows the recipe of the reference's simulator (utils/SimulateBindingSite.java:40-130:
background sequence, then each of a site's label motifs inserted with a fixed probability at a
fixed offset) with an i.i.d. background instead of the genome Markov model, which needs files
that are not available offline. Labels form a ring: P labels ``L0..L{P-1}`` and C = P
sub-classes ``Li;L(i+1 mod P)``, so every label keeps two children (MakeArff.java:78-93) and
every leaf has two parents; ``outgroup=True`` appends a parentless ``Random`` sub-class
(the ``--makerandregs`` case, SeqUnwinderConfig.java:380-386).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

_ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)
_BG = np.array([0.3, 0.2, 0.2, 0.3])


@dataclass
class SynthData:
    bases: np.ndarray        # uint8 ASCII, n_sites*length, concatenated
    offsets: np.ndarray      # int64 [n_sites+1]
    annotations: List[str]   # per site, "Li;Lj" or "Random"
    site_class: np.ndarray   # int32 [n_sites] index into class_annotations
    class_annotations: List[str]

    @property
    def n_sites(self) -> int:
        return len(self.offsets) - 1


CONFIGS = {
    # name: (n_sites, n_labels, kmin, kmax, outgroup, T)
    "cfg1": dict(n_sites=2000, n_labels=3, kmin=4, kmax=5, outgroup=False, threads=5),
    "cfg2": dict(n_sites=20000, n_labels=8, kmin=4, kmax=5, outgroup=False, threads=8),
    "cfg3": dict(n_sites=200000, n_labels=12, kmin=4, kmax=7, outgroup=False, threads=8),
    "cfg4": dict(n_sites=1999999, n_labels=16, kmin=4, kmax=6, outgroup=True, threads=8),
}
SEED_BASE = 20260119


def generate(n_sites: int, n_labels: int, length: int = 150, seed: int = SEED_BASE,
             outgroup: bool = False, n_frac: float = 0.0, motif_len: int = 8,
             insert_prob: float = 0.7) -> SynthData:
    rng = np.random.default_rng(seed)
    P = n_labels
    class_ann = [f"L{i};L{(i + 1) % P}" for i in range(P)]
    class_labels = [[i, (i + 1) % P] for i in range(P)]
    if outgroup:
        class_ann.append("Random")
        class_labels.append([])
    C = len(class_ann)

    # balanced classes in a seeded random order (sites of one class must not all fall into
    # the same ADMM block, which i % T blocking would do for a round-robin order with T == C)
    site_class = (np.arange(n_sites) % C).astype(np.int32)
    rng.shuffle(site_class)

    seq = _ALPHABET[rng.choice(4, size=(n_sites, length), p=_BG)]

    # one PPM per label, Dirichlet(0.3) columns
    ppm = rng.dirichlet(np.full(4, 0.3), size=(P, motif_len))
    for c in range(C):
        rows = np.nonzero(site_class == c)[0]
        labs = class_labels[c]
        m = len(labs)
        for j, lab in enumerate(labs):
            take = rows[rng.random(rows.size) < insert_prob]
            off = (j * length) // m + 15
            if take.size == 0 or off + motif_len > length:
                continue
            u = rng.random((take.size, motif_len))
            cdf = np.cumsum(ppm[lab], axis=1)           # [motif_len, 4]
            letters = (u[:, :, None] > cdf[None, :, :]).sum(axis=2).clip(0, 3)
            seq[take[:, None], off + np.arange(motif_len)[None, :]] = _ALPHABET[letters]

    if n_frac > 0:
        bad = np.nonzero(rng.random(n_sites) < n_frac)[0]
        pos = rng.integers(0, length, size=bad.size)
        seq[bad, pos] = ord("N")

    offsets = (np.arange(n_sites + 1, dtype=np.int64) * length)
    annotations = [class_ann[c] for c in site_class]
    return SynthData(np.ascontiguousarray(seq).reshape(-1), offsets, annotations, site_class,
                     class_ann)


def generate_config(name: str, n_sites: int | None = None, **kw) -> SynthData:
    cfg = CONFIGS[name]
    idx = int(name[-1])
    return generate(n_sites or cfg["n_sites"], cfg["n_labels"], seed=SEED_BASE + idx,
                    outgroup=cfg["outgroup"], **kw)
