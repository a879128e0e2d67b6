"""Shared test helpers: build trainer inputs from the synthetic generator the way the reference's
host code does (MakeArff.setLabels -> KmerCounter -> RemoveUseless -> Classifier.buildClassifier
statistics), using the CPU oracle for the arithmetic. Test infrastructure only."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from oracle import pyoracle as O
from sequnwinder_b200 import structure, synth


@dataclass
class Problem:
    data: np.ndarray       # standardised, column 0 == 1
    raw: np.ndarray        # un-standardised, column 0 == 1
    cls: np.ndarray
    weights: np.ndarray
    st: O.StructureArrays
    crs: structure.ClassRelationStructure
    num_classes: int
    x0: np.ndarray
    smx0: np.ndarray
    mean: np.ndarray
    sd: np.ndarray
    threads: int
    counts: np.ndarray
    keep: np.ndarray


def make_problem(cfg: str = "cfg1", n_sites: int | None = None, max_features: int | None = None,
                 seed_offset: int = 0, outgroup: bool | None = None) -> Problem:
    c = synth.CONFIGS[cfg]
    d = synth.generate(n_sites or c["n_sites"], c["n_labels"],
                       seed=synth.SEED_BASE + int(cfg[-1]) + seed_offset,
                       outgroup=c["outgroup"] if outgroup is None else outgroup)
    counts, has_n = O.kmer_count(d.bases, d.offsets, c["kmin"], c["kmax"])
    assert not has_n.any()
    des = structure.make_design(d.annotations)
    crs = structure.ClassRelationStructure(des.design, des.num_layers)
    st = O.StructureArrays(**crs.to_arrays())
    cls = np.array([des.subgroup_names.index(s) for s in des.subgroups_at_sites], dtype=np.int32)
    w = np.array([des.subgroup_weights[s] for s in des.subgroups_at_sites], dtype=np.float64)
    keep = np.nonzero(counts.max(0) != counts.min(0))[0]   # weka RemoveUseless: constant columns
    if max_features is not None:
        keep = keep[:max_features]
    raw = np.concatenate([np.ones((counts.shape[0], 1)), counts[:, keep].astype(np.float64)], 1)
    data, mean, sd = O.standardize(raw, w)
    C = len(des.subgroup_names)
    x0, smx0 = O.init_params(cls, raw.shape[1], C, st)
    return Problem(data, raw, cls, w, st, crs, C, x0, smx0, mean, sd, c["threads"], counts, keep)
