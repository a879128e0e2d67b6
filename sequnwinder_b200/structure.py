"""Label hierarchy ("design file") handling on the host side of the trainer boundary.

Mirrors the two reference pieces that produce and consume the class structure:

* ``make_design``  — loaddata/MakeArff.java:51-232 (``setLabels``): subgroup names, the design
  text, per-subgroup instance weights.
* ``ClassRelationStructure`` — framework/Classifier.java:633-736: parser of that text.

All citations are relative to /root/reference/src/org/seqcode/projects/sequnwinder/.
The reference iterates ``java.util.HashMap`` key sets when it numbers labels
(MakeArff.java:97-101); ``_java_hashmap_order`` reproduces that order so that label node
indices match the Java host's.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np


def _java_string_hash(s: str) -> int:
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h


def _java_hashmap_order(keys_in_insertion_order: Sequence[str], max_size: int | None = None):
    """Iteration order of a Java 8 HashMap<String,?> (no treeified bins)."""
    n = max_size if max_size is not None else len(keys_in_insertion_order)
    cap = 16
    while n > (cap * 3) // 4:
        cap *= 2

    def bucket(k):
        h = _java_string_hash(k)
        h ^= h >> 16
        return h & (cap - 1)

    return [k for _, _, k in sorted((bucket(k), i, k)
                                    for i, k in enumerate(keys_in_insertion_order))]


@dataclass
class Node:
    """Classifier.java:702-735"""
    nodeIndex: int
    nodeName: str
    isLeaf: bool
    layer: int
    parents: List[int] = field(default_factory=list)
    children: List[int] = field(default_factory=list)


class ClassRelationStructure:
    """Parsed design file (Classifier.java:648-683).

    ``leafs`` and ``layers`` keep design-file order, as the reference's lists do; parents are
    recorded for leaves only and children for non-leaves only (Classifier.java:665,672).
    """

    def __init__(self, design_text: str, num_layers: int):
        self.numLayers = int(num_layers)
        self.leafs: List[Node] = []
        self.layers: List[List[Node]] = [[] for _ in range(self.numLayers)]
        self.allNodes: Dict[int, Node] = {}
        self.file_order: List[Node] = []
        for line in design_text.splitlines():
            if not line or line.startswith("#"):
                continue
            pieces = line.split("\t")
            if len(pieces) < 6:
                raise ValueError("Incorrect class structure format!!!")
            node = Node(int(pieces[0]), pieces[1], int(pieces[2]) == 1, int(pieces[5]))
            if node.isLeaf and pieces[3] != "-":
                node.parents = [int(s) for s in pieces[3].split(",")]
            if (not node.isLeaf) and pieces[4] != "-":
                node.children = [int(s) for s in pieces[4].split(",")]
            self.allNodes[node.nodeIndex] = node
            if node.isLeaf:
                self.leafs.append(node)
            self.layers[node.layer].append(node)
            self.file_order.append(node)

    @property
    def numNodes(self) -> int:
        return len(self.allNodes)

    def node_names(self) -> List[str]:
        return [self.allNodes[i].nodeName for i in sorted(self.allNodes)]

    def to_arrays(self):
        """CSR arrays in design-file order, the form the C ABI takes (include/sequnwinder_b200.h)."""
        fo = self.file_order
        node_index = np.array([n.nodeIndex for n in fo], dtype=np.int32)
        is_leaf = np.array([1 if n.isLeaf else 0 for n in fo], dtype=np.int32)
        layer = np.array([n.layer for n in fo], dtype=np.int32)
        parent_ptr = np.zeros(len(fo) + 1, dtype=np.int32)
        child_ptr = np.zeros(len(fo) + 1, dtype=np.int32)
        pidx: List[int] = []
        cidx: List[int] = []
        for i, n in enumerate(fo):
            pidx.extend(n.parents)
            cidx.extend(n.children)
            parent_ptr[i + 1] = len(pidx)
            child_ptr[i + 1] = len(cidx)
        return dict(num_nodes=len(fo), num_layers=self.numLayers, node_index=node_index,
                    is_leaf=is_leaf, layer=layer, parent_ptr=parent_ptr,
                    parent_idx=np.array(pidx, dtype=np.int32), child_ptr=child_ptr,
                    child_idx=np.array(cidx, dtype=np.int32))


@dataclass
class Design:
    design: str                    # the design-file text (MakeArff.java:144-208)
    num_layers: int                # 1 when no label survives (MakeArff.java:137-142)
    subgroup_names: List[str]      # index == leaf nodeIndex == class value
    subgroups_at_sites: List[str]  # per site, after renaming
    subgroup_weights: Dict[str, float]
    label_index: Dict[str, int]


def _matches(sub: str, lab: str) -> bool:
    # MakeArff.java:160,197
    return (sub.startswith(lab + "#") or sub.endswith("#" + lab) or ("#" + lab + "#") in sub
            or sub == lab)


def make_design(annotations: Sequence[str]) -> Design:
    """MakeArff.setLabels (MakeArff.java:51-232) for a list of ``"LabA;LabB"`` annotations."""
    label_names: List[str] = []          # insertion order of HashMap labelNames
    sub_names: List[str] = []
    subs_at: List[str] = []
    for s in annotations:
        labels = s.split(";")
        subgroup = "#".join(labels)
        if subgroup not in sub_names:
            sub_names.append(subgroup)
        for ln in labels:
            if ln not in label_names:
                label_names.append(ln)
        subs_at.append(subgroup)
    max_labels = len(label_names)

    # labels with a single subgroup are removed (MakeArff.java:78-93)
    cnt: Dict[str, int] = {}
    for sn in sub_names:
        for piece in sn.split("#"):
            cnt[piece] = cnt.get(piece, 0) + 1
    label_names = [l for l in label_names if cnt.get(l, 0) != 1]

    # label indices follow HashMap iteration order (MakeArff.java:97-101)
    ordered = _java_hashmap_order(label_names, max_labels)
    label_index = {s: i + len(sub_names) for i, s in enumerate(ordered)}

    # a subgroup named like a label becomes "<name>Only#<name>" (MakeArff.java:108-125)
    for i, sn in enumerate(sub_names):
        if sn in label_index:
            mod = sn + "Only#" + sn
            sub_names[i] = mod
            subs_at = [mod if v == sn else v for v in subs_at]

    num_layers = 1 if not label_index else 2

    lines = ["#Index\tName\tisSubGroup\tAssignedLabs\tAssignedSGs\tLayer"]
    for index, s in enumerate(sub_names):
        is_root = not any(_matches(s, lab) for lab in label_index)
        if not is_root:
            labs = ",".join(str(label_index[a]) for a in s.split("#") if a in label_index)
        else:
            labs = "-"
        lines.append(f"{index}\t{s}\t1\t{labs}\t-\t0")
    for lab in sorted(label_index, key=lambda k: label_index[k]):
        kids = ",".join(str(i) for i, sub in enumerate(sub_names) if _matches(sub, lab))
        lines.append(f"{label_index[lab]}\t{lab}\t0\t-\t{kids}\t1")
    design = "\n".join(lines)

    # instance weights: largest subgroup count / this subgroup's count (MakeArff.java:213-229)
    counts: Dict[str, float] = {}
    for s in subs_at:
        counts[s] = counts.get(s, 0.0) + 1.0
    biggest = max(counts.values()) if counts else 1.0
    weights = {k: biggest / v for k, v in counts.items()}
    for k, v in counts.items():
        if v == biggest:
            weights[k] = 1.0
    return Design(design, num_layers, sub_names, subs_at, weights, label_index)
