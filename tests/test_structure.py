"""CPU: host-side mirrors of MakeArff.setLabels (loaddata/MakeArff.java:51-232) and
Classifier.ClassRelationStructure (framework/Classifier.java:633-736)."""
import numpy as np

from sequnwinder_b200.structure import ClassRelationStructure, make_design, _java_string_hash


def test_java_string_hash_known_values():
    assert _java_string_hash("") == 0
    assert _java_string_hash("a") == 97
    assert _java_string_hash("Thread0") == (sum(ord(c) * 31 ** (6 - i) for i, c in enumerate("Thread0"))) & 0xFFFFFFFF


def test_ring_design():
    ann = ["L0;L1", "L1;L2", "L2;L0", "L0;L1", "L2;L0"]
    d = make_design(ann)
    assert d.num_layers == 2
    assert d.subgroup_names == ["L0#L1", "L1#L2", "L2#L0"]
    crs = ClassRelationStructure(d.design, d.num_layers)
    assert crs.numNodes == 6 and [n.nodeIndex for n in crs.leafs] == [0, 1, 2]
    for leaf in crs.leafs:
        assert len(leaf.parents) == 2 and not leaf.children
        for pid in leaf.parents:
            assert leaf.nodeIndex in crs.allNodes[pid].children
    for lab in crs.layers[1]:
        assert not lab.parents and len(lab.children) == 2      # Classifier.java:665: leaves only
    assert d.subgroup_weights["L0#L1"] == 1.0 and d.subgroup_weights["L1#L2"] == 2.0


def test_singleton_label_removed_and_outgroup_is_root():
    ann = ["A;B", "A;C", "Random"]
    d = make_design(ann)
    # B, C and Random each have one subgroup -> removed (MakeArff.java:78-93); A stays
    assert set(d.label_index) == {"A"}
    crs = ClassRelationStructure(d.design, d.num_layers)
    leaf = {n.nodeName: n for n in crs.leafs}
    assert leaf["Random"].parents == []
    assert leaf["A#B"].parents == [d.label_index["A"]]


def test_subgroup_named_like_label_is_renamed():
    ann = ["A", "A;B", "B;C", "C"]
    d = make_design(ann)
    assert "AOnly#A" in d.subgroup_names and "COnly#C" in d.subgroup_names
    assert "AOnly#A" in d.subgroups_at_sites


def test_no_labels_means_single_layer():
    d = make_design(["X", "Y", "X"])
    assert d.num_layers == 1 and not d.label_index


def test_to_arrays_roundtrip():
    d = make_design(["L0;L1", "L1;L2", "L2;L0"])
    crs = ClassRelationStructure(d.design, d.num_layers)
    a = crs.to_arrays()
    assert a["num_nodes"] == 6 and a["parent_ptr"][-1] == 6 and a["child_ptr"][-1] == 6
    assert a["is_leaf"].tolist() == [1, 1, 1, 0, 0, 0]
    assert np.array_equal(a["layer"], [0, 0, 0, 1, 1, 1])
