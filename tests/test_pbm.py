"""pbm: marshalling ade graphs.  The codec is hand-written
(no protoc here); it is pinned by the reference's own fixture pbm/data/graph.pb /
graph.txt (copied to tests/golden/pbm/: test DATA of pbm/test/test_save.cpp:116-132 and
test_load.cpp:90-118, not source), which it must decode, print and re-encode to the byte."""
import os

import numpy as np
import pytest

import util
from util import age, llo, oracle

import workloads

GOLDEN = os.path.join(util.ROOT, "tests", "golden", "pbm")


def tree_lines(nodes, idx, depth=0, out=None):
    """the PrettyEquation rendering the reference's test compares (dbg/ade.hpp), trimmed"""
    out = [] if out is None else out
    node = nodes[idx]
    if "functor" in node:
        text = "(%s)" % node["functor"]["opname"]
    else:
        text = "([%s])" % "\\".join(str(d) for d in node["source"]["shape"])
    out.append(text if depth == 0 else "`--" + text)
    if "functor" in node:
        for arg in node["functor"]["args"]:
            tree_lines(nodes, arg["idx"], depth + 1, out)
    return out


def test_reference_fixture_decodes_prints_and_reencodes_to_the_byte():
    raw = open(os.path.join(GOLDEN, "graph.pb"), "rb").read()
    assert llo.pbm_reencode(raw) == raw
    graph = llo.pbm_decode(raw)
    nodes = graph["nodes"]
    assert len(nodes) == 18
    assert [n["labels"] for n in nodes if "source" in n] == [
        ["subtree", "src2"], ["global", "osrc"], ["subtree", "src"], ["global", "osrc2"],
        ["subtree2", "src"], ["subtree2", "src2"], ["subtree2", "src3"]]
    # sources come first, functors after their arguments (pbm/src/save.cpp:20-30)
    for i, n in enumerate(nodes):
        if "functor" in n:
            assert all(a["idx"] < i for a in n["functor"]["args"])
            assert all(len(a["coord"]) == 81 and len(a["shaper"]) in (0, 81) for a in n["functor"]["args"])
    dest = {tuple(n["labels"]): i for i, n in enumerate(nodes) if n["labels"]}
    got = tree_lines(nodes, dest[("subtree", "dest")]) + tree_lines(nodes, dest[("subtree2", "dest")])
    want = []
    for line in open(os.path.join(GOLDEN, "graph.txt")):
        line = line.strip().lstrip("| ")   # the tree rails carry no information beyond depth-first order
        if line:
            want.append(line)
    assert got == want
    # the permute({1,0}) shaper of the fixture's matmul argument is a real 9x9 map
    matmul = next(n for n in nodes if "functor" in n and n["functor"]["opname"] == "@")
    shaper = np.array(matmul["functor"]["args"][0]["shaper"]).reshape(9, 9)
    assert shaper[0][1] == 1 and shaper[1][0] == 1 and shaper[8][8] == 1


def test_llo_graph_round_trips_with_labels_and_values():
    X, Y, params = workloads.mlp_data(6, sizes=(5, 7, 4, 3), dtype=np.float64)
    loss, pvars, (vx, vy) = workloads.mlp_graph(age, llo, X, Y, params)
    grad = llo.derive(loss, pvars[0])
    data = llo.save_graph([loss, grad], [(loss, ["model", "loss"]), (grad, ["model", "dW1"]),
                                         (pvars[0], ["model", "W1"])], "mlp")
    assert llo.pbm_reencode(data) == data
    loaded = llo.load_graph(data)
    assert loaded["label"] == "mlp"
    by_path = {tuple(path): tens for path, tens in loaded["labelled"]}
    assert set(by_path) == {("model", "loss"), ("model", "dW1"), ("model", "W1")}
    assert len(loaded["roots"]) == 2
    for path, original in ((("model", "loss"), loss), (("model", "dW1"), grad)):
        want = oracle.evaluate(original, np.dtype(np.float64))
        got = oracle.evaluate(by_path[path], np.dtype(np.float64))
        np.testing.assert_array_equal(want, got)
    # leaves keep dtype, shape and bytes; the loaded leaf carries its last label
    w1 = by_path[("model", "W1")]
    np.testing.assert_array_equal(oracle.evaluate(w1, np.dtype(np.float64)), params[0])
    assert str(w1).startswith("W1(")
    # the loaded graph is an ordinary graph: it derives and optimizes
    again = llo.derive(by_path[("model", "loss")], w1)
    np.testing.assert_array_equal(oracle.evaluate(again, np.dtype(np.float64)),
                                  oracle.evaluate(grad, np.dtype(np.float64)))


def test_wide_dimensions_use_the_versioned_field_and_narrow_ones_the_reference_bytes():
    narrow = llo.variable(np.arange(6, dtype=np.int32).reshape(2, 3), "n")
    wide = llo.variable(np.arange(600, dtype=np.float32).reshape(2, 300), "w")
    root = age.add(narrow, narrow)
    data = llo.save_graph([root, wide])
    nodes = llo.pbm_decode(data)["nodes"]
    sources = [n["source"] for n in nodes if "source" in n]
    by_len = {len(s["data"]): s for s in sources}
    assert list(by_len[24]["shape"]) == [3, 2, 1, 1, 1, 1, 1, 1] and by_len[24]["wide_shape"] == []
    assert by_len[2400]["shape"] == b"" and by_len[2400]["wide_shape"] == [300, 2, 1, 1, 1, 1, 1, 1]
    loaded = llo.load_graph(data)
    shapes = sorted(tuple(int(d) for d in t.shape()) for t in loaded["nodes"] if not t.opname())
    assert shapes == [(2, 3), (2, 300)]
    back = [t for t in loaded["nodes"] if not t.opname() and tuple(t.shape()) == (2, 300)][0]
    np.testing.assert_array_equal(oracle.evaluate(back, np.dtype(np.float32)),
                                  np.arange(600, dtype=np.float32).reshape(2, 300))


def test_truncated_files_fail_loudly():
    data = llo.save_graph([age.neg(llo.variable(np.ones((2, 2)), "v"))])
    with pytest.raises(RuntimeError):
        llo.load_graph(data[:-5])
