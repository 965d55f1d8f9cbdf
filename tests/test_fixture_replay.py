"""tools/replay_fixture.py: the PLY reader, the igl::edges restatement and the ground-truth
comparer the reference's tests are built on (tests/io.cpp, tests/ground_truth.cpp), on a
synthetic fixture written here -- the upstream sample data is not available offline."""
import importlib.util
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("replay_fixture", os.path.join(ROOT, "tools", "replay_fixture.py"))
rf = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rf)


@pytest.mark.parametrize("binary,dtype", [(True, "double"), (True, "float"), (False, "double")])
def test_ply_round_trip(tmp_path, scene_small, binary, dtype):
    V, F = np.ascontiguousarray(scene_small["V0"]), np.ascontiguousarray(scene_small["F"])
    p = str(tmp_path / "a.ply")
    rf.write_ply(p, V, F, binary=binary, dtype=dtype)
    V2, F2 = rf.read_ply(p)
    assert np.array_equal(F2, F)
    if dtype == "double":
        assert np.array_equal(V2, V)                       # doubles survive bit for bit
    else:
        assert np.array_equal(V2, V.astype(np.float32).astype(np.float64))


def test_igl_edges_order(scene_small):
    F = np.ascontiguousarray(scene_small["F"])
    E = rf.igl_edges(F)
    want = set()
    for a, b, c in F.tolist():
        for i, j in ((a, b), (b, c), (c, a)):
            want.add((min(i, j), max(i, j)))
    assert set(map(tuple, E.tolist())) == want and len(E) == len(want)
    assert np.all(E[:, 0] < E[:, 1])
    key = E[:, 1].astype(np.int64) * (1 << 32) + E[:, 0]   # column-major upper triangle
    assert np.all(np.diff(key) > 0)


def test_ground_truth_subset_check(tmp_path):
    pairs = np.array([[0, 1], [2, 5], [3, 4]])
    gt = tmp_path / "gt.json"
    gt.write_text(json.dumps([[10, 21], [12, 25]]))
    assert len(rf.ground_truth_missing(pairs, str(gt), 10, 20)) == 0
    gt.write_text(json.dumps([[10, 21], [12, 26]]))
    miss = rf.ground_truth_missing(pairs, str(gt), 10, 20)
    assert miss.tolist() == [[12, 26]]


@pytest.mark.gpu
def test_replay_of_a_synthetic_fixture(tmp_path, orc, scene_small):
    """Two PLY frames + ground-truth JSON (the oracle's overlaps, offset like the reference's
    Mathematica files) replayed through the GPU path."""
    s = scene_small
    F = np.ascontiguousarray(s["F"])
    rf.write_ply(str(tmp_path / "f0.ply"), s["V0"], F)
    rf.write_ply(str(tmp_path / "f1.ply"), s["V1"], F)
    E = rf.igl_edges(F)
    scene = {"V0": s["V0"], "V1": s["V1"], "F": s["F"], "E": np.asfortranarray(E)}
    want = orc.ccd(scene)
    nV, nE = len(s["V0"]), len(E)
    (tmp_path / "vf.json").write_text(json.dumps((want["vf"] + [0, nV + nE]).tolist()))
    (tmp_path / "ee.json").write_text(json.dumps((want["ee"] + [nV, nV]).tolist()))
    r = rf.replay(str(tmp_path / "f0.ply"), str(tmp_path / "f1.ply"), str(tmp_path / "vf.json"),
                  str(tmp_path / "ee.json"))
    assert r["n_vf"] == len(want["vf"]) and r["n_ee"] == len(want["ee"])
    assert r["vf_gt_missing"] == 0 and r["ee_gt_missing"] == 0
    assert r["toi"] == want["toi"]
