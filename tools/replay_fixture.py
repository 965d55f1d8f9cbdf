#!/usr/bin/env python
"""Replay one of the reference's own test fixtures through the B200 path.

The reference tests (tests/test_broad_phase.cpp:14-78, tests/test_broad_phase.cu:22-126,
tests/test_narrow_phase.cu:17-72) read two PLY frames of a simulation with libigl, derive the
edges with igl::edges, and compare against Mathematica ground-truth JSON files from
Continuous-Collision-Detection/Sample-Scalable-CCD-Data (not available offline).  With the data:

  python tools/replay_fixture.py cloth-ball/frames/cloth_ball92.ply cloth-ball/frames/cloth_ball93.ply \
      --vf-gt cloth-ball/boxes/92vf.json --ee-gt cloth-ball/boxes/92ee.json \
      --expect-vf 1655541 --expect-ee 5197332 --expect-toi 3.814697265625e-06

checks what those tests check: box counts, exact overlap counts, ground truth is a subset of the
result (ground_truth.cpp:55-63, with the id offsets of test_broad_phase.cpp:66-74) and the
earliest TOI (test_narrow_phase.cu:41-45,65: ms=0, max_iter=-1, tol=1e-6, allow_zero_toi).

The file readers are plain numpy (no libigl here): PLY ascii / binary_little_endian with float or
double vertices and triangle faces; igl_edges() restates igl::edges (adjacency matrix, then the
upper-triangular entries in column-major order: edges (i, j), i < j, sorted by j then i).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4",
              "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply(path):
    """-> (V float64 (n, 3), F int32 (m, 3)); triangle meshes only."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1]["props"].append(("list", tok[2], tok[3], tok[4]))
                else:
                    elements[-1]["props"].append(("scalar", tok[1], tok[2]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        V = F = None
        for el in elements:
            n, props = el["count"], el["props"]
            if all(p[0] == "scalar" for p in props):
                names = [p[2] for p in props]
                if fmt == "ascii":
                    rows = np.array([f.readline().split() for _ in range(n)], dtype=np.float64)
                    rows = rows.reshape(n, len(props))
                    cols = {nm: rows[:, i] for i, nm in enumerate(names)}
                else:
                    dt = np.dtype([(nm, "<" + _PLY_TYPES[p[1]]) for nm, p in zip(names, props)])
                    rec = np.frombuffer(f.read(dt.itemsize * n), dtype=dt, count=n)
                    cols = {nm: rec[nm].astype(np.float64) for nm in names}
                if el["name"] == "vertex":
                    V = np.stack([cols["x"], cols["y"], cols["z"]], axis=1)
            else:
                if len(props) != 1:
                    raise ValueError(f"{path}: element {el['name']}: mixed list properties")
                _, ct, it, _ = props[0]
                faces = []
                if fmt == "ascii":
                    for _ in range(n):
                        t = f.readline().split()
                        faces.append([int(x) for x in t[1:1 + int(t[0])]])
                else:
                    cdt, idt = np.dtype("<" + _PLY_TYPES[ct]), np.dtype("<" + _PLY_TYPES[it])
                    for _ in range(n):
                        k = int(np.frombuffer(f.read(cdt.itemsize), dtype=cdt)[0])
                        faces.append(np.frombuffer(f.read(idt.itemsize * k), dtype=idt).tolist())
                if el["name"] == "face":
                    if any(len(fc) != 3 for fc in faces):
                        raise ValueError(f"{path}: only triangle meshes are supported")
                    F = np.array(faces, dtype=np.int32).reshape(-1, 3)
        if V is None or F is None:
            raise ValueError(f"{path}: needs a vertex and a face element")
        return V, F


def write_ply(path, V, F, binary=True, dtype="double"):
    """Small writer (tests / making fixtures from synthetic scenes)."""
    V = np.asarray(V, dtype=np.float64)
    F = np.asarray(F, dtype=np.int32)
    head = ["ply", "format " + ("binary_little_endian" if binary else "ascii") + " 1.0",
            "comment written by tools/replay_fixture.py", f"element vertex {len(V)}",
            f"property {dtype} x", f"property {dtype} y", f"property {dtype} z",
            f"element face {len(F)}", "property list uchar int vertex_indices", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode())
        if binary:
            f.write(V.astype("<" + _PLY_TYPES[dtype]).tobytes())
            rec = np.zeros(len(F), dtype=[("n", "u1"), ("v", "<i4", 3)])
            rec["n"] = 3
            rec["v"] = F
            f.write(rec.tobytes())
        else:
            for v in V:
                f.write(("%r %r %r\n" % tuple(float(x) for x in v)).encode())
            for t in F:
                f.write(("3 %d %d %d\n" % tuple(int(x) for x in t)).encode())


def igl_edges(F):
    """igl::edges(F, E): unique undirected edges (i < j), ordered by j, then i."""
    F = np.asarray(F)
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    e = np.unique(np.sort(e, axis=1), axis=0)
    return e[np.lexsort((e[:, 0], e[:, 1]))].astype(np.int32)


def ground_truth_missing(pairs, gt_file, offset_a=0, offset_b=0):
    """Ground-truth pairs (JSON list of [a, b]) that are NOT in `pairs` after adding the id
    offsets of test_broad_phase.cpp:66-74; the reference requires this to be empty."""
    gt = np.array(json.load(open(gt_file)), dtype=np.int64).reshape(-1, 2)
    mine = np.asarray(pairs, dtype=np.int64).reshape(-1, 2) + np.array([offset_a, offset_b])
    key = lambda a: a[:, 0] * (1 << 32) + a[:, 1]
    return gt[~np.isin(key(gt), key(mine))]


def replay(file_t0, file_t1, vf_gt=None, ee_gt=None, params=None, device=0):
    """Runs the GPU path on the two frames; returns a dict with everything the reference's
    tests look at."""
    from _pkg import load_package
    sccd = load_package()
    V0, F = read_ply(file_t0)
    V1, F1 = read_ply(file_t1)
    if V0.shape != V1.shape or not np.array_equal(F, F1):
        raise ValueError("the two frames are not the same mesh")
    E = igl_edges(F)
    params = params or dict(ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True)
    fo = lambda a: np.asfortranarray(a)
    ctx = sccd.Context(device)
    ctx.upload_mesh(fo(V0), fo(V1), fo(E), fo(F))
    toi = ctx.ccd(**params)
    ctx.build_boxes(params["ms"])
    vf, ee = ctx.broad_phase(0), ctx.broad_phase(1)
    ctx.close()
    nV, nE = len(V0), len(E)
    out = {"n_vertices": nV, "n_edges": nE, "n_faces": len(F), "n_vf": len(vf), "n_ee": len(ee),
           "toi": toi}
    if vf_gt:   # (vertex, face + nV + nE)
        out["vf_gt_missing"] = len(ground_truth_missing(vf, vf_gt, 0, nV + nE))
    if ee_gt:   # (edge + nV, edge + nV)
        out["ee_gt_missing"] = len(ground_truth_missing(ee, ee_gt, nV, nV))
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    ap.add_argument("frame_t0")
    ap.add_argument("frame_t1")
    ap.add_argument("--vf-gt")
    ap.add_argument("--ee-gt")
    ap.add_argument("--expect-vf", type=int)
    ap.add_argument("--expect-ee", type=int)
    ap.add_argument("--expect-toi", type=float)
    a = ap.parse_args()
    r = replay(a.frame_t0, a.frame_t1, a.vf_gt, a.ee_gt)
    print(json.dumps(r))
    ok = r.get("vf_gt_missing", 0) == 0 and r.get("ee_gt_missing", 0) == 0
    if a.expect_vf is not None:
        ok = ok and r["n_vf"] == a.expect_vf
    if a.expect_ee is not None:
        ok = ok and r["n_ee"] == a.expect_ee
    if a.expect_toi is not None:   # test_narrow_phase.cu:65 uses Catch::Approx
        ok = ok and abs(r["toi"] - a.expect_toi) <= 1e-12 + 1e-10 * abs(a.expect_toi)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
