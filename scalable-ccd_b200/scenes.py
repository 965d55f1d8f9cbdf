"""Deterministic synthetic two-frame scenes for the five BASELINE.json configs.

The reference's own fixtures (cloth-ball 92->93 etc., tests/test_broad_phase.cpp:23-38)
are downloaded PLY files that do not exist offline, so parity and benchmarks run on
these generators.  All arrays follow the reference's
conventions (cuda/ccd.cuh:26-38): V0/V1 are nV x 3 float64, E is nE x 2 int32,
F is nF x 3 int32, all COLUMN-major (Eigen default), i.e. numpy order="F".
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- meshes
def edges_from_faces(F: np.ndarray) -> np.ndarray:
    """Unique undirected edges of a triangle list (what igl::edges yields,
    reference tests/io.cpp:10-22), rows sorted lexicographically."""
    F = np.asarray(F, dtype=np.int64)
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    e = np.unique(e, axis=0)
    return e.astype(np.int32)


def grid_cloth(n: int):
    """n x n vertex grid on [-1,1]^2 in the xy plane, 2(n-1)^2 triangles."""
    xs = np.linspace(-1.0, 1.0, n)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    V = np.stack([X.ravel(), Y.ravel(), np.zeros(n * n)], axis=1)
    idx = np.arange(n * n).reshape(n, n)
    a = idx[:-1, :-1].ravel()
    b = idx[1:, :-1].ravel()
    c = idx[1:, 1:].ravel()
    d = idx[:-1, 1:].ravel()
    F = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], axis=0)
    return V, F.astype(np.int32)


def uv_sphere(nseg: int, nring: int, r: float = 1.0):
    """UV sphere: 2 + (nring-1)*nseg vertices, 2*nseg*(nring-1) triangles."""
    V = [[0.0, 0.0, r]]
    for i in range(1, nring):
        th = np.pi * i / nring
        for j in range(nseg):
            ph = 2 * np.pi * j / nseg
            V.append([r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)])
    V.append([0.0, 0.0, -r])
    V = np.array(V)
    F = []
    ring = lambda i, j: 1 + (i - 1) * nseg + (j % nseg)
    south = len(V) - 1
    for j in range(nseg):
        F.append([0, ring(1, j), ring(1, j + 1)])
        F.append([south, ring(nring - 1, j + 1), ring(nring - 1, j)])
    for i in range(1, nring - 1):
        for j in range(nseg):
            F.append([ring(i, j), ring(i + 1, j), ring(i + 1, j + 1)])
            F.append([ring(i, j), ring(i + 1, j + 1), ring(i, j + 1)])
    return V, np.array(F, dtype=np.int32)


def icosphere(level: int, r: float = 1.0):
    t = (1.0 + np.sqrt(5.0)) / 2.0
    V = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
         [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]],
        dtype=np.float64)
    F = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
         [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
         [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    for _ in range(level):
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        uniq, inv = np.unique(es, axis=0, return_inverse=True)
        mid = V[uniq[:, 0]] + V[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        nV = len(V)
        V = np.concatenate([V, mid], axis=0)
        nF = len(F)
        m01 = nV + inv[:nF]
        m12 = nV + inv[nF:2 * nF]
        m20 = nV + inv[2 * nF:]
        F = np.concatenate([
            np.stack([F[:, 0], m01, m20], 1), np.stack([F[:, 1], m12, m01], 1),
            np.stack([F[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return V * r, F.astype(np.int32)


def _pack(V0, V1, F, E=None):
    if E is None:
        E = edges_from_faces(F)
    return {
        "V0": np.asfortranarray(V0, dtype=np.float64),
        "V1": np.asfortranarray(V1, dtype=np.float64),
        "E": np.asfortranarray(E, dtype=np.int32),
        "F": np.asfortranarray(F, dtype=np.int32),
    }


# --------------------------------------------------------------------------- configs
def cloth_on_sphere(n: int, seed: int, sphere: str = "uv", drape: float = 0.0,
                    z0: float = 0.52, z1: float = 0.44):
    """Cloth grid (n x n) falling from z0 to z1 onto a static sphere of radius 0.5.
    C1 = (n=101, seed 0, UV sphere 24x12); C2 = (n=409, seed 1, icosphere level 5,
    draped)."""
    rng = np.random.default_rng(seed)
    Vc, Fc = grid_cloth(n)
    shape = drape * np.sin(3.0 * Vc[:, 0]) * np.cos(2.0 * Vc[:, 1])
    V0c = Vc.copy()
    V0c[:, 2] = z0 + shape + rng.uniform(-1e-3, 1e-3, len(Vc))
    V1c = Vc.copy() + rng.uniform(-1e-3, 1e-3, Vc.shape)
    V1c[:, 2] += z1 + shape
    if sphere == "uv":
        Vs, Fs = uv_sphere(24, 12, 0.5)
    else:
        Vs, Fs = icosphere(5, 0.5)
    # The sphere deforms slightly (per-vertex jitter at t1): for a perfectly rigid or
    # static body the reference's edge-edge tolerance degenerates (tol[0] = tol[1] = inf,
    # root_finder.cu:82-87) and every nearby non-incident edge pair reports toi = 0,
    # which would make the narrow phase trivial.
    Vs1 = Vs + rng.uniform(-1e-4, 1e-4, Vs.shape)
    V0 = np.concatenate([V0c, Vs], axis=0)
    V1 = np.concatenate([V1c, Vs1], axis=0)
    F = np.concatenate([Fc, Fs + len(Vc)], axis=0)
    return _pack(V0, V1, F)


def scene_c1(seed: int = 0):
    """Config 1: ~20K-triangle cloth over a UV sphere, ~62K boxes."""
    return cloth_on_sphere(101, seed, "uv")


def scene_c2(seed: int = 1, n: int = 409):
    """Config 2: ~1M-primitive cloth-ball scene (167,281+ V / 332,928+ F / 500,208+ E)."""
    return cloth_on_sphere(n, seed, "ico", drape=0.02)


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def blob_pile(n_inst: int, seed: int, slab: bool = False, radius: float = 0.05,
              spacing: float = 2.6):
    """Configs 3/4: n_inst rigid instances of a 602-primitive closed blob
    (noise-displaced UV sphere: 102 V / 300 E / 200 F), heavy-tailed placement
    (80% dense Gaussian core, 20% Pareto-tailed shell), random rotations, rigid
    velocity toward the core.  slab=True stretches the pile along x (config 4) so a
    sorted-range split across GPUs is meaningful.  `spacing` (in blob radii) sets the
    core density and hence the overlap count."""
    rng = np.random.default_rng(seed)
    Vb, Fb = uv_sphere(10, 11, 1.0)
    Vb = Vb * (1.0 + 0.25 * rng.uniform(-1, 1, (len(Vb), 1)))
    Eb = edges_from_faces(Fb)
    nvb, nfb, neb = len(Vb), len(Fb), len(Eb)

    n_core = int(0.8 * n_inst)
    n_tail = n_inst - n_core
    # Centres sit on jittered cubic lattices so that NO two blobs touch at t0 (a blob's
    # extent is at most 1.25 * radius; an initial interpenetration would make toi = 0 and
    # the narrow phase trivial): the dense core on a lattice of `spacing` radii -- the
    # lattice points closest to the origin, an ellipsoid 12:1:1 for the slab -- and the
    # sparse shell on a 3x coarser lattice with Pareto-distributed radii.
    stretch = np.array([12.0 if slab else 1.0, 1.0, 1.0])

    def lattice(n, h, r_min):
        k = int(np.ceil((2.0 * n / stretch[0]) ** (1.0 / 3.0))) + 2
        gx = np.arange(-int(k * stretch[0]), int(k * stretch[0]) + 1)
        gy = np.arange(-k, k + 1)
        X, Y, Z = np.meshgrid(gx, gy, gy, indexing="ij")
        pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1).astype(np.float64) * h
        r = np.linalg.norm(pts / stretch, axis=1)
        pts, r = pts[r >= r_min], r[r >= r_min]
        return pts[np.argsort(r, kind="stable")[:n]]

    h = spacing * radius
    core = lattice(n_core, h, 0.0)
    r_core = np.linalg.norm(core / stretch, axis=1).max() + 2.0 * h
    shell = lattice(8 * n_tail, 3.0 * h, r_core)
    # heavy tail: keep shell sites with probability falling off like a Pareto law
    rs = np.linalg.norm(shell / stretch, axis=1)
    keep = rng.random(len(shell)) < (r_core / rs) ** 2.5
    shell = shell[keep][:n_tail]
    if len(shell) < n_tail:  # not enough accepted sites: top up with the nearest rejected ones
        extra = lattice(8 * n_tail, 3.0 * h, r_core)[: n_tail - len(shell)]
        shell = np.concatenate([shell, extra + 1.5 * h], axis=0)
    C = np.concatenate([core, shell], axis=0)
    C += rng.uniform(-0.04, 0.04, C.shape) * radius
    C = C[rng.permutation(n_inst)]
    R = _random_rotations(rng, n_inst)
    P = np.einsum("nij,vj->nvi", R, Vb * radius) + C[:, None, :]
    target = C.copy()
    target[:, 1:] = 0.0
    if not slab:
        target[:, 0] = 0.0
    vel = (target - C)
    vel *= (0.6 * radius) / np.maximum(np.linalg.norm(vel, axis=1, keepdims=True), 1e-12)
    vel *= rng.uniform(0.2, 1.0, (n_inst, 1))
    V0 = P.reshape(-1, 3)
    # rigid velocity + a small non-rigid wobble (see cloth_on_sphere for why)
    V1 = (P + vel[:, None, :] + rng.uniform(-2e-3, 2e-3, P.shape) * radius).reshape(-1, 3)
    off = (np.arange(n_inst, dtype=np.int64) * nvb)[:, None, None]
    F = (Fb[None].astype(np.int64) + off).reshape(-1, 3).astype(np.int32)
    E = (Eb[None].astype(np.int64) + off).reshape(-1, 2).astype(np.int32)
    return _pack(V0, V1, F, E)


def scene_c3(seed: int = 2, n_inst: int = 10_000):
    """Config 3: ~6.0M AABBs, dense heavy-tailed overlaps."""
    return blob_pile(n_inst, seed)


def scene_c4(seed: int = 3, n_inst: int = 83_000):
    """Config 4: ~50M AABBs in a slab along x for 2/4/8-GPU range splitting."""
    return blob_pile(n_inst, seed, slab=True)


def queries_c5(n: int, seed: int = 4, parallel: bool = True):
    """Config 5: adversarial narrow-phase queries given directly as vertex arrays.

    Returns (ee, vf): float64 arrays of shape (n, 24) laid out like the first 192
    bytes of the reference's CCDData (cuda/narrow_phase/ccd_data.cuh:8-18):
    v0s v1s v2s v3s v0e v1e v2e v3e.  EE: parallel / near-parallel edges sliding past
    each other at offsets 10^U(-12,-6); VF: vertex trajectories grazing the face
    plane / face edges within 10^U(-12,-6).

    parallel=False keeps the edges skew (tilt 10^U(-3,-1)): with a minimum separation
    ms > 0, exactly parallel overlapping edges have a whole contact LINE that the solver
    must resolve at tolerance (>1e8 boxes per query, for the reference as well), which is
    only usable together with max_iter."""
    rng = np.random.default_rng(seed)
    off = 10.0 ** rng.uniform(-12, -6, n)
    sign = rng.choice([-1.0, 1.0], n)
    # ---- edge-edge: edge A along x at height +-off above edge B (near parallel)
    ee = np.zeros((n, 8, 3))
    tilt = 10.0 ** rng.uniform(-10, -3, n) * rng.choice([0.0, 1.0], n)
    if not parallel:
        tilt = 10.0 ** rng.uniform(-3, -1, n)
    shift = rng.uniform(-0.3, 0.3, (n, 3))
    a0 = np.stack([-0.5 + 0 * off, 0 * off, off * sign], 1)
    a1 = np.stack([0.5 + 0 * off, tilt, off * sign], 1)
    b0 = np.stack([-0.4 + 0 * off, -tilt, 0 * off], 1)
    b1 = np.stack([0.6 + 0 * off, 0 * off, 0 * off], 1)
    move = np.stack([rng.uniform(-0.2, 0.2, n), rng.uniform(-0.2, 0.2, n),
                     -sign * off * rng.uniform(0.0, 3.0, n)], 1)
    ee[:, 0], ee[:, 1], ee[:, 2], ee[:, 3] = a0 + shift, a1 + shift, b0 + shift, b1 + shift
    ee[:, 4], ee[:, 5] = a0 + shift + move, a1 + shift + move
    ee[:, 6], ee[:, 7] = b0 + shift, b1 + shift
    # ---- vertex-face: vertex slides along / grazes the plane of a static-ish face
    vf = np.zeros((n, 8, 3))
    f0 = np.array([0.0, 0.0, 0.0]) + shift
    f1 = np.array([1.0, 0.0, 0.0]) + shift
    f2 = np.array([0.0, 1.0, 0.0]) + shift
    uv = rng.uniform(-0.1, 0.6, (n, 2))          # some start outside, near an edge
    p0 = np.stack([uv[:, 0], uv[:, 1], off * sign], 1) + shift
    dz = -sign * off * rng.uniform(0.0, 3.0, n)   # 1/3 never reach the plane
    p1 = p0 + np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.3, 0.3, n), dz], 1)
    wob = 10.0 ** rng.uniform(-12, -7, (n, 3)) * rng.choice([0.0, 1.0], (n, 1))
    vf[:, 0], vf[:, 1], vf[:, 2], vf[:, 3] = p0, f0, f1, f2
    vf[:, 4], vf[:, 5], vf[:, 6], vf[:, 7] = p1, f0 + wob, f1 - wob, f2 + wob
    return (np.ascontiguousarray(ee.reshape(n, 24)),
            np.ascontiguousarray(vf.reshape(n, 24)))


SCENES = {
    "c1": scene_c1,
    "c2": scene_c2,
    "c3": scene_c3,
    "c4": scene_c4,
}


def num_boxes(scene) -> int:
    return scene["V0"].shape[0] + scene["E"].shape[0] + scene["F"].shape[0]
