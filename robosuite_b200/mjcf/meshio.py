"""Mesh file readers (binary/ascii STL, OBJ, legacy MuJoCo MSH) and mass-property / convex-hull helpers.

Host-side, compile-time only (not on the per-step path).  The reference hands mesh files to MuJoCo's compiler via
the composed MJCF (`robosuite/models/assets/**/meshes`); this module is the loader's replacement for that step.
"""
import struct
import numpy as np


def _read_stl(data: bytes) -> tuple:
    # binary STL: 80-byte header, uint32 ntri, then 50 bytes per triangle
    if len(data) >= 84:
        ntri = struct.unpack_from("<I", data, 80)[0]
        if 84 + 50 * ntri == len(data):
            rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                                count=ntri, offset=84)
            v = rec["v"].reshape(-1, 3).astype(np.float64)
            f = np.arange(3 * ntri, dtype=np.int64).reshape(-1, 3)
            return v, f
    # ascii STL
    verts = []
    for line in data.decode("utf-8", "ignore").splitlines():
        s = line.split()
        if len(s) == 4 and s[0] == "vertex":
            verts.append([float(s[1]), float(s[2]), float(s[3])])
    v = np.asarray(verts, dtype=np.float64)
    return v, np.arange(len(v), dtype=np.int64).reshape(-1, 3)


def _read_obj(data: bytes) -> tuple:
    verts, faces = [], []
    for line in data.decode("utf-8", "ignore").splitlines():
        s = line.split()
        if not s:
            continue
        if s[0] == "v":
            verts.append([float(s[1]), float(s[2]), float(s[3])])
        elif s[0] == "f":
            idx = [int(t.split("/")[0]) for t in s[1:]]
            idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
            for k in range(1, len(idx) - 1):  # fan triangulation
                faces.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.int64).reshape(-1, 3)


def _read_msh(data: bytes) -> tuple:
    # legacy MuJoCo .msh: int32 nvertex, nnormal, ntexcoord, nface; then float32 arrays; int32 faces
    nv, nn, nt, nf = struct.unpack_from("<4i", data, 0)
    off = 16
    v = np.frombuffer(data, dtype="<f4", count=3 * nv, offset=off).reshape(-1, 3).astype(np.float64)
    off += 12 * nv + 12 * nn + 8 * nt
    f = np.frombuffer(data, dtype="<i4", count=3 * nf, offset=off).reshape(-1, 3).astype(np.int64)
    return v, f


def load_mesh(path: str):
    with open(path, "rb") as fh:
        data = fh.read()
    ext = path.lower().rsplit(".", 1)[-1]
    if ext == "stl":
        v, f = _read_stl(data)
    elif ext == "obj":
        v, f = _read_obj(data)
    elif ext == "msh":
        v, f = _read_msh(data)
    else:
        raise ValueError(f"unsupported mesh format: {path}")
    # merge repeated vertices (MuJoCo does this for STL)
    uv, inv = np.unique(v, axis=0, return_inverse=True)
    # keep first-occurrence order to be deterministic
    first = np.full(len(uv), len(v), dtype=np.int64)
    np.minimum.at(first, inv, np.arange(len(v)))
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return uv[order], rank[inv][f]


def mesh_mass_properties(v, f):
    """Volume, centre of mass and unit-density inertia tensor about the COM of a closed triangle mesh
    (signed tetrahedra against the origin)."""
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))
    vol = vol6.sum() / 6.0
    if abs(vol) < 1e-18:
        return 0.0, v.mean(axis=0), np.zeros((3, 3))
    com = ((a + b + c) * vol6[:, None]).sum(axis=0) / (24.0 * vol)
    # second moments: integral of x x^T over each tetra (0,a,b,c) = vol6/120 * (sum_ij (1+delta_ij) p_i p_j^T)
    S = np.zeros((3, 3))
    for p, q in ((a, a), (b, b), (c, c)):
        S += np.einsum("i,ij,ik->jk", vol6, p, q) * 2.0
    for p, q in ((a, b), (a, c), (b, c)):
        m = np.einsum("i,ij,ik->jk", vol6, p, q)
        S += m + m.T
    S /= 120.0
    if vol < 0:
        vol, S = -vol, -S
    S -= vol * np.outer(com, com)
    inertia = np.trace(S) * np.eye(3) - S
    return vol, com, inertia


def convex_hull(v):
    """Hull vertices (subset of v, original order) and triangle faces indexed into that subset.
    MuJoCo itself calls qhull for this step; scipy.spatial.ConvexHull is the same library."""
    from scipy.spatial import ConvexHull

    h = ConvexHull(v)
    idx = np.sort(np.unique(h.simplices.ravel()))
    remap = -np.ones(len(v), dtype=np.int64)
    remap[idx] = np.arange(len(idx))
    faces = remap[h.simplices]
    hv = v[idx]
    # orient faces outward
    c = hv.mean(axis=0)
    a, b, cc = hv[faces[:, 0]], hv[faces[:, 1]], hv[faces[:, 2]]
    n = np.cross(b - a, cc - a)
    flip = np.einsum("ij,ij->i", n, a - c) < 0
    faces[flip] = faces[flip][:, ::-1]
    return hv, faces
