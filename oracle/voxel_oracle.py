"""oracle/voxel_oracle.py -- TEST INFRASTRUCTURE ONLY: CPU (numpy) statement of the STL -> voxel-flag step.

What it stands for in the reference (relative to /root/reference/src): voxelize(mesh, domain, borderWidth)
offLattice/voxelizer.hh:107-142 with VoxelizeMeshFunctional3D (:527-610) and detectBorderLine (:640-690), and
the mesh preparation DEFscaledMesh + TriangularSurfaceMesh::inflate (offLattice/triangleBoundary3D.hh:255-301,
triangularSurfaceMesh.hh:1419-1425, :1552-1577).

The reference floods the lattice sequentially from an outside seed, deciding each voxel from an already
decided neighbour by the parity of the surface crossings of the connecting segment.  This oracle states the
same classification as the parity of the crossings of the lattice line x = ix, y = iy below the voxel, with a
symbolic perturbation for lines through mesh edges and vertices.  PINNED: tests/test_voxel_oracle.py checks it
bit for bit against tests/golden/voxel_*.npz, voxel matrices dumped by the UNMODIFIED reference
(oracle/_ref/ref_voxel, oracle/gen_voxel_golden.py) for two STL files of the reference at five resolutions /
orientations / margins / border widths.

The mesh preparation follows the reference operation by operation, including two quirks of this fork that decide
where inflate() moves a vertex: TriangleToDef::fixOrientation never changes a triangle (offLattice/triangleToDef.hh:240-330)
and getNeighborTriangleIds (triangularSurfaceMesh.hh:495-518) walks half-edges that, on inconsistently wound patches
(the caps of cylinderSimple.stl), reach only part of a vertex's one-ring.  Pinned on the triangles the reference hands
to voxelize() (fixture key "tri"): agreement to 1e-10 lattice units.
"""
import numpy as np

OUTSIDE, OUTER_BORDER, INSIDE, INNER_BORDER = 1, 2, 3, 4      # offLattice/voxelizer.h:38-48


def _edge(ux, uy, vx, vy, px, py):
    """edge function of u -> v at p with the end points in canonical order (identical value, opposite sign,
    for the two triangles sharing the edge); u, v scalars, p arrays"""
    swap = (ux > vx) or (ux == vx and uy > vy)
    if swap:
        ux, uy, vx, vy = vx, vy, ux, uy
    e = (vx - ux) * (py - uy) - (vy - uy) * (px - ux)
    return -e if swap else e


def voxelize(tri, dims, border_width=1):
    """tri float64 [n,3,3] in lattice units -> int8 [nx,ny,nz] voxelFlag values"""
    nx, ny, nz = [int(d) for d in dims]
    marks = np.zeros((nx, ny, nz + 1), dtype=np.uint8)
    for t in np.asarray(tri, dtype=np.float64):
        (ax, ay, az), (bx, by, bz), (cx, cy, cz) = [[float(c) for c in v] for v in t]
        area = float(_edge(ax, ay, bx, by, np.float64(cx), np.float64(cy)))
        if area == 0.0:
            continue
        x0, x1 = max(0, int(np.ceil(min(ax, bx, cx)))), min(nx - 1, int(np.floor(max(ax, bx, cx))))
        y0, y1 = max(0, int(np.ceil(min(ay, by, cy)))), min(ny - 1, int(np.floor(max(ay, by, cy))))
        if x0 > x1 or y0 > y1:
            continue
        ix, iy = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1), indexing="ij")
        px, py = ix.astype(np.float64), iy.astype(np.float64)
        sgn = -1.0 if area < 0 else 1.0
        ok = np.ones(px.shape, dtype=bool)
        es = []
        for (ux, uy, vx, vy) in ((bx, by, cx, cy), (cx, cy, ax, ay), (ax, ay, bx, by)):
            e = sgn * _edge(ux, uy, vx, vy, px, py)
            dx, dy = sgn * (vx - ux), sgn * (vy - uy)
            tie = (dy < 0) or (dy == 0 and dx > 0)
            ok &= (e > 0) | ((e == 0) & tie)
            es.append(e)
        if not ok.any():
            continue
        with np.errstate(invalid="ignore", divide="ignore"):
            zc = (es[0] * az + es[1] * bz + es[2] * cz) / (es[0] + es[1] + es[2])
        iz0 = np.floor(zc[ok]) + 1.0
        iz0 = np.where(iz0 < 0, 0, iz0)
        keep = iz0 <= nz
        np.bitwise_xor.at(marks, (ix[ok][keep], iy[ok][keep], iz0[keep].astype(np.int64)), 1)
    inside = np.bitwise_xor.accumulate(marks, axis=2)[:, :, :nz].astype(bool)
    return flags_from_inside(inside, border_width)


def flags_from_inside(inside, bw):
    """detectBorderLine offLattice/voxelizer.hh:640-690 (the matrix has a one-cell envelope of `outside`;
    neighbours further out are not looked at)"""
    nx, ny, nz = inside.shape
    pad = np.pad(inside, bw, constant_values=False)
    known = np.zeros(pad.shape, dtype=bool)
    lo = max(bw - 1, 0)
    known[lo:2 * bw + nx - lo, lo:2 * bw + ny - lo, lo:2 * bw + nz - lo] = True
    nb_in = np.zeros_like(inside)
    nb_out = np.zeros_like(inside)
    for dx in range(-bw, bw + 1):
        for dy in range(-bw, bw + 1):
            for dz in range(-bw, bw + 1):
                if dx == dy == dz == 0:
                    continue
                sl = (slice(bw + dx, bw + dx + nx), slice(bw + dy, bw + dy + ny), slice(bw + dz, bw + dz + nz))
                nb_in |= pad[sl] & known[sl]
                nb_out |= (~pad[sl]) & known[sl]
    flags = np.where(inside, INSIDE, OUTSIDE).astype(np.int8)
    flags[(~inside) & nb_in] = OUTER_BORDER
    flags[inside & nb_out] = INNER_BORDER
    return flags


MESH_EPS = 10.0 * float(np.finfo(np.float32).eps)       # getEpsilon<T>(FLT) core/globalDefs.h:104-127


def _is_zero(x):
    """util::fpequal(x, 0, eps) core/util.h:428-457"""
    return abs(x) <= 2.0 * MESH_EPS


def fix_facets(tris, normals):
    """TriangleSet::checkForDegenerateTrianglesAndFixOrientationNoAbort offLattice/triangleSet.hh:497-548: drop facets
    with a vanishing edge or area, swap vertices 1, 2 where the facet's own normal disagrees with the stated one"""
    out = []
    for t, n in zip(np.asarray(tris, dtype=np.float64), np.asarray(normals, dtype=np.float64)):
        e01, e02, e12 = t[1] - t[0], t[2] - t[0], t[2] - t[1]
        cn = np.cross(e01, e02)
        nrm = lambda v: float(np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))
        if _is_zero(nrm(e01)) or _is_zero(nrm(e02)) or _is_zero(nrm(e12)) or _is_zero(nrm(cn)):
            continue
        if cn[0] * n[0] + cn[1] * n[1] + cn[2] * n[2] < 0.0:
            t = t[[0, 2, 1]]
        out.append(t)
    return np.array(out, dtype=np.float64).reshape(-1, 3, 3)


def read_stl(path):
    """TriangleSet<T>(fileName, FLT) offLattice/triangleSet.hh:130-350 -> float64 [n,3,3]"""
    import struct
    raw = open(path, "rb").read()
    if b"endfacet" in raw[:4096]:
        normals, verts = [], []
        for ln in raw.decode("ascii", "replace").splitlines():
            w = ln.split()
            if len(w) >= 5 and w[0] == "facet" and w[1] == "normal":
                normals.append([float(x) for x in w[2:5]])
            elif len(w) >= 4 and w[0] == "vertex":
                verts.append([float(x) for x in w[1:4]])
        return fix_facets(np.array(verts, dtype=np.float64).reshape(-1, 3, 3), np.array(normals, dtype=np.float64))
    nt = struct.unpack("<I", raw[80:84])[0]
    rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=nt, offset=84)
    return fix_facets(rec["v"].astype(np.float64), rec["n"].astype(np.float64))


def _lt(x, y):
    return (x < y) and (0.5 * (abs(x - y) + abs(y - x)) > MESH_EPS)


def _vertex_less(a, b):
    """TriangleToDef::VsLessThan offLattice/triangleToDef.hh:65-99"""
    eq = lambda x, y: (not _lt(x, y)) and (not _lt(y, x))
    return _lt(a[0], b[0]) or (eq(a[0], b[0]) and _lt(a[1], b[1])) or (eq(a[0], b[0]) and eq(a[1], b[1]) and _lt(a[2], b[2]))


class _Key:
    def __init__(self, v):
        self.v = v

    def __lt__(self, other):
        return _vertex_less(self.v, other.v)


def surface_mesh(soup):
    """constructSurfaceMesh / TriangleToDef offLattice/triangleToDef.hh:200-235 for closed meshes: unique vertices in
    order of first appearance, vertex indices per triangle, neighbouring half-edge `ne` (half-edge 3t+k runs from local
    vertex k to k+1), emanating half-edge per vertex"""
    import bisect
    tris = np.asarray(soup, dtype=np.float64).reshape(-1, 3, 3)
    nt = len(tris)
    keys, ids, verts = [], [], []
    idx = np.zeros((nt, 3), dtype=np.int64)
    for t in range(nt):
        for k in range(3):
            v = tuple(float(c) for c in tris[t, k])
            key = _Key(v)
            p = bisect.bisect_left(keys, key)
            if p < len(keys) and not _vertex_less(v, keys[p].v) and not _vertex_less(keys[p].v, v):
                idx[t, k] = ids[p]
            else:
                keys.insert(p, key)
                ids.insert(p, len(verts))
                idx[t, k] = len(verts)
                verts.append(v)
    table = {}
    for t in range(nt):
        for le in range(3):
            i0, i1 = int(idx[t, le]), int(idx[t, (le + 1) % 3])
            key = (min(i0, i1), max(i0, i1))
            if key not in table:
                table[key] = [t, -1]
            elif table[key][1] == -1:
                table[key][1] = t
            else:
                raise ValueError("the surface mesh contains non-manifold edges")
    ne = -np.ones(3 * nt, dtype=np.int64)

    def local_edge(t, a, b):
        for le in range(3):
            if {int(idx[t, le]), int(idx[t, (le + 1) % 3])} == {a, b}:
                return le
        raise ValueError("problem with the surface mesh connectivity")
    for (a, b), (t1, t2) in table.items():
        if t2 != -1:
            l1, l2 = local_edge(t1, a, b), local_edge(t2, a, b)
            ne[3 * t1 + l1] = 3 * t2 + l2
            ne[3 * t2 + l2] = 3 * t1 + l1
    em = -np.ones(len(verts), dtype=np.int64)
    for t in range(nt):
        for le in range(3):
            if em[idx[t, le]] == -1:
                em[idx[t, le]] = 3 * t + le
    return np.array(verts, dtype=np.float64), idx, ne, em


def _neighbor_triangles(v, ne, em):
    """TriangularSurfaceMesh::getNeighborTriangleIds triangularSurfaceMesh.hh:495-518"""
    prev = lambda e: e + 2 if e % 3 == 0 else e - 1
    ee = int(em[v])
    out = [ee // 3]
    e = int(ne[prev(ee)])
    if e < 0:
        e = ee
    while e != ee:
        out.append(e // 3)
        e = int(ne[prev(e)])
        if e < 0:
            e = ee
        if len(out) > len(ne):
            raise ValueError("half-edge walk does not close")
    return out


def mesh_to_lattice(soup, resolution, reference_direction, margin, extra_layer=0, inflate=1e-3):
    """DEFscaledMesh(set, resolution, referenceDirection, margin, extraLayer) (offLattice/triangleBoundary3D.hh:255-301,
    toLatticeUnits triangularSurfaceMesh.hh:1552-1577) + inflate (:1419-1425, vertex by vertex in place):
    float64 [n,3,3] of the STL -> (float64 [n,3,3] in lattice units, dims of the voxel matrix)"""
    verts, idx, ne, em = surface_mesh(soup)
    lo, hi = verts.min(axis=0), verts.max(axis=0)
    scale = float(resolution) / (hi[reference_direction] - lo[reference_direction])
    off = float(margin + extra_layer)
    P = ((verts + (-lo)) * scale) + off
    if inflate:
        for v in range(len(P)):
            n = np.zeros(3)
            for t in _neighbor_triangles(v, ne, em):
                a, b, c = P[idx[t, 0]], P[idx[t, 1]], P[idx[t, 2]]
                e1, e2 = b - a, c - a
                m = np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]])
                n = n + m / np.sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2])
            P[v] = P[v] + inflate * (n / np.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]))
    dims = tuple(int(e) + 1 + 2 * (margin + extra_layer) for e in (P.max(axis=0) - P.min(axis=0)))
    return P[idx], dims
