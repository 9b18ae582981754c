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

Known deviation of the mesh preparation: the reference's inflate() moves a vertex along the normalised sum of the
unit normals of the triangles its half-edge walk reaches (getNeighborTriangleIds, triangularSurfaceMesh.hh:495-518),
which on meshes with inconsistently oriented patches (cylinderSimple.stl: the caps) is only part of the one-ring;
this file and the product sum over the whole one-ring.  Vertices then differ by up to the inflate amount (1e-3
lattice units); the voxel matrices of all golden cases are nevertheless identical.
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


def mesh_to_lattice(soup, resolution, reference_direction, margin, extra_layer=0, inflate=1e-3):
    """float32 [n,3,3] as stored in the STL -> (float64 [n,3,3] in lattice units, dims)"""
    v = np.asarray(soup, dtype=np.float32).astype(np.float64).reshape(-1, 3)
    uniq, inv = np.unique(v, axis=0, return_inverse=True)
    idx = inv.reshape(-1, 3)
    lo, hi = uniq.min(axis=0), uniq.max(axis=0)
    scale = float(resolution) / (hi[reference_direction] - lo[reference_direction])
    off = float(margin + extra_layer)
    P = (uniq - lo) * scale + off
    if inflate:
        t = P[idx]
        n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
        n /= np.linalg.norm(n, axis=1)[:, None]
        vn = np.zeros_like(P)
        for k in range(3):
            np.add.at(vn, idx[:, k], n)
        vn /= np.linalg.norm(vn, axis=1)[:, None]
        P = P + inflate * vn
    dims = tuple(int(e) + 1 + 2 * (margin + extra_layer) for e in (P.max(axis=0) - P.min(axis=0)))
    return P[idx], dims
