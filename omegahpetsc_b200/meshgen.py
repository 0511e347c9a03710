"""Synthetic inputs of BASELINE.json configs 4 and 5: the unit square split into
nx*ny quads, each cut on the same diagonal, vertices row-major; the unit cube split into
nx*ny*nz hexes x 6 Kuhn tetrahedra.  These stand in for Omega_h::build_box (ex12.cpp:805), whose
Hilbert reordering is Omega_h-internal; parity is "same inputs -> same outputs", so the generator
only has to be deterministic.  Also the strip / slab partitioner used for the multi-GPU runs
(chain topology, like the XGC picparts)."""
import numpy as np


def box2d(nx, ny, lx=1.0, ly=1.0, pinned=False):
    vx = np.arange(nx + 1, dtype=np.float64) * (lx / nx)
    vy = np.arange(ny + 1, dtype=np.float64) * (ly / ny)
    coords = _alloc(((nx + 1) * (ny + 1), 2), np.float64, pinned)
    coords[:, 0] = np.tile(vx, ny + 1)
    coords[:, 1] = np.repeat(vy, nx + 1)
    cells = _alloc((2 * nx * ny, 3), np.int32, pinned)
    v00 = (np.arange(ny, dtype=np.int32)[:, None] * (nx + 1) + np.arange(nx, dtype=np.int32)[None, :]).reshape(-1)
    cells[0::2, 0] = v00
    cells[0::2, 1] = v00 + 1
    cells[0::2, 2] = v00 + nx + 2
    cells[1::2, 0] = v00
    cells[1::2, 1] = v00 + nx + 2
    cells[1::2, 2] = v00 + nx + 1
    return cells, coords


_KUHN = ((0, 1, 3, 7), (0, 3, 2, 7), (0, 2, 6, 7), (0, 6, 4, 7), (0, 4, 5, 7), (0, 5, 1, 7))


def box3d(nx, ny, nz, pinned=False):
    nvx, nvy, nvz = nx + 1, ny + 1, nz + 1
    coords = _alloc((nvx * nvy * nvz, 3), np.float64, pinned)
    coords[:, 0] = np.tile(np.arange(nvx) / nx, nvy * nvz)
    coords[:, 1] = np.tile(np.repeat(np.arange(nvy) / ny, nvx), nvz)
    coords[:, 2] = np.repeat(np.arange(nvz) / nz, nvx * nvy)
    k, j, i = np.meshgrid(np.arange(nz, dtype=np.int32), np.arange(ny, dtype=np.int32),
                          np.arange(nx, dtype=np.int32), indexing="ij")
    base = ((k * nvy + j) * nvx + i).reshape(-1)
    sx, sy, sz = 1, nvx, nvx * nvy
    corner = [base + (c & 1) * sx + ((c >> 1) & 1) * sy + ((c >> 2) & 1) * sz for c in range(8)]
    cells = _alloc((base.size * 6, 4), np.int32, pinned)
    for t, tet in enumerate(_KUHN):
        for a, c in enumerate(tet):
            cells[t::6, a] = corner[c]
    return cells, coords


def _alloc(shape, dtype, pinned):
    if not pinned:
        return np.empty(shape, dtype=dtype)
    import torch
    t = torch.empty(shape, dtype={np.float64: torch.float64, np.int32: torch.int32}[dtype], pin_memory=True)
    return t.numpy()


def chain_partition(cells, coords, nparts, axis=None):
    """Element partition into `nparts` strips/slabs of (nearly) equal element count along one
    coordinate axis, vertex owner = lowest part among the elements touching it (so ghosts of part
    r live on parts < r, the XGC picpart convention)."""
    dim = coords.shape[1]
    if axis is None:
        axis = dim - 1
    cent = coords[cells][:, :, axis].mean(axis=1)
    order = np.argsort(cent, kind="stable")
    elem_owner = np.empty(cells.shape[0], dtype=np.int32)
    bounds = np.linspace(0, cells.shape[0], nparts + 1).astype(np.int64)
    for r in range(nparts):
        elem_owner[order[bounds[r]:bounds[r + 1]]] = r
    vert_owner = np.full(coords.shape[0], nparts, dtype=np.int32)
    np.minimum.at(vert_owner, cells.reshape(-1), np.repeat(elem_owner, cells.shape[1]))
    return elem_owner, vert_owner


def local_part(cells, coords, elem_owner, vert_owner, rank):
    """Rank-local mesh for the owner-computes assembly: every element touching a vertex owned by
    `rank` (owned elements + one layer of overlap), its vertices in ascending global order, and
    the ghost list (local vertex, owner rank, GLOBAL vertex id of the root)."""
    touch = (vert_owner[cells] == rank).any(axis=1)
    lc = cells[touch]
    gverts = np.unique(lc.reshape(-1))
    g2l = np.full(coords.shape[0], -1, dtype=np.int64)
    g2l[gverts] = np.arange(gverts.size)
    lcells = g2l[lc].astype(np.int32)
    lcoords = np.ascontiguousarray(coords[gverts])
    ghost = np.nonzero(vert_owner[gverts] != rank)[0].astype(np.int32)
    return dict(cells=lcells, coords=lcoords, gverts=gverts, ghost_local=ghost,
                ghost_owner=vert_owner[gverts[ghost]].astype(np.int32), ghost_gvert=gverts[ghost],
                elem_global=np.nonzero(touch)[0])
