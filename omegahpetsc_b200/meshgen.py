"""Synthetic inputs of BASELINE.json configs 4 and 5: the unit square split into
nx*ny quads, each cut on the same diagonal, vertices row-major; the unit cube split into
nx*ny*nz hexes x 6 Kuhn tetrahedra.  These stand in for Omega_h::build_box (ex12.cpp:805), whose
Hilbert reordering is Omega_h-internal; parity is "same inputs -> same outputs", so the generator
only has to be deterministic.  The strip / slab partitioner for the multi-GPU runs is in partition.py."""
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


def _part1by1(x):
    x = x.astype(np.uint64) & np.uint64(0xffff)
    x = (x | (x << np.uint64(8))) & np.uint64(0x00ff00ff)
    x = (x | (x << np.uint64(4))) & np.uint64(0x0f0f0f0f)
    x = (x | (x << np.uint64(2))) & np.uint64(0x33333333)
    x = (x | (x << np.uint64(1))) & np.uint64(0x55555555)
    return x


def hilbert_index(i, j, bits):
    """Distance of the grid points (i, j), 0 <= i, j < 2**bits, along the Hilbert curve of that order (the classic
    quadrant-rotation recurrence, vectorised over the points)."""
    x, y = i.astype(np.int64).copy(), j.astype(np.int64).copy()
    d = np.zeros(x.shape, dtype=np.int64)
    s = np.int64(1) << np.int64(bits - 1)
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate the quadrant so that the curve enters and leaves it at the right corners
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        x &= s - 1
        y &= s - 1
        s >>= np.int64(1)
    return d


def hilbert_reorder(cells, coords, sizes, pinned=False):
    """The same renumbering along a HILBERT curve, the curve Omega_h::build_box itself sorts by (ex12.cpp:805,
    Omega_h-internal): consecutive vertices are always grid neighbours, which a Morton curve does not guarantee."""
    nx, ny = sizes
    i = np.rint(coords[:, 0] * nx).astype(np.int64)
    j = np.rint(coords[:, 1] * ny).astype(np.int64)
    bits = max(1, int(max(nx, ny)).bit_length())
    return _reorder_by_key(cells, coords, hilbert_index(i, j, bits), pinned)


def morton_reorder(cells, coords, sizes, pinned=False):
    """Renumber the vertices (and sort the cells) of a 2-D box along a Morton (Z-order) curve.  Stands in for the
    space-filling-curve reordering of Omega_h::build_box (ex12.cpp:805): the mesh is the same, the numbering
    is not row-major any more, so |column - row| of the matrix is no longer bounded by the grid width."""
    nx, ny = sizes
    i = np.rint(coords[:, 0] * nx).astype(np.int64)
    j = np.rint(coords[:, 1] * ny).astype(np.int64)
    key = _part1by1(i) | (_part1by1(j) << np.uint64(1))
    return _reorder_by_key(cells, coords, key, pinned)


def _reorder_by_key(cells, coords, key, pinned):
    order = np.argsort(key, kind="stable")              # new vertex k = old vertex order[k]
    new_of_old = np.empty(order.size, dtype=np.int32)
    new_of_old[order] = np.arange(order.size, dtype=np.int32)
    c2 = new_of_old[cells]
    ckey = c2.min(axis=1)
    corder = np.argsort(ckey, kind="stable")
    out_cells = _alloc(cells.shape, np.int32, pinned)
    out_coords = _alloc(coords.shape, np.float64, pinned)
    out_cells[:] = c2[corder]
    out_coords[:] = coords[order]
    return out_cells, out_coords


def _alloc(shape, dtype, pinned):
    if not pinned:
        return np.empty(shape, dtype=dtype)
    import torch
    t = torch.empty(shape, dtype={np.float64: torch.float64, np.int32: torch.int32}[dtype], pin_memory=True)
    return t.numpy()
