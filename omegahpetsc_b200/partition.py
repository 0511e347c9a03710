"""Host-side partition plumbing for the multi-GPU runs (one process per GPU).

The reference gets its partitions from Omega_h (`balance()`, ex12.cpp:819) or from pumipic picpart
files (ex12.cpp:810-813) and describes ghosts to PETSc as a point SF (ex12.cpp:864-876,917-939).
Here the synthetic boxes are cut into strips / slabs (a chain of parts, the topology of the XGC
picparts) and each rank's local mesh is the set of elements touching a vertex it
owns: owned elements plus one layer of overlap, so that every owned matrix row can be assembled
without communication ("owner computes")."""
import numpy as np


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


def ownership_from_picparts(elem_owner_tags, vert_owner_tags):
    """pumipic picpart files of one mesh agree on the `ownership` tags: take rank 0's."""
    return np.asarray(elem_owner_tags, dtype=np.int32), np.asarray(vert_owner_tags, dtype=np.int32)


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


def resolve_ghost_roots(part, vert_owner, rank, world, all_gather_object):
    """The `iremote` array of PetscSFSetGraph (ex12.cpp:868-871): for every ghost, the LOCAL vertex
    index it has on its owner.  Each rank publishes (global id, local id) of the vertices it owns;
    `all_gather_object(list_out, obj)` is torch.distributed's (any backend)."""
    mine = vert_owner[part["gverts"]] == rank
    tables = [None] * world
    all_gather_object(tables, (part["gverts"][mine], np.nonzero(mine)[0]))
    owner_vertex = np.empty(part["ghost_local"].size, dtype=np.int32)
    for r in range(world):
        sel = part["ghost_owner"] == r
        if sel.any():
            g, l = tables[r]
            pos = np.searchsorted(g, part["ghost_gvert"][sel])
            if not np.array_equal(g[pos], part["ghost_gvert"][sel]):
                raise ValueError("rank %d does not hold vertices that rank %d ghosts from it" % (r, rank))
            owner_vertex[sel] = l[pos]
    return owner_vertex
