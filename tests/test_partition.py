"""CPU tests of omegahpetsc_b200/partition.py and meshgen.py (no GPU, no process group): strip partitions of
any width, the overlap closure each rank assembles on, ghost resolution with an in-process all-gather."""
import numpy as np
import pytest

from omegahpetsc_b200 import meshgen, partition


def fake_all_gather(world_parts, vert_owner):
    """Emulates torch.distributed.all_gather_object for resolve_ghost_roots by precomputing every rank's table."""
    tables = []
    for r, part in enumerate(world_parts):
        mine = vert_owner[part["gverts"]] == r
        tables.append((part["gverts"][mine], np.nonzero(mine)[0]))

    def gather(out, _obj):
        for i, t in enumerate(tables):
            out[i] = t
    return gather


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("dim", [2, 3])
def test_chain_partition_and_overlap_closure(world, dim):
    cells, coords = meshgen.box2d(17, 12) if dim == 2 else meshgen.box3d(4, 3, 9)
    eo, vo = partition.chain_partition(cells, coords, world)
    counts = np.bincount(eo, minlength=world)
    assert counts.min() > 0 and counts.max() - counts.min() <= 1, "strips of equal element count"
    assert vo.min() == 0 and vo.max() == world - 1
    # vertex owner = lowest part among its elements (XGC picpart convention: ghosts point to lower ranks)
    lowest = np.full(coords.shape[0], world)
    np.minimum.at(lowest, cells.reshape(-1), np.repeat(eo, cells.shape[1]))
    assert np.array_equal(vo, lowest)
    parts = [partition.local_part(cells, coords, eo, vo, r) for r in range(world)]
    gather = fake_all_gather(parts, vo)
    owned_total = 0
    covered = np.zeros(cells.shape[0], int)
    for r, p in enumerate(parts):
        gv = p["gverts"]
        assert np.array_equal(gv, np.unique(gv)), "local vertices in ascending global order"
        touch = (vo[cells] == r).any(axis=1)
        assert np.array_equal(p["elem_global"], np.nonzero(touch)[0])
        covered[p["elem_global"]] += 1
        owned = vo[gv] == r
        owned_total += int(owned.sum())
        # every cell containing an owned vertex is local: owned stars are complete
        star_sizes_global = np.bincount(cells.reshape(-1), minlength=coords.shape[0])
        star_sizes_local = np.bincount(p["cells"].reshape(-1), minlength=gv.size)
        assert np.array_equal(star_sizes_local[owned], star_sizes_global[gv][owned])
        roots = partition.resolve_ghost_roots(p, vo, r, world, gather)
        for k in range(p["ghost_local"].size):
            q = p["ghost_owner"][k]
            assert q != r and parts[q]["gverts"][roots[k]] == gv[p["ghost_local"][k]]
    assert owned_total == coords.shape[0], "every vertex owned exactly once"
    assert covered.min() >= 1, "every element is assembled by at least one rank"
    if world > 1:
        assert covered.max() >= 2, "interface elements are shared (one overlap layer)"


def test_generators_closed_forms():
    cells, coords = meshgen.box2d(5, 3)
    assert cells.shape == (30, 3) and coords.shape == (24, 2) and cells.dtype == np.int32
    X = coords[cells]
    area = 0.5 * ((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 1, 1] - X[:, 0, 1]) * (X[:, 2, 0] - X[:, 0, 0]))
    assert (area > 0).all() and area.sum() == pytest.approx(1.0, rel=1e-14)
    cells, coords = meshgen.box3d(2, 3, 2)
    assert cells.shape == (72, 4) and coords.shape == (36, 3)
    X = coords[cells]
    vol = np.einsum("ij,ij->i", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0]) / 6.0
    assert (vol > 0).all() and vol.sum() == pytest.approx(1.0, rel=1e-14)


def test_config3_at_8_gpus_is_the_24_part_fixture_merged():
    """BASELINE config 3 has no 2- or 8-part fixtures: they are derived by merging consecutive parts of the reference's
    24-part XGC partition.  Host-only check of that derivation (tests/mgpu_worker.load)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import mgpu_worker
    for world, counts in ((8, [431, 1285, 1889, 1897, 2340, 2820, 3364, 9728]), (2, [5502, 18252])):
        cells, coords, eo, vo = mgpu_worker.load("24k_24p", world)
        assert cells.shape == (23754, 3) and coords.shape == (12060, 2)
        assert np.bincount(eo, minlength=world).tolist() == counts        # "owned elements 8p: 431 ... 9728"
        for r in range(world):                                            # chain: a part's ghosts live on lower parts only
            gv = np.unique(cells[eo == r])
            assert (vo[gv] <= r).all()
        nghost = [int((vo[np.unique(cells[eo == r])] != r).sum()) for r in range(world)]
        if world == 8:
            assert nghost == [0, 47, 96, 141, 176, 214, 256, 306]          # SURVEY 8d ghost counts
