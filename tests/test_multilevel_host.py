"""Host side of the multilevel preconditioner: the aggregation hierarchy (no device needed)."""
import numpy as np

import sim3opt_b200 as s3
from sim3opt_b200 import synth


def test_hierarchy_sizes_and_aggregates():
    g = synth.sphere(10, 1000, seed=42)
    nv = len(g["est"])
    sizes, blocks, agg = s3.host_multilevel(nv, g["fixed"], g["v0"], g["v1"])
    nf = int((np.asarray(g["fixed"]) == 0).sum())
    assert len(agg) == nf and len(sizes) >= 2
    assert np.all(np.diff(np.concatenate([[nf], sizes])) < 0)        # strictly coarser every level
    assert sizes[-1] <= 16                                          # coarsest level fits the dense inverse
    assert agg.min() == 0 and agg.max() == sizes[0] - 1
    assert np.all(np.bincount(agg, minlength=sizes[0]) >= 1)        # every aggregate is non-empty
    assert np.all(blocks >= sizes)                                  # every level keeps its diagonal blocks
    # aggregates are connected neighbourhoods: every member is the seed or adjacent to a member
    colptr, rowidx, hidx = s3.host_structure(nv, g["fixed"], g["v0"], g["v1"])
    cols = np.repeat(np.arange(nf), np.diff(colptr))
    off = rowidx != cols
    same = agg[rowidx[off]] == agg[cols[off]]
    touched = np.zeros(nf, bool)
    touched[rowidx[off][same]] = True
    touched[cols[off][same]] = True
    sizes0 = np.bincount(agg)
    assert np.all(touched[sizes0[agg] > 1])
    # deterministic
    sizes2, blocks2, agg2 = s3.host_multilevel(nv, g["fixed"], g["v0"], g["v1"])
    assert np.array_equal(sizes, sizes2) and np.array_equal(blocks, blocks2) and np.array_equal(agg, agg2)


def test_tiny_graph_has_no_levels(kitti_k1):
    g = synth.sphere(2, 6, seed=1)      # 12 poses: nothing to coarsen
    sizes, blocks, agg = s3.host_multilevel(len(g["est"]), g["fixed"], g["v0"], g["v1"])
    assert len(sizes) == 0
    sizes, blocks, agg = s3.host_multilevel(len(kitti_k1["est"]), kitti_k1["fixed"], kitti_k1["v0"], kitti_k1["v1"])
    assert len(sizes) >= 2 and sizes[0] < 771 // 2


def test_partitioned_hierarchy_respects_vertex_ranges():
    """World-size-N hierarchy (the partitioned solve): aggregates stay inside a rank's vertex range and
    are numbered rank by rank, so every rank owns a contiguous range of coarse rows."""
    g = synth.sphere(8, 500, seed=3)
    nv = len(g["est"])
    nf = int((np.asarray(g["fixed"]) == 0).sum())
    for world in (2, 4, 8):
        sizes, blocks, agg = s3.host_multilevel(nv, g["fixed"], g["v0"], g["v1"], world=world)
        seg = -(-nf // world)
        owner = np.arange(nf) // seg
        # one owner per aggregate
        lo = np.full(sizes[0], world, np.int64); hi = np.full(sizes[0], -1, np.int64)
        np.minimum.at(lo, agg, owner); np.maximum.at(hi, agg, owner)
        assert np.array_equal(lo, hi)
        assert np.all(np.diff(lo) >= 0)                  # coarse numbering follows the ranks
        assert np.all(np.diff(np.concatenate([[nf], sizes])) < 0)


def test_partitioned_galerkin_contributions_are_local():
    """The invariant the partitioned multilevel PCG relies on (csrc/amg.cu, localize_fine_level): every fine block
    (gi < gj) that contributes to an upper level-1 block (min(I,J), max(I,J)) is stored on the rank that owns that
    coarse row -- i.e. that rank owns the block's row vertex gi (cut edges live in the row of their lower end),
    and when the coarse block is taken transposed (I > J) both ends are its own."""
    for gname, g in (("sphere", synth.sphere(8, 500, seed=3)), ("manhattan", synth.manhattan3d(3000, seed=4))):
        nv = len(g["est"])
        nf = int((np.asarray(g["fixed"]) == 0).sum())
        colptr, rowidx, hidx = s3.host_structure(nv, g["fixed"], g["v0"], g["v1"])
        cols = np.repeat(np.arange(nf), np.diff(colptr))
        off = rowidx != cols
        gi, gj = rowidx[off], cols[off]                       # upper blocks: row < column
        assert np.all(gi < gj)
        for world in (2, 3, 8):
            sizes, blocks, agg = s3.host_multilevel(nv, g["fixed"], g["v0"], g["v1"], world=world)
            seg = -(-nf // world)
            owner_v = np.arange(nf) // seg
            owner_c = np.zeros(sizes[0], np.int64)
            owner_c[agg] = owner_v                            # aggregates are rank-confined (tested above)
            I, J = agg[gi], agg[gj]
            row_owner = owner_c[np.minimum(I, J)]
            assert np.all(row_owner == owner_v[gi]), gname    # the block sits in a row this rank owns
            flipped = I > J
            assert np.all(owner_v[gj[flipped]] == owner_v[gi[flipped]]), gname
