"""CPU: the oracle's neighbour lists against the expectations written down in the reference's
(orphaned, but still the only written spec) test/test_spatialvb.cc:81-581."""
import numpy as np

import oracle


def coords_of(points):
    return np.ascontiguousarray(np.array(points, dtype=np.int32).T)


def test_1d_line():
    """test_spatialvb.cc CalcNeighboursOneVoxel / 1-D cases: interior voxels see both sides, ends one."""
    n = 5
    c = coords_of([(x, 0, 0) for x in range(n)])
    n1, n2 = oracle.neighbours(c, 1)
    assert n1[0] == [2] and n1[n - 1] == [n - 1]
    for v in range(1, n - 1):
        assert n1[v] == [v + 2, v]  # +x first, then -x (inference_vb.cc:863-869), 1-based ids
    assert n2[0] == [3] and sorted(n2[2]) == [1, 5]


def test_single_voxel_has_no_neighbours():
    n1, n2 = oracle.neighbours(coords_of([(3, 4, 5)]), 3)
    assert n1 == [[]] and n2 == [[]]


def test_2d_grid_no_wraparound():
    """x/y wrap-around must be rejected (inference_vb.cc:906-925): the last voxel of a row is not a
    neighbour of the first voxel of the next row."""
    nx, ny = 4, 3
    pts = [(x, y, 0) for y in range(ny) for x in range(nx)]
    n1, _ = oracle.neighbours(coords_of(pts), 2)
    vid = lambda x, y: y * nx + x + 1
    assert n1[vid(3, 0) - 1] == [vid(2, 0), vid(3, 1)]
    assert n1[vid(0, 1) - 1] == [vid(1, 1), vid(0, 2), vid(0, 0)]
    assert sorted(n1[vid(1, 1) - 1]) == sorted([vid(2, 1), vid(0, 1), vid(1, 2), vid(1, 0)])


def test_3d_cube_counts_and_second_neighbours():
    n = 3
    pts = [(x, y, z) for z in range(n) for y in range(n) for x in range(n)]
    n1, n2 = oracle.neighbours(coords_of(pts), 3)
    centre = 1 * 9 + 1 * 3 + 1
    assert len(n1[centre]) == 6 and len(n1[0]) == 3
    # second neighbours keep duplicates: the centre reaches each of its 12 edge-diagonal voxels twice and
    # has no straight second neighbours in a 3^3 cube -> 24 entries
    assert len(n2[centre]) == 24
    # spatial_dims = 2 restricts to in-plane neighbours
    n1d2, _ = oracle.neighbours(coords_of(pts), 2)
    assert len(n1d2[centre]) == 4


def test_irregular_mask():
    pts = [(0, 0, 0), (1, 0, 0), (3, 0, 0), (1, 1, 0), (1, 0, 1)]
    pts = sorted(pts, key=lambda p: (p[2], p[1], p[0]))
    n1, _ = oracle.neighbours(coords_of(pts), 3)
    ids = {p: i + 1 for i, p in enumerate(pts)}
    assert sorted(n1[ids[(1, 0, 0)] - 1]) == sorted([ids[(0, 0, 0)], ids[(1, 1, 0)], ids[(1, 0, 1)]])
    assert n1[ids[(3, 0, 0)] - 1] == []
