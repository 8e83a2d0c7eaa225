"""GPU: batched SpatialIndex::get_nearest_neighbours (location_hash_2d.rs:151-238), ring-search quirks
included, against the oracle and the reference's own test (location_hash_2d.rs:310-339)."""
import numpy as np
import pytest

import oracle_ffi as O
import rmf_crowdsim_b200 as R

pytestmark = pytest.mark.gpu


def _grid100(h):
    ids, pts = [], []
    for x in range(10):
        for y in range(10):
            ids.append(10 * x + y)
            pts.append((x + 0.5, y + 0.5))
    return np.array(ids, dtype=np.uint64), np.array(pts)


def test_nearest_neighbours_reference_test():
    g = R.LocationHash2D(10.0, 10.0, 0.5, (0.0, 0.0), capacity=128)
    ids, pts = _grid100(g)
    g.add_or_update_many(ids, pts)
    assert g.get_nearest_neighbours(1, (0.6, 0.6)) == [0]
    assert g.get_nearest_neighbours(4, (1.7, 1.6)) == [11, 21, 12, 10]


@pytest.mark.parametrize("cell,n_pts", [(0.5, 300), (1.0, 2000), (3.0, 500)])
def test_knn_matches_oracle_including_ring_quirks(cell, n_pts):
    rng = np.random.default_rng(int(cell * 10) + n_pts)
    w = h = 30.0
    off = (-3.0, -4.0)
    g = R.LocationHash2D(w, h, cell, off, capacity=4096)
    o = O.OracleSim(w, h, cell, off)
    # in-grid points, a few duplicates of position (distance ties -> stable order) and cell-boundary points
    pts = rng.uniform([off[0], off[1]], [off[0] + w - 1e-9, off[1] + h - 1e-9], size=(n_pts, 2))
    pts[10:20] = pts[0:10]
    pts[20:30] = np.floor(pts[20:30] / cell) * cell
    pts = pts[(pts[:, 0] >= off[0]) & (pts[:, 1] >= off[1])]
    ids = rng.permutation(len(pts)).astype(np.uint64)
    g.add_or_update_many(ids, pts)
    for i, p in zip(ids, pts):
        o.index_add_or_update(int(i), p)
    q = np.concatenate([rng.uniform([off[0] - 5, off[1] - 5], [off[0] + w + 5, off[1] + h + 5], size=(200, 2)),
                        pts[:20], np.array([[off[0], off[1]], [1e9, 1e9], [-1e9, 3.0]])])
    for k in (1, 3, 8, 40):
        got_ids, got_counts = g.query_knn(q, k)
        for r in range(len(q)):
            want = o.query_knn(k, q[r])
            assert int(got_counts[r]) == len(want), (k, r, q[r])
            assert list(got_ids[r, :len(want)]) == list(want), (k, r, q[r])


def test_knn_on_an_empty_index_and_k_zero():
    g = R.LocationHash2D(8.0, 8.0, 1.0, (0.0, 0.0), capacity=16)
    assert g.get_nearest_neighbours(3, (4.0, 4.0)) == []
    g.add_or_update(7, (1.5, 1.5))
    assert g.get_nearest_neighbours(0, (4.0, 4.0)) == []
    assert g.get_nearest_neighbours(1, (4.0, 4.0)) == [7]
    # The ring walk visits cell (x-s, y-s) twice (half-open side loops), and a row index >= n_y aliases into
    # the next column (signed_idx_to_data_idx has no y bound, location_hash_2d.rs:74-85): cell (0, 9) of the
    # fifth ring IS data cell 9 = (1, 1).  The reference therefore returns the same agent three times.
    o = O.OracleSim(8.0, 8.0, 1.0, (0.0, 0.0))
    o.index_add_or_update(7, (1.5, 1.5))
    assert list(o.query_knn(3, (4.0, 4.0))) == [7, 7, 7]
    assert g.get_nearest_neighbours(3, (4.0, 4.0)) == [7, 7, 7]
