"""GPU: the SpatialIndex surface (spatial_index/spatial_index.rs:4-14) under seeded random sequences of
add_or_update / remove_agent / get_neighbours_in_radius / get_nearest_neighbours, against the oracle's LocationHash2D.
Query results are compared as LISTS: cells x-major then y, ascending id inside a cell (radius); ring walk with its
half-open sides and the stable distance sort (nearest)."""
import numpy as np
import pytest

import oracle_ffi as O
import rmf_crowdsim_b200 as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,cell", [(0, 1.0), (1, 0.5), (2, 2.5), (3, 4.0)])
def test_random_index_updates_and_queries_match_the_oracle(seed, cell):
    rng = np.random.default_rng(900 + seed)
    w = h = 40.0
    off = (float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5)))
    g = R.LocationHash2D(w, h, cell, off, capacity=2048)
    o = O.OracleSim(w, h, cell, off)
    live = {}

    def rand_points(k):
        # left of / below the origin is legal (the insert cell saturates to 0, location_hash_2d.rs:56-57)
        return rng.uniform([off[0] - 2.0, off[1] - 2.0], [off[0] + w - 1e-6, off[1] + h - 1e-6], size=(k, 2))

    next_id = 0
    for it in range(60):
        op = rng.choice(["add", "move", "remove", "query", "query"])
        if op == "add" or not live:
            k = int(rng.integers(1, 60))
            pts = rand_points(k)
            m = k // 8  # some exactly on cell edges
            pts[:m] = np.floor((pts[:m] - off) / cell) * cell + off
            ids = np.arange(next_id, next_id + k, dtype=np.uint64)
            next_id += k
            g.add_or_update_many(ids, pts)
            for i, p in zip(ids, pts):
                o.index_add_or_update(int(i), p)
                live[int(i)] = p
        elif op == "move":
            ids = rng.choice(list(live), size=min(len(live), int(rng.integers(1, 30))), replace=False).astype(np.uint64)
            pts = rand_points(len(ids))
            g.add_or_update_many(ids, pts)
            for i, p in zip(ids, pts):
                o.index_add_or_update(int(i), p)
                live[int(i)] = p
        elif op == "remove":
            for i in rng.choice(list(live), size=min(len(live), int(rng.integers(1, 10))), replace=False):
                g.remove_agent(int(i))
                o.index_remove(int(i))
                del live[int(i)]
        else:
            q = rng.uniform([off[0] - 4.0, off[1] - 4.0], [off[0] + w + 4.0, off[1] + h + 4.0], size=(24, 2))
            for r, i in enumerate(list(live)[:6]):  # some queries exactly on an indexed point
                q[r] = live[i]
            radius = rng.choice([0.3, 1.0, 2.2, 5.5, 12.0], size=len(q))
            offsets, ids = g.query_radius(q, radius)
            for r in range(len(q)):
                want = o.query_radius(radius[r], q[r])
                got = ids[int(offsets[r]):int(offsets[r + 1])]
                assert list(got) == list(want), (it, r, q[r], radius[r])
            for k in (1, 5, 17):
                got_ids, got_counts = g.query_knn(q, k)
                for r in range(len(q)):
                    want = o.query_knn(k, q[r])
                    assert int(got_counts[r]) == len(want), (it, k, r)
                    assert list(got_ids[r, :len(want)]) == list(want), (it, k, r)


def test_out_of_bounds_update_is_refused_and_changes_nothing():
    g = R.LocationHash2D(10.0, 10.0, 1.0, (0.0, 0.0), capacity=16)
    g.add_or_update(3, (2.5, 2.5))
    with pytest.raises(R.CrowdsimError) as e:
        g.add_or_update(3, (20.0, 2.5))
    assert "Index out of bounds" in str(e.value)
    assert g.get_neighbours_in_radius(1.0, (2.5, 2.5)) == [3]
