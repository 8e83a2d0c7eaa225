"""CPU: pins the oracle against the reference's own eight tests (constants restated from
/root/reference: lib.rs:422-453, location_hash_2d.rs:310-397, zanlungo.rs:224-236,
tests/event_listeners_test.rs:64-111) through the oracle's C API, and runs its C++ selftest."""
import subprocess

import numpy as np
import pytest

import oracle_ffi as O


def test_selftest_binary():
    O.build()
    r = subprocess.run([O.SELFTEST], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS" in r.stdout


def _grid100(o):
    pts = {}
    i = 0
    for x in range(10):
        for y in range(10):
            p = (x + 0.5, y + 0.5)
            o.index_add_or_update(i, p)
            pts[i] = p
            i += 1
    return pts


def test_nearest_neighbours():
    o = O.OracleSim(10, 10, 0.5, (0, 0))
    pts = _grid100(o)
    assert list(o.query_knn(1, (0.6, 0.6))) == [0]
    got = list(o.query_knn(4, (1.7, 1.6)))
    d = sorted((np.hypot(p[0] - 1.7, p[1] - 1.6), i) for i, p in pts.items())
    assert got == [i for _, i in d[:4]] == [11, 21, 12, 10]


def test_radius_search():
    o = O.OracleSim(10, 10, 0.5, (0, 0))
    pts = _grid100(o)
    gt = {i for i, p in pts.items() if np.sqrt((p[0] - 4) ** 2 + (p[1] - 4) ** 2) < 1.1}
    assert set(o.query_radius(1.1, (4, 4))) == gt == {33, 34, 43, 44}


def test_update():
    o = O.OracleSim(2, 2, 1, (0, 0))
    o.index_add_or_update(1, (0, 0))
    assert list(o.query_radius(1.0, (0, 0))) == [1]
    o.index_add_or_update(1, (1, 0))
    assert list(o.query_radius(1.0, (0, 0))) == []  # distance 1 is not < 1


def test_remove():
    o = O.OracleSim(1, 1, 1, (0, 0))
    o.index_add_or_update(1, (0, 0))
    assert len(o.query_radius(1.1, (0, 0))) == 1
    o.index_remove(1)
    assert len(o.query_radius(1.1, (0, 0))) == 0


def test_time_to_collision():
    assert O.ttc(4.0, (1, 0), (-10, 0)) == 6.0
    assert O.ttc(4.0, (1, 0), (10, 0)) == float("inf")


@pytest.mark.parametrize("mode", [O.DEFERRED, O.IN_LOOP])
def test_step_integration(mode):
    o = O.OracleSim(1000, 1000, 20, (-500, -500), index_mode=mode)
    ids = o.add_agents([(0, 0)], o.hl_constant((1, 0)), o.lp_none(), 100)
    assert list(ids) == [0] and o.agent_count() == 1
    o.step(1, 0)
    st = o.read_state()
    assert o.agent_count() == 1
    assert np.hypot(st["x"][0] - 1.0, st["y"][0]) < 1e-5


@pytest.mark.parametrize("mode", [O.DEFERRED, O.IN_LOOP])
def test_event_listener_source_sink_api(mode):
    o = O.OracleSim(1000, 1000, 20, (-500, -500), index_mode=mode)
    o.add_source_sink((0, 0), 1.0, 1.0, o.hl_constant((1, 0)), o.lp_none(), [(20, 0)], False, 5.0)
    added = removed = 0
    for steps in range(20):
        assert o.agent_count() == steps and added == steps
        o.step(1, 0)
        s, _, d = o.poll_events()
        added += len(s)
        removed += len(d)
    for steps in range(20, 40):
        assert o.agent_count() == 20 and added == steps and removed == steps - 20
        o.step(1, 0)
        s, _, d = o.poll_events()
        added += len(s)
        removed += len(d)


def test_duration_as_secs_f64():
    L = O.lib()
    assert L.orc_duration_as_secs_f64(0, 16_666_667) == 0.0 + 16666667.0 / 1e9
    assert L.orc_duration_as_secs_f64(3, 500_000_000) == 3.5


def test_insert_vs_query_cell_mismatch():
    """A1 truncates (negatives saturate to 0), A3 floors: an agent left of the grid origin is stored
    in column 0 but a query centred on it looks at column -1 (location_hash_2d.rs:56 vs :69)."""
    o = O.OracleSim(10, 10, 1, (0, 0))
    o.index_add_or_update(7, (-0.5, 0.5))
    assert o.cell_of([(-0.5, 0.5)])[0] == 0
    assert list(o.query_radius(0.4, (-0.5, 0.5))) == []      # stencil = column -1 only: invalid
    assert list(o.query_radius(0.6, (-0.5, 0.5))) == [7]     # stencil reaches column 0


def test_width_stride_aliasing():
    """idx = x_idx * n_x + y_idx with n_x = width cells (location_hash_2d.rs:59): for a 4 x 8 grid
    (n_x = 4, n_y = 8) cell (0, 5) aliases cell (1, 1)."""
    o = O.OracleSim(4, 8, 1, (0, 0))
    assert o.cell_of([(0.5, 5.5)])[0] == o.cell_of([(1.5, 1.5)])[0] == 5
    # and x_idx = 3, y_idx = 7 -> 19 < 32 is accepted although 3*8+7 would be the "intended" 31
    assert o.cell_of([(3.5, 7.5)])[0] == 19
    # 8 x 4 grid: n_x = 8, len = 32; x_idx = 4 -> idx = 32 + y >= len => out of bounds
    o2 = O.OracleSim(8, 4, 1, (0, 0))
    assert o2.cell_of([(4.5, 0.5)])[0] == -1
    assert o2.cell_of([(3.5, 3.5)])[0] == 27


def test_flat_parallel_cpu_port_is_bit_identical_to_the_oracle():
    """oracle/flat_parallel.cpp (the multi-threaded CPU implementation bench.py reports next to the reference-style
    baseline) against the oracle's deferred mode on a 1 600-agent crowd: t_i and the new state bit for bit, for one
    thread and for several, over three steps."""
    import oracle_ffi as O
    import parity as P
    from rmf_crowdsim_b200 import scenes as SC

    scene = SC.uniform_crowd(40, "shuffled", margin=16.0, seed=8, lp=("zanlungo", 0.05, 1.0, 0.0, 0.5, 50.0, 0.1))
    o = P.build_oracle(scene)
    o.enable_trace(True)
    st = [{k: np.ascontiguousarray(v.copy()) for k, v in zip(("x", "y", "vx", "vy"),
                                                              (scene.xy[:, 0], scene.xy[:, 1], scene.vxy[:, 0],
                                                               scene.vxy[:, 1]))} for _ in range(2)]
    finite = 0
    for _ in range(3):
        o.step(*scene.dt)
        so, tr = o.read_state(), o.read_trace()
        for s, threads in zip(st, (1, 5)):
            t_i = O.flat_step(scene, s["x"], s["y"], s["vx"], s["vy"], scene.dt, threads, want_t_i=True)
            assert np.array_equal(t_i.view(np.uint64), tr["t_i"].view(np.uint64))
            for k in ("x", "y", "vx", "vy"):
                assert np.array_equal(s[k].view(np.uint64), so[k].view(np.uint64)), (k, threads)
        finite += int(np.isfinite(tr["t_i"]).sum())
    assert finite > 100


def test_flat_port_trace_and_windows_with_real_ids():
    """The trace outputs of oracle/flat_parallel.cpp (what the benchmark-size GPU parity tests compare against): list
    lengths, t_i and forces equal the oracle's trace bit for bit; and a window cut out of the crowd, stepped with the
    agents' real ids, reproduces the full crowd's results for every agent whose neighbours are all inside."""
    import oracle_ffi as O
    import parity as P
    from rmf_crowdsim_b200 import scenes as SC

    scene = SC.uniform_crowd(40, "shuffled", margin=16.0, seed=8)
    o = P.build_oracle(scene)
    o.enable_trace(True)
    o.step(*scene.dt)
    tr, so = o.read_trace(), o.read_state()
    cols = [np.ascontiguousarray(a.copy()) for a in (scene.xy[:, 0], scene.xy[:, 1], scene.vxy[:, 0], scene.vxy[:, 1])]
    ft = O.flat_step_trace(scene, *cols, scene.dt, 3)
    assert np.array_equal(ft["nbc"].astype(np.int64), np.diff(tr["nb_offsets"].astype(np.int64)))
    for k in ("t_i", "fx", "fy"):
        assert np.array_equal(ft[k].view(np.uint64), tr[k].view(np.uint64)), k
    for c, k in zip(cols, ("x", "y", "vx", "vy")):
        assert np.array_equal(c.view(np.uint64), so[k].view(np.uint64)), k
    assert np.isfinite(ft["t_i"]).sum() > 100 and np.any(ft["fx"] != 0.0)
    # window: agents of [10, 30)^2 plus the ring within eyesight + slack, real ids (ascending)
    x0, y0 = scene.xy[:, 0], scene.xy[:, 1]
    ring = scene.eyesight + 0.5
    inner = (x0 >= 10) & (x0 < 30) & (y0 >= 10) & (y0 < 30)
    outer = (x0 >= 10 - ring) & (x0 < 30 + ring) & (y0 >= 10 - ring) & (y0 < 30 + ring)
    ids = np.nonzero(outer)[0].astype(np.uint64)
    wcols = [np.ascontiguousarray(a[outer].copy()) for a in (scene.xy[:, 0], scene.xy[:, 1], scene.vxy[:, 0],
                                                             scene.vxy[:, 1])]
    wt = O.flat_step_trace(scene, *wcols, scene.dt, 2, ids=ids)
    keep = inner[outer]
    assert keep.sum() > 300
    for k in ("t_i", "fx", "fy"):
        assert np.array_equal(wt[k][keep].view(np.uint64), ft[k][inner].view(np.uint64)), k
    assert np.array_equal(wt["nbc"][keep], ft["nbc"][inner])
    for c, full in zip(wcols, cols):
        assert np.array_equal(c[keep].view(np.uint64), full[inner].view(np.uint64))
    with pytest.raises(O.OracleError):
        O.flat_step_trace(scene, *wcols, scene.dt, 2, ids=ids[::-1].copy())
