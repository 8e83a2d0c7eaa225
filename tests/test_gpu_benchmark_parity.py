"""GPU: the CUDA path against the CPU oracle AT BENCHMARK SIZE (BASELINE.json configs 3, 4 and 5).

The literal oracle (hash maps, one thread) needs minutes for 2^20 agents, so the crowds of configs 3 and 4 are
checked against oracle/flat_parallel.cpp -- the oracle's own Zanlungo arithmetic and radius test on flat arrays and
all host cores, proven bit-identical to the oracle's deferred mode by tests/test_oracle_reference.py (CPU suite):
  * C3: all 2^20 agents, two committed steps (the second from the oracle's state);
  * C4: the whole 2^24-agent crowd is stepped on the GPU, a 1024 m x 1024 m window of it (2^20 agents, plus the ring
    of agents within eyesight of the window) by the oracle with the agents' real ids;
  * C5: a SourceSink stream with > 100 000 live agents, spawning and despawning every step, device-side route
    follower, for 60 steps after the fill against the literal oracle (lib.rs:199-254, 305-336, 378-380).
Tolerances are north_star's: neighbour-list lengths equal, t_i bit-exact, forces / velocities / positions <= 1e-9
relative (CUDA's exp differs from glibc's by <= 2 ulp; everything else is the same IEEE operation).
"""
import os

import numpy as np
import pytest

import oracle_ffi as O
import parity as P
import rmf_crowdsim_b200 as R
from rmf_crowdsim_b200 import scenes as SC

pytestmark = pytest.mark.gpu

THREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _cols(scene, sel=None):
    xy, vxy = scene.xy, scene.vxy
    if sel is not None:
        xy, vxy = xy[sel], vxy[sel]
    return [np.ascontiguousarray(a, dtype=np.float64) for a in (xy[:, 0], xy[:, 1], vxy[:, 0], vxy[:, 1])]


def _check(tr_g, st_g, tr_o, cols_o, sel_g=None):
    """GPU trace / state (ascending id, optionally restricted by sel_g) against the flat oracle's arrays."""
    pick = (lambda a: a) if sel_g is None else (lambda a: a[sel_g])
    nbc_g = np.diff(tr_g["nb_offsets"].astype(np.int64))
    assert np.array_equal(pick(nbc_g), tr_o["nbc"].astype(np.int64)), "neighbour-list lengths differ"
    assert np.array_equal(pick(tr_g["t_i"]).view(np.uint64), tr_o["t_i"].view(np.uint64)), "t_i not bit-exact"
    fmag = np.sqrt(tr_o["fx"] ** 2 + tr_o["fy"] ** 2)
    fmag = np.where(np.isfinite(fmag), fmag, 0.0)
    ef = max(P.rel_err(pick(tr_g["fx"]), tr_o["fx"], scale=fmag), P.rel_err(pick(tr_g["fy"]), tr_o["fy"], scale=fmag))
    x, y, vx, vy = cols_o
    vmag = np.sqrt(vx ** 2 + vy ** 2)
    vmag = np.where(np.isfinite(vmag), vmag, 0.0)
    ev = max(P.rel_err(pick(st_g["vx"]), vx, scale=vmag), P.rel_err(pick(st_g["vy"]), vy, scale=vmag))
    ep = max(P.rel_err(pick(st_g["x"]), x), P.rel_err(pick(st_g["y"]), y))
    assert ef <= P.REL_TOL and ev <= P.REL_TOL and ep <= P.REL_TOL, (ef, ev, ep)
    return int(np.isfinite(tr_o["t_i"]).sum())


def test_c3_all_agents_against_the_flat_oracle():
    scene = SC.config_c3("shuffled")
    assert scene.n == 1 << 20
    g = SC.build_simulation(scene)
    g.set_trace(True)
    cols = _cols(scene)
    finite = 0
    for step in range(2):
        if step:  # identical inputs for every compared step
            g.set_state(None, *cols)
        tr_o = O.flat_step_trace(scene, *cols, scene.dt, THREADS)
        g.step(R.Duration(*scene.dt))
        tr_g, st_g = g.read_trace(neighbours=False), g.read_state()
        assert np.array_equal(tr_g["id"], np.arange(scene.n, dtype=np.uint64))
        finite += _check(tr_g, st_g, tr_o, cols)
        assert int(tr_g["nb_offsets"][-1]) == g.stats().neighbour_total == int(tr_o["nbc"].sum())
    assert finite > 0.3 * scene.n  # the force pass was busy: ~42 % of the agents have a finite t_i


def test_c4_window_of_2_pow_20_agents_against_the_flat_oracle():
    scene = SC.config_c4("shuffled")
    assert scene.n == 1 << 24
    g = SC.build_simulation(scene)
    g.set_trace(True)
    g.step(R.Duration(*scene.dt))
    tr_g, st_g = g.read_trace(neighbours=False), g.read_state()
    assert np.array_equal(tr_g["id"], np.arange(scene.n, dtype=np.uint64))
    del g
    # window [lo, hi)^2 of the start-of-step positions and everybody within eyesight (+ slack) of it
    lo, hi, ring = 1536.0, 2560.0, scene.eyesight + 0.5
    x0, y0 = scene.xy[:, 0], scene.xy[:, 1]
    inner = (x0 >= lo) & (x0 < hi) & (y0 >= lo) & (y0 < hi)
    outer = (x0 >= lo - ring) & (x0 < hi + ring) & (y0 >= lo - ring) & (y0 < hi + ring)
    assert int(inner.sum()) == 1 << 20
    ids = np.nonzero(outer)[0].astype(np.uint64)  # array index == id (scenes.uniform_crowd): ascending
    cols = _cols(scene, outer)
    tr_o = O.flat_step_trace(scene, *cols, scene.dt, THREADS, ids=ids)
    keep = inner[outer]
    tr_o = {k: v[keep] for k, v in tr_o.items()}
    cols = [c[keep] for c in cols]
    finite = _check(tr_g, st_g, tr_o, cols, sel_g=inner)
    assert finite > 0.3 * (1 << 20)


def _stream_pair(lp_zanlungo: bool, cols=64, rows=64):
    """cols x rows source sinks, one spawn per source and step (MonotonicCrowd(2/s), dt = 0.5 s; the previous agent
    has walked 0.5 m > 0.4 m by then), 13.5 m two-segment routes followed on the device, sink radius 0.6 m:
    ~27 live agents per source in steady state."""
    pitch_x, pitch_y, margin = 18.0, 4.0, 16.0
    dom = float(np.ceil((max(cols * pitch_x, rows * pitch_y) + 2 * margin) / 2.0) * 2.0)
    o = O.OracleSim(dom, dom, 2.0, (-margin, -margin))
    g = R.Simulation(R.LocationHash2D(dom, dom, 2.0, (-margin, -margin), capacity=cols * rows * 40))
    za = (0.05, 1.0, 0.0, 0.5, 1.0, 0.2)
    keep = []
    for c in range(cols):
        for r in range(rows):
            x0, y0 = c * pitch_x + 1.0, r * pitch_y + 1.0
            route = [(x0 + 6.03, y0 + 0.7), (x0 + 13.51, y0)]
            o.add_source_sink((x0, y0), 0.6, 2.0, o.hl_route(route), o.lp_zanlungo(*za) if lp_zanlungo else o.lp_none(),
                              [route[-1]], False, 2.0)
            hl, lp = R.RouteFollowPlan(route), (R.Zanlungo(*za) if lp_zanlungo else R.NoLocalPlan())
            keep.append((hl, lp))
            g.add_source_sink(R.SourceSink((x0, y0), 0.6, R.MonotonicCrowd(2.0), hl, lp, [route[-1]], False, 2.0))
    g._keep = keep
    return g, o


def test_c5_stream_of_100k_live_agents_against_the_oracle():
    g, o = _stream_pair(False)
    dt = (0, 500_000_000)
    n_src = 64 * 64
    spawned = destroyed = 0
    checked = 0
    for step in range(90):
        g.step_async(R.Duration(*dt))
        o.step(*dt)
        s, _, d = o.poll_events()
        spawned += len(s)
        destroyed += len(d)
        g.sync()
        st = g.stats()
        assert st.spawned == len(s) and st.destroyed == len(d), step
        g._dispatch_events()  # drains the device-side event lists (4096 spawns + despawns per step)
        if step < 30 and step % 10 != 9:
            continue  # the fill: compared every tenth step
        # > 100 000 live agents from here on; every step is compared
        so, sg = o.read_state(), g.read_state()
        assert np.array_equal(sg["id"], so["id"]), step
        # NoLocalPlan + route follower: only IEEE +,-,*,/,sqrt -- bit for bit
        for k in ("x", "y", "vx", "vy"):
            assert np.array_equal(sg[k].view(np.uint64), so[k].view(np.uint64)), (k, step)
        assert np.array_equal(sg["next_waypoint"].astype(np.uint64), so["next_waypoint"].astype(np.uint64))
        if step >= 30:
            assert len(so["id"]) > 100_000, (step, len(so["id"]))
            checked += 1
    assert checked == 60 and destroyed > 50 * n_src // 2 and spawned > 80 * n_src


def test_c5_stream_with_zanlungo_against_the_oracle():
    """The same stream with the Zanlungo planner on a 32 x 32 lattice (~27 000 live agents): agents of one source
    walk in single file, so t_i is finite only where a faster follower closes in after a route bend; states are
    re-synchronised before every compared step (forces carry exp)."""
    g, o = _stream_pair(True, cols=32, rows=32)
    dt = (0, 500_000_000)
    g.set_trace(True)
    o.enable_trace(True)
    finite = 0
    for step in range(45):
        if o.agent_count():
            P.resync(g, o)
        g.step(R.Duration(*dt))
        o.step(*dt)
        s, _, d = o.poll_events()
        assert g.agent_count() == o.agent_count(), step
        r = P.compare_states(g.read_state(), o.read_state())
        assert r["vel_rel_err"] <= P.REL_TOL and r["pos_rel_err"] <= P.REL_TOL
        if step % 5 == 4:
            tr = P.compare_traces(g.read_trace(), o.read_trace())
            assert tr["force_rel_err"] <= P.REL_TOL
            finite += tr["finite_tti"]
    assert g.agent_count() > 20_000
