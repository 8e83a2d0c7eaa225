"""GPU: steady-state steps as CUDA graphs (RCS_OPT_GRAPHS, rcs_host_step.inl).

A step whose launch sequence repeats is captured the second time its key comes up and replayed as one graph launch
from then on; the host-side half of the step (buffer roles, pending list, counters) is replayed next to it.  Every
case here runs the same calls on a handle with graphs and on one without and compares every bit, and checks through
rcs_graph_stats that the graph path really ran."""
import numpy as np
import pytest

import rmf_crowdsim_b200 as R
from rmf_crowdsim_b200 import _native as N
from rmf_crowdsim_b200 import scenes as SC

pytestmark = pytest.mark.gpu


def _pair(scene):
    sims = []
    for on in (1, 0):
        g = SC.build_simulation(scene)
        g.set_option(N.RCS_OPT_GRAPHS, on)
        sims.append(g)
    return sims


def _same(a, b, order=N.RCS_ORDER_ID):
    sa, sb = a.read_state(order=order), b.read_state(order=order)
    for k in ("id", "x", "y", "vx", "vy", "next_waypoint"):
        assert np.array_equal(sa[k].view(np.uint64) if sa[k].dtype == np.float64 else sa[k],
                              sb[k].view(np.uint64) if sb[k].dtype == np.float64 else sb[k]), k
    assert a.stats().neighbour_total == b.stats().neighbour_total
    assert a.stats().steps == b.stats().steps


def test_committed_steps_replay_one_graph():
    scene = SC.uniform_crowd(96, "lane", margin=16.0, seed=2)
    a, b = _pair(scene)
    dt = R.Duration(0, 100_000_000)
    for g in (a, b):
        for _ in range(14):
            g.step_async(dt)
        g.sync()
    _same(a, b, N.RCS_ORDER_STORAGE)
    launches, captures = a.graph_stats()
    # step 1 uploads counts and groups (not steady), step 2 runs kernel by kernel, step 3 is captured, 11 replays
    assert captures == 1 and launches == 11
    assert b.graph_stats() == (0, 0)
    assert a.launch_count() == b.launch_count()  # kernels inside graph launches are counted
    # a different dt is a different key; a state injection starts a new epoch
    for g in (a, b):
        for _ in range(5):
            g.step_async(R.Duration(0, 50_000_000))
        g.sync()
    _same(a, b)
    assert a.graph_stats() == (14, 2)
    st = a.read_state()
    for g in (a, b):  # a state injection changes the data, not the launch sequence: the first graph is still good
        g.set_state(None, st["x"] + 0.125, st["y"], st["vx"], st["vy"])
        for _ in range(6):
            g.step_async(dt)
        g.sync()
    _same(a, b)
    assert a.graph_stats() == (20, 2)


def test_frozen_steps_alternate_between_two_graphs():
    scene = SC.uniform_crowd(64, "shuffled", margin=16.0, seed=4)
    a, b = _pair(scene)
    dt = R.Duration(*scene.dt)
    before = a.read_state()
    for g in (a, b):
        for _ in range(12):
            g.step_async(dt, no_commit=True)
        g.sync()
    _same(a, b)
    after = a.read_state()
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(before[k].view(np.uint64), after[k].view(np.uint64))
    assert a.stats().finite_tti_count == b.stats().finite_tti_count > 0
    launches, captures = a.graph_stats()
    assert captures == 2 and launches == 7  # steps 4 and 5 are captured (the two buffer roles), 6..12 replay
    # one committed step from the frozen snapshot, with and without graphs
    for g in (a, b):
        g.step(dt)
    _same(a, b)


def test_streaming_path_and_trace_toggle():
    scene = SC.uniform_crowd(64, "shuffled", margin=16.0, seed=5, lp=("none",))
    a, b = _pair(scene)
    dt = R.Duration(0, 200_000_000)
    for g in (a, b):
        for _ in range(9):
            g.step_async(dt)
        g.sync()
    _same(a, b)
    assert a.graph_stats() == (4, 2)  # position / velocity buffers alternate: two graphs
    for g in (a, b):
        g.set_trace(True)  # traced steps run kernel by kernel (and through the index)
        g.step(dt)
        g.set_trace(False)
        for _ in range(5):
            g.step_async(dt)
        g.sync()
    _same(a, b)


def test_source_sink_stream_with_graphs():
    """Churn: spawns and despawns every step; the launch bound grows until it reaches the capacity, from then on the
    steps repeat.  A sync in the middle compacts the arrays (new buffer roles, new counts: new graphs)."""
    sims = []
    for on in (1, 0):
        g = R.Simulation(R.LocationHash2D(64.0, 64.0, 2.0, (0.0, 0.0), capacity=96))
        g.set_option(N.RCS_OPT_GRAPHS, on)
        keep = []
        for k in range(4):
            hl, lp = R.ConstantVelocityPlan((1.0, 0.0)), R.Zanlungo(0.05, 1.0, 0.0, 0.5, 1.0, 0.2)
            keep.append((hl, lp))
            g.add_source_sink(R.SourceSink((4.0, 8.0 + 10.0 * k), 0.6, R.MonotonicCrowd(2.0), hl, lp,
                                           [(14.0, 8.0 + 10.0 * k)], False, 2.0))
        g._keep = keep
        sims.append(g)
    a, b = sims
    dt = R.Duration(0, 500_000_000)
    for rounds in range(3):
        for g in (a, b):
            for _ in range(40):
                g.step_async(dt)
            g.sync()
        assert a.agent_count() == b.agent_count() > 0
        _same(a, b)
        a._dispatch_events()  # drain the event lists
        b._dispatch_events()
    assert a.graph_stats()[0] > 40
