"""GPU: SourceSink spawn / despawn stream (lib.rs:199-254, 305-336, 378-380) against the oracle and the
reference's own integration test (tests/event_listeners_test.rs:64-111)."""
import numpy as np
import pytest

import oracle_ffi as O
import parity as P
import rmf_crowdsim_b200 as R
from rmf_crowdsim_b200 import _native as _N

N_ERR_CAPACITY = _N.RCS_ERR_CAPACITY

pytestmark = pytest.mark.gpu


class Recorder(R.EventListener):
    """MockEventListener of tests/event_listeners_test.rs:37-62."""

    def __init__(self):
        self.added, self.removed = [], []

    def agent_spawned(self, position, agent):
        self.added.append((agent, position))

    def agent_destroyed(self, agent):
        self.removed.append(agent)


def test_event_listener_source_sink_api():
    """The reference test, literally: MonotonicCrowd(1.0), source (0,0), waypoint (20,0), radius_sink 1,
    eyesight 5, NoLocalPlan, stub planner (1,0), dt = 1 s."""
    sim = R.Simulation(R.LocationHash2D(1000.0, 1000.0, 20.0, (-500.0, -500.0), capacity=64))
    rec = Recorder()
    sim.add_event_listener(rec)
    ss = R.SourceSink(source=(0.0, 0.0), radius_sink=1.0, crowd_generator=R.MonotonicCrowd(1.0),
                      high_level_planner=R.ConstantVelocityPlan((1.0, 0.0)), local_planner=R.NoLocalPlan(),
                      waypoints=[(20.0, 0.0)], loop_forever=False, agent_eyesight_range=5.0)
    sim.add_source_sink(ss)
    for steps in range(20):
        assert sim.agent_count() == steps
        assert len(rec.added) == steps
        sim.step(R.Duration(1, 0))
    for steps in range(20, 40):
        assert sim.agent_count() == 20
        assert len(rec.added) == steps
        assert len(rec.removed) == steps - 20
        sim.step(R.Duration(1, 0))
    assert [a for a, _ in rec.added] == list(range(40))
    assert rec.removed == list(range(20))
    assert all(p == (0.0, 0.0) for _, p in rec.added)


def _pair(n_sources=12, zan=True, loop=False, seed=0):
    """Identical source sinks in the oracle and the CUDA simulation.
    NoLocalPlan: sources on a ring walking through a waypoint near the centre to the opposite side.
    Zanlungo: two lanes 0.15 m apart, same direction, different speeds -- the faster stream overtakes the
    slower one inside agent_radius laterally, so t_i is finite and forces act, yet the crowd stays finite
    (checked with the oracle; a crossing Zanlungo crowd goes non-finite within a few steps)."""
    rng = np.random.default_rng(seed)
    w = h = 64.0
    off = (-32.0, -32.0)
    o = O.OracleSim(w, h, 2.0, off)
    g = R.Simulation(R.LocationHash2D(w, h, 2.0, off, capacity=4096))
    za = (0.05, 1.0, 0.0, 0.5, 1.0, 0.2)
    specs = []
    if zan:
        for y, sp in [(0.0, 1.5), (0.15, 1.0)]:
            specs.append(((-20.0, y), [(0.0, y), (20.0, y)], (sp, 0.0)))
    else:
        for k in range(n_sources):
            ang = 2 * np.pi * k / n_sources
            src = (20.0 * np.cos(ang), 20.0 * np.sin(ang))
            mid = (float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-1.5, 1.5)))
            dst = (-src[0], -src[1])
            specs.append((src, [mid, dst], ((dst[0] - src[0]) / 40.0 * 1.5, (dst[1] - src[1]) / 40.0 * 1.5)))
    keep = []
    for src, wps, vel in specs:
        ohl, olp = o.hl_constant(vel), (o.lp_zanlungo(*za) if zan else o.lp_none())
        o.add_source_sink(src, 3.0, 2.0, ohl, olp, wps, loop, 2.0)
        ghl, glp = R.ConstantVelocityPlan(vel), (R.Zanlungo(*za) if zan else R.NoLocalPlan())
        keep.append((ghl, glp))
        g.add_source_sink(R.SourceSink(src, 3.0, R.MonotonicCrowd(2.0), ghl, glp, wps, loop, 2.0))
    g._keep = keep
    return g, o


@pytest.mark.parametrize("zan", [False, True])
def test_spawn_despawn_stream_matches_oracle(zan):
    g, o = _pair(zan=zan)
    rec = Recorder()
    g.add_event_listener(rec)
    spawned_o, destroyed_o = [], []
    dt = (0, 500_000_000)  # round(0.5 * 2.0) = 1 agent per source per step
    saw_despawn = False
    finite = 0
    if zan:
        g.set_trace(True)
        o.enable_trace(True)
    for step in range(85):
        if o.agent_count():
            P.resync(g, o)
        g.step(R.Duration(*dt))
        o.step(*dt)
        s, sxy, d = o.poll_events()
        spawned_o += list(s)
        destroyed_o += sorted(d)  # canonical order inside a step: ascending id
        saw_despawn |= len(d) > 0
        assert g.agent_count() == o.agent_count(), step
        r = P.compare_states(g.read_state(), o.read_state())
        assert r["vel_rel_err"] <= P.REL_TOL and r["pos_rel_err"] <= P.REL_TOL
        st = g.stats()
        assert st.spawned == len(s) and st.destroyed == len(d)
        if zan:
            tr = P.compare_traces(g.read_trace(), o.read_trace())
            assert tr["force_rel_err"] <= P.REL_TOL
            finite += tr["finite_tti"]
    assert saw_despawn and (finite > 0 or not zan)
    assert [a for a, _ in rec.added] == [int(v) for v in spawned_o]
    assert rec.removed == [int(v) for v in destroyed_o]


def test_loop_forever_resets_waypoint_and_never_despawns():
    g, o = _pair(n_sources=4, zan=False, loop=True)
    for step in range(60):
        if o.agent_count():
            P.resync(g, o)
        g.step(R.Duration(0, 500_000_000))
        o.step(0, 500_000_000)
        _, _, d = o.poll_events()
        assert len(d) == 0
        P.compare_states(g.read_state(), o.read_state())  # includes next_waypoint
    assert g.stats().destroyed == 0


def test_spawn_is_blocked_while_someone_stands_on_the_source():
    """lib.rs:212-216: no spawn while ANY agent is within 0.4 m of the source (strict <)."""
    for d, expect in [(0.39, 0), (0.4, None), (0.41, 1), (0.25, 0)]:  # 50.4 - 50.0 is 0.39999999999999858 in f64
        o = O.OracleSim(100.0, 100.0, 2.0, (0.0, 0.0))
        g = R.Simulation(R.LocationHash2D(100.0, 100.0, 2.0, (0.0, 0.0), capacity=64))
        o.add_agents([(50.0 + d, 50.0)], o.hl_constant((0.0, 0.0)), o.lp_none(), 1.0)
        g.add_agents([(50.0 + d, 50.0)], R.ConstantVelocityPlan((0.0, 0.0)), R.NoLocalPlan(), 1.0)
        o.add_source_sink((50.0, 50.0), 1.0, 1.0, o.hl_constant((1.0, 0.0)), o.lp_none(), [(90.0, 50.0)], False, 1.0)
        g.add_source_sink(R.SourceSink((50.0, 50.0), 1.0, R.MonotonicCrowd(1.0), R.ConstantVelocityPlan((1.0, 0.0)),
                                       R.NoLocalPlan(), [(90.0, 50.0)], False, 1.0))
        g.step(R.Duration(1, 0))
        o.step(1, 0)
        if expect is not None:
            assert o.agent_count() == 1 + expect
        assert g.agent_count() == o.agent_count()


def test_source_outside_the_grid_is_the_reference_error():
    g = R.Simulation(R.LocationHash2D(10.0, 10.0, 1.0, (0.0, 0.0), capacity=16))
    with pytest.raises(R.CrowdsimError) as e:
        g.add_source_sink(R.SourceSink((50.0, 50.0), 1.0, R.MonotonicCrowd(1.0), R.ConstantVelocityPlan((1.0, 0.0)),
                                       R.NoLocalPlan(), [(5.0, 5.0)], False, 1.0))
    assert str(e.value) == "Failed to add agents from source"


def test_monotonic_crowd_rounding():
    """source_sink.rs:97-100: round(dt * rate) with no carry: 0.49 never spawns, 0.5 does."""
    for dt_ns, rate, expect in [(490_000_000, 1.0, 0), (500_000_000, 1.0, 1), (16_666_667, 60.0, 1),
                                (16_666_667, 29.0, 0)]:
        g = R.Simulation(R.LocationHash2D(100.0, 100.0, 2.0, (0.0, 0.0), capacity=64))
        g.add_source_sink(R.SourceSink((50.0, 50.0), 1.0, R.MonotonicCrowd(rate), R.ConstantVelocityPlan((30.0, 0.0)),
                                       R.NoLocalPlan(), [(90.0, 50.0)], False, 1.0))
        assert R.MonotonicCrowd(rate).get_number_to_spawn(R.Duration(0, dt_ns)) == expect
        g.step(R.Duration(0, dt_ns))
        assert g.agent_count() == expect


def test_route_follower_on_the_device_matches_oracle():
    """the per-step half of RMFPlanner (rmf/mod.rs:197-215) -- unit velocity towards
    the current route point, advance within 0.1 m -- with routes supplied as polylines, spawned by source sinks
    (set_target at spawn, lib.rs:242-249) whose single waypoint is the sink."""
    w = h = 64.0
    off = (-32.0, -32.0)
    o = O.OracleSim(w, h, 2.0, off)
    g = R.Simulation(R.LocationHash2D(w, h, 2.0, off, capacity=2048))
    keep = []
    routes = [
        ((-20.0, -10.0), [(-10.0, -10.0), (-10.0, 0.0), (5.0, 0.0), (20.0, 10.0)], [(20.0, 10.0)]),
        ((20.0, 12.0), [(10.0, 12.0), (0.0, 5.0), (-15.0, 5.0)], [(-15.0, 5.0)]),
        ((0.0, -25.0), [(0.0, -12.0), (3.0, -2.0), (0.0, 20.0)], [(0.0, 20.0)]),
    ]
    for src, route, wps in routes:
        o.add_source_sink(src, 0.6, 2.0, o.hl_route(route), o.lp_none(), wps, False, 2.0)
        hl, lp = R.RouteFollowPlan(route), R.NoLocalPlan()
        keep.append((hl, lp))
        g.add_source_sink(R.SourceSink(src, 0.6, R.MonotonicCrowd(2.0), hl, lp, wps, False, 2.0))
    # one polyline per planner: a route-follower source sink with intermediate waypoints (for which the reference
    # plans a new route each, rmf/mod.rs:217-237) is refused
    with pytest.raises(R.CrowdsimError):
        g.add_source_sink(R.SourceSink((5.0, 25.0), 0.6, R.MonotonicCrowd(2.0), R.RouteFollowPlan([(9.0, 25.0), (15.0, 25.0)]),
                                       R.NoLocalPlan(), [(9.0, 25.0), (15.0, 25.0)], False, 2.0))
    destroyed = 0
    for step in range(150):
        if o.agent_count():
            P.resync(g, o)
        g.step(R.Duration(0, 500_000_000))  # unit speed: 0.5 m per step
        o.step(0, 500_000_000)
        _, _, d = o.poll_events()
        destroyed += len(d)
        assert g.agent_count() == o.agent_count(), step
        r = P.compare_states(g.read_state(), o.read_state())  # includes next_waypoint
        assert r["vel_rel_err"] <= 1e-12 and r["pos_rel_err"] <= 1e-12  # only IEEE +,-,*,/,sqrt involved
    assert destroyed > 0
    # plain agents of a route planner are not in its cache: None -> velocity 0 (rmf/mod.rs:211-214) ...
    hl = R.RouteFollowPlan([(5.0, 5.0), (9.0, 5.0)])
    g2 = R.Simulation(R.LocationHash2D(32.0, 32.0, 2.0, (0.0, 0.0), capacity=16))
    ids = g2.add_agents([(1.0, 5.0), (2.0, 9.0)], hl, R.NoLocalPlan(), 1.0)
    g2.step(R.Duration(1, 0))
    st = g2.read_state()
    assert list(st["x"]) == [1.0, 2.0] and list(st["vx"]) == [0.0, 0.0]
    # ... until set_target enters them
    g2.route_set_target([ids[0]])
    g2.step(R.Duration(1, 0))
    st = g2.read_state()
    assert list(st["x"]) == [2.0, 2.0] and list(st["vx"]) == [1.0, 0.0] and list(st["next_waypoint"]) == [0, 0]


def test_source_sink_with_a_host_evaluated_planner_moves_its_agents():
    """A SourceSink whose HighLevelPlanner is evaluated on the host: spawned agents are registered with the planner
    (set_target at spawn, lib.rs:242-249) and move from the step after their spawn on; the source is cleared, so the
    next spawns follow."""

    class East(R.HighLevelPlanner):
        def __init__(self):
            self.targets, self.removed = {}, []

        def set_target(self, agent, point, tolerance):
            self.targets[agent.agent_id] = (point, tolerance)

        def get_desired_velocity(self, agent, time):
            return (1.0, 0.0) if agent.agent_id in self.targets else None

        def remove_agent_id(self, agent):
            self.removed.append(agent)

    hl = East()
    g = R.Simulation(R.LocationHash2D(64.0, 64.0, 2.0, (0.0, 0.0), capacity=256))
    g.add_source_sink(R.SourceSink((10.0, 10.0), 0.6, R.MonotonicCrowd(1.0), hl, R.NoLocalPlan(), [(14.0, 10.0)], False,
                                   2.0))
    with pytest.raises(R.CrowdsimError):
        g.add_source_sink(R.SourceSink((10.0, 10.0), 0.6, R.MonotonicCrowd(1.0), East(), R.NoLocalPlan(), [(14.0, 10.0)],
                                       False, 2.0))
    for _ in range(12):
        g.step(R.Duration(1, 0))
    st = g.read_state()
    assert len(hl.targets) >= 6 and hl.targets[0] == ((14.0, 10.0), (0.6, 0.6))
    assert hl.removed and hl.removed[0] == 0          # the first agent reached the sink at x = 14
    assert np.all(st["vx"][:-1] == 1.0) and np.all(np.diff(st["x"]) < 0)  # single file, one metre apart


def test_add_agents_after_async_steps_that_spawned_checks_the_real_count():
    """rcs_add_agents must count the agents that steps still in flight have spawned (they are only known on the
    device until the sync) before it checks the capacity and picks the insert offset."""
    cap = 24
    g = R.Simulation(R.LocationHash2D(64.0, 64.0, 2.0, (0.0, 0.0), capacity=cap))
    hl, lp = R.ConstantVelocityPlan((1.0, 0.0)), R.NoLocalPlan()
    g._keep = (hl, lp)
    g.add_source_sink(R.SourceSink((10.0, 10.0), 0.6, R.MonotonicCrowd(1.0), hl, lp, [(60.0, 10.0)], False, 2.0))
    for _ in range(20):
        g.step_async(R.Duration(1, 0))  # 20 spawns, none of them seen by the host yet
    with pytest.raises(R.CrowdsimError) as e:
        g.add_agents(np.full((5, 2), 30.0) + np.arange(5)[:, None], hl, lp, 2.0)  # 20 + 5 > 24
    assert e.value.code == N_ERR_CAPACITY
    assert g.agent_count() == 20
    ids = g.add_agents(np.full((4, 2), 30.0) + np.arange(4)[:, None], hl, lp, 2.0)
    assert list(ids) == [20, 21, 22, 23] and g.agent_count() == 24
    st = g.read_state()
    assert np.array_equal(st["id"], np.arange(24, dtype=np.uint64))
    assert np.array_equal(st["x"][:20], 10.0 + np.arange(20, 0, -1))  # the spawned agents are intact
