"""GPU: spatial strips through the single-process transport: every rank count gives
results BIT-IDENTICAL to a single handle, and the neighbour lists / t_i of owned agents match the oracle."""
import numpy as np
import pytest

import parity as P
import rmf_crowdsim_b200 as R
from rmf_crowdsim_b200 import scenes as SC
from rmf_crowdsim_b200.strips import LocalStripGroup, strip_columns

pytestmark = pytest.mark.gpu


def _bits(st):
    return {k: (v.view(np.uint64) if v.dtype == np.float64 else v) for k, v in st.items()}


def _same(a, b):
    a, b = _bits(a), _bits(b)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_strips_are_bit_identical_to_one_handle_lane_crowd(world):
    """40 committed steps of the lane-ordered crowd: agents stream across every strip boundary (1.3 m/s,
    cells of 2 m), so migration through the redundant ring is exercised in both directions."""
    scene = SC.uniform_crowd(40, "lane", margin=8.0, seed=9)
    single = SC.build_simulation(scene)
    grp = LocalStripGroup(scene, world)
    dt = R.Duration(0, 100_000_000)  # 0.13 m per step: a boundary crossing every few steps
    before = [set(sm.read_state()["id"].tolist()) for sm in grp.sims]
    for _ in range(40):
        single.step(dt)
        grp.step(dt)
    _same(single.read_state(), grp.read_state())
    assert sum(grp.agent_counts()) == scene.n
    after = [set(sm.read_state()["id"].tolist()) for sm in grp.sims]
    assert world == 1 or all(a != b for a, b in zip(after, before))  # agents really migrated, in both directions
    assert set().union(*after) == set(range(scene.n))


@pytest.mark.parametrize("world", [2, 4])
def test_strips_shuffled_crowd_forces_match_single_handle_and_oracle(world):
    scene = SC.uniform_crowd(48, "shuffled", margin=8.0, seed=4)
    single = SC.build_simulation(scene)
    grp = LocalStripGroup(scene, world)
    single.set_trace(True)
    grp.set_trace(True)
    dt = R.Duration(*scene.dt)
    for _ in range(2):
        single.step(dt)
        grp.step(dt)
        tg, ts = grp.read_trace(), single.read_trace()
        for k in ("id", "nb_offsets", "nb_ids"):
            assert np.array_equal(tg[k], ts[k]), k
        for k in ("t_i", "fx", "fy"):
            assert np.array_equal(tg[k].view(np.uint64), ts[k].view(np.uint64)), k
        _same(single.read_state(), grp.read_state())
    # first step against the oracle (identical inputs): neighbour sets and t_i bit-exact, forces 1e-9
    grp2 = LocalStripGroup(scene, world)
    grp2.set_trace(True)
    o2 = P.build_oracle(scene)
    o2.enable_trace(True)
    grp2.step(dt)
    o2.step(*scene.dt)
    r = P.compare_traces(grp2.read_trace(), o2.read_trace())
    assert r["finite_tti"] > 0 and r["force_rel_err"] <= P.REL_TOL
    s = P.compare_states(grp2.read_state(), o2.read_state())
    assert s["vel_rel_err"] <= P.REL_TOL and s["pos_rel_err"] <= P.REL_TOL


@pytest.mark.parametrize("cell,eyesight,s", [(1.0, 2.0, 1.0), (4.0, 2.0, 0.5), (0.5, 1.7, 1.0)])
def test_strips_with_wide_or_crowded_stencils_match_a_single_handle(cell, eyesight, s):
    """eyesight > cell (halo reach of several columns) and crowded cells (chunked cooperative kernel) on three
    ranks: neighbour lists, t_i, forces and state bit-identical to one handle."""
    rng = np.random.default_rng(23)
    scene = SC.uniform_crowd(48, "shuffled", s=s, cell=cell, eyesight=eyesight, margin=8.0, seed=5,
                             lp=("zanlungo", 0.05, 1.0, 0.0, 0.5, 50.0, 0.05))  # heavy agents: the crowd stays sane
    scene.vxy = scene.vxy + rng.uniform(-0.3, 0.3, size=scene.vxy.shape)
    single = SC.build_simulation(scene)
    grp = LocalStripGroup(scene, 3)
    single.set_trace(True)
    grp.set_trace(True)
    dt = R.Duration(0, 10_000_000)
    finite = 0
    for _ in range(3):
        single.step(dt)
        grp.step(dt)
        tg, ts = grp.read_trace(), single.read_trace()
        for k in ("id", "nb_offsets", "nb_ids"):
            assert np.array_equal(tg[k], ts[k]), k
        for k in ("t_i", "fx", "fy"):
            assert np.array_equal(tg[k].view(np.uint64), ts[k].view(np.uint64)), k
        finite += int(np.isfinite(ts["t_i"]).sum())
        _same(single.read_state(), grp.read_state())
    assert finite > 50


def test_agents_walk_into_strips_that_started_empty():
    """The crowd occupies the two middle strips of four; the outer ranks own nothing at first (they still register
    the crowd's planner group: ghosts carry group numbers), receive their first ghosts after a few steps and adopt
    the agents that walk in.  Bit-identical to one handle throughout."""
    scene = SC.uniform_crowd(24, "lane", margin=24.0, seed=17)
    single = SC.build_simulation(scene)
    grp = LocalStripGroup(scene, 4)
    assert grp.agent_counts()[0] == 0 and grp.agent_counts()[3] == 0
    dt = R.Duration(0, 500_000_000)  # 0.65 m per step
    for k in range(30):
        single.step(dt)
        grp.step(dt)
        if k % 10 == 9:
            _same(single.read_state(), grp.read_state())
    counts = grp.agent_counts()
    assert counts[0] > 0 and counts[3] > 0 and sum(counts) == scene.n


def test_no_commit_keeps_the_owned_snapshot():
    scene = SC.uniform_crowd(32, "shuffled", margin=8.0, seed=2)
    grp = LocalStripGroup(scene, 3)
    before = grp.read_state()
    for _ in range(3):
        grp.step(R.Duration(*scene.dt), no_commit=True)
    _same(before, grp.read_state())


def test_strip_ranges_tile_the_grid():
    sim = R.Simulation(R.LocationHash2D(100.0, 100.0, 2.0, (0.0, 0.0), capacity=16))
    for world in (1, 2, 3, 7, 8):
        edges = [strip_columns(sim, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 50
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))


def test_agent_outside_the_strip_is_rejected():
    scene = SC.uniform_crowd(16, "lane", margin=8.0)
    grp = LocalStripGroup(scene, 2)
    sm = grp.sims[0]
    with pytest.raises(R.CrowdsimError):
        # column far right belongs to rank 1
        from rmf_crowdsim_b200.strips import add_agents_with_ids
        add_agents_with_ids(sm, np.array([10**6], dtype=np.uint64), np.array([[20.0, 3.0]]), None,
                            *sm._scene_planners, 2.0)


def test_failure_on_one_rank_stops_the_group():
    """An agent pushed out of the grid on the last rank: that rank reports the reference's error, the other
    ranks stop within `world` steps through the flag in the halo header."""
    scene = SC.uniform_crowd(16, "lane", margin=8.0)
    scene.hl = ("constant", (50.0, 0.0))  # 5 m per step to the right: leaves the 32 m grid quickly
    grp = LocalStripGroup(scene, 2)
    codes = set()
    with pytest.raises(R.CrowdsimError) as e:
        for _ in range(20):
            grp.step(R.Duration(0, 100_000_000))
    codes.add(e.value.code)
    assert codes & {R._native.RCS_ERR_OUT_OF_BOUNDS, R._native.RCS_ERR_HALO}


@pytest.mark.parametrize("seed", range(8))
def test_random_strip_configurations_match_one_handle(seed):
    """Seeded random grid (cell size against eyesight: halo reach 1-4 columns), crowd size, rank count and step length;
    committed steps with agents crossing strip boundaries in both directions.  Bit-identical to one handle."""
    rng = np.random.default_rng(700 + seed)
    cell = float(rng.choice([1.0, 2.0, 3.0]))
    eyesight = float(rng.choice([1.5, 2.0, 2.7]))
    side = int(rng.choice([40, 48, 56]))
    world = int(rng.choice([2, 3, 4, 5]))
    # lane-ordered: a shuffled crowd walks through itself within a dozen committed steps and the model then throws
    # its 1e15 forces; the velocity jitter below still puts neighbours on collision courses
    scene = SC.uniform_crowd(side, "lane", cell=cell, eyesight=eyesight, margin=12.0, seed=50 + seed,
                             lp=("zanlungo", 0.05, 1.0, 0.0, 0.5, 100.0, 0.05))
    scene.vxy = scene.vxy + rng.uniform(-0.3, 0.3, size=scene.vxy.shape)
    single = SC.build_simulation(scene)
    grp = LocalStripGroup(scene, world)
    single.set_trace(True)
    grp.set_trace(True)
    dt = R.Duration(0, int(rng.choice([100_000_000, 200_000_000])))
    for k in range(12):
        single.step(dt)
        grp.step(dt)
        if k % 4 == 3:
            tg, ts = grp.read_trace(), single.read_trace()
            for key in ("id", "nb_offsets", "nb_ids"):
                assert np.array_equal(tg[key], ts[key]), (key, k)
            for key in ("t_i", "fx", "fy"):
                assert np.array_equal(tg[key].view(np.uint64), ts[key].view(np.uint64)), (key, k)
            _same(single.read_state(), grp.read_state())
    assert sum(grp.agent_counts()) == scene.n


def _stream_scene():
    """An empty 96 m x 96 m domain (48 cell columns of 2 m): the agents all come from source sinks."""
    return SC.Scene(name="stream", width=96.0, height=96.0, cell=2.0, offset=(0.0, 0.0), xy=np.zeros((0, 2)),
                    vxy=np.zeros((0, 2)), eyesight=2.0, hl=("constant", (0.0, 0.0)),
                    lp=("zanlungo", 0.05, 1.0, 0.0, 0.5, 1.0, 0.2), seed=0)


def _stream_sources(zan):
    """Eight source sinks whose agents walk +x or -x across the strip boundaries (columns 16 and 32 for three
    ranks), two of them as an overtaking pair 0.15 m apart (finite t_i with the Zanlungo planner); route followers
    and constant-velocity planners; the sinks lie in other strips than the sources."""
    za = (0.05, 1.0, 0.0, 0.5, 1.0, 0.2)
    specs = []
    for k, (x0, x1, y) in enumerate([(5.0, 85.0, 10.0), (85.0, 9.0, 20.0), (41.0, 90.0, 30.0), (53.0, 7.0, 40.0),
                                     (11.0, 70.0, 50.0), (11.0, 70.0, 50.15), (75.0, 40.0, 60.0), (21.0, 50.0, 70.0)]):
        sp = 1.5 if k == 4 else 1.0  # the faster lane has the lower ids (it yields): stays finite, as in _pair()
        v = (sp if x1 > x0 else -sp, 0.0)
        # route followers in single file go NaN under the Zanlungo planner: constant planners there
        route = k % 2 == 0 and k != 4 and not zan
        specs.append(((x0, y), [((x0 + x1) / 2, y + (0.5 if route else 0.0)), (x1, y)], v, route))

    def maker(spec):
        src, wps, v, route = spec

        def make():
            hl = R.RouteFollowPlan(wps) if route else R.ConstantVelocityPlan(v)
            lp = R.Zanlungo(*za) if zan else R.NoLocalPlan()
            return R.SourceSink(src, 0.6, R.MonotonicCrowd(2.0), hl, lp, [wps[-1]], False, 2.0)
        return make
    return [maker(sp) for sp in specs]


class _Rec(R.EventListener):
    def __init__(self):
        self.added, self.removed = [], []

    def agent_spawned(self, position, agent):
        self.added.append((agent, position))

    def agent_destroyed(self, agent):
        self.removed.append(agent)


@pytest.mark.parametrize("zan", [False, True])
def test_source_sinks_on_strips_match_one_handle(zan):
    """SourceSink spawn / despawn on a strip-partitioned crowd (lib.rs:199-254, 305-336): each rank spawns for the
    sources in its columns, the ids come from the step's global spawn set (a bitmap summed over the ranks), agents
    migrate through the strips to sinks owned by other ranks.  Ids, states and events equal one handle's."""
    scene = _stream_scene()
    single = SC.build_simulation(scene, capacity=4096)
    grp = LocalStripGroup(scene, 3, capacity=4096, halo_capacity=1024)
    rec_s, rec_g = _Rec(), _Rec()
    single.add_event_listener(rec_s)
    grp.add_event_listener(rec_g)
    keep = []
    for make in _stream_sources(zan):
        ss = make()
        keep.append(ss)
        single.add_source_sink(ss)
        grp.add_source_sink(make)
    dt = R.Duration(0, 500_000_000)
    for step in range(130):
        single.step(dt)
        grp.step(dt)
        grp.dispatch_events()
        if step % 10 == 9 or step > 120:
            sa, sb = single.read_state(), grp.read_state()
            assert np.array_equal(sa["id"], sb["id"]), step
            for k in ("x", "y", "vx", "vy"):
                assert np.array_equal(sa[k].view(np.uint64), sb[k].view(np.uint64)), (k, step)
            assert np.array_equal(sa["next_waypoint"], sb["next_waypoint"])
    assert sorted(rec_g.added) == sorted(rec_s.added) and len(rec_s.added) > 500
    assert sorted(rec_g.removed) == sorted(rec_s.removed) and len(rec_s.removed) > 100
    counts = grp.agent_counts()
    assert sum(counts) == single.agent_count() and min(counts) > 0


def test_source_sink_next_to_a_strip_boundary_is_refused():
    scene = _stream_scene()
    grp = LocalStripGroup(scene, 3, capacity=256, halo_capacity=128)

    def make(x):
        return lambda: R.SourceSink((x, 10.0), 0.6, R.MonotonicCrowd(2.0), R.ConstantVelocityPlan((1.0, 0.0)),
                                    R.NoLocalPlan(), [(x + 5.0, 10.0)], False, 2.0)
    with pytest.raises(R.CrowdsimError):
        grp.add_source_sink(make(32.2))  # column 16 is rank 1's first: the 0.4 m probe reaches into column 15
    grp.add_source_sink(make(33.0))      # probe columns 16..16
    grp.add_source_sink(make(0.1))       # the domain's edge is nobody's boundary


def test_caller_supplied_strip_boundaries_match_one_handle():
    """rcs_dist_set_boundaries: strips balanced by agent count instead of equal column counts (the crowd does not fill
    the grid) give the single handle's bits as well; bad boundaries are refused."""
    scene = SC.uniform_crowd(48, "lane", margin=16.0, seed=9)  # 80 m domain, 40 columns, agents in columns 8..31
    single = SC.build_simulation(scene)
    bounds = [0, 16, 24, 40]
    grp = LocalStripGroup(scene, 3, capacity=scene.n, halo_capacity=2048, boundaries=bounds)
    assert [strip_columns(sm, r, 3) for r, sm in enumerate(grp.sims)] == [(0, 16), (16, 24), (24, 40)]
    dt = R.Duration(0, 200_000_000)
    for _ in range(12):
        single.step(dt)
        grp.step(dt)
    sa, sb = single.read_state(), grp.read_state()
    assert np.array_equal(sa["id"], sb["id"])
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(sa[k].view(np.uint64), sb[k].view(np.uint64)), k
    with pytest.raises(R.CrowdsimError):
        LocalStripGroup(scene, 3, capacity=scene.n, halo_capacity=2048, boundaries=[0, 24, 16, 40])
