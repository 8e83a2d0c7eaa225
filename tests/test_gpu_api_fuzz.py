"""GPU: seeded random SEQUENCES of API calls -- step, add_agents, remove_agents, set_state, source sinks spawning and
despawning in between -- against the oracle after every call.  The arithmetic is covered elsewhere; this is about the
host-side bookkeeping around it: device-side counts and upper bounds, lazy compaction of flagged agents, id -> slot
tables, sequential id allocation shared by add_agents and the source sinks (lib.rs:128-129), event lists."""
import numpy as np
import pytest

import oracle_ffi as O
import parity as P
import rmf_crowdsim_b200 as R

pytestmark = pytest.mark.gpu

ZAN = (0.05, 1.0, 0.0, 0.5, 200.0, 0.05)


class Recorder(R.EventListener):
    def __init__(self):
        self.added, self.removed = [], []

    def agent_spawned(self, position, agent):
        self.added.append((int(agent), tuple(position)))

    def agent_destroyed(self, agent):
        self.removed.append(int(agent))


def _check(g, o, where):
    assert g.agent_count() == o.agent_count(), where
    sg, so = g.read_state(), o.read_state()
    assert np.array_equal(sg["id"], so["id"]), where
    r = P.compare_states(sg, so)
    assert r["vel_rel_err"] <= P.REL_TOL and r["pos_rel_err"] <= P.REL_TOL, where


@pytest.mark.parametrize("seed", range(6))
def test_random_call_sequences_keep_the_bookkeeping_in_step_with_the_oracle(seed):
    rng = np.random.default_rng(500 + seed)
    w = h = 64.0
    off = (0.0, 0.0)
    o = O.OracleSim(w, h, 2.0, off)
    g = R.Simulation(R.LocationHash2D(w, h, 2.0, off, capacity=4096))
    rec = Recorder()
    g.add_event_listener(rec)
    # group A: a Zanlungo lattice in the lower half; group B: NoLocalPlan walkers added in the upper half
    hl_a, lp_a = R.ConstantVelocityPlan((0.3, 0.1)), R.Zanlungo(*ZAN)
    hl_b, lp_b = R.ParityVelocityPlan((0.5, 0.0)), R.NoLocalPlan()
    ohl_a, olp_a = o.hl_constant((0.3, 0.1)), o.lp_zanlungo(*ZAN)
    ohl_b, olp_b = o.hl_parity((0.5, 0.0)), o.lp_none()
    gi, gj = np.meshgrid(np.arange(18), np.arange(12), indexing="ij")
    lattice = np.stack([gi.reshape(-1), gj.reshape(-1)], axis=1) * 1.5 + 8.0 + rng.uniform(-0.3, 0.3, (216, 2))
    assert list(o.add_agents(lattice, ohl_a, olp_a, 2.0)) == g.add_agents(lattice, hl_a, lp_a, 2.0)
    # two source sinks whose agents walk 6 m along y = 34 / y = 36 and leave at the second waypoint
    keep = []
    for y, speed in [(34.0, 1.0), (36.0, 1.6)]:
        hl, ohl = R.ConstantVelocityPlan((speed, 0.0)), o.hl_constant((speed, 0.0))
        keep.append(hl)
        sid_o = o.add_source_sink((6.0, y), 0.6, 10.0, ohl, olp_b, [(9.0, y), (12.0, y)], False, 1.0)
        sid_g = g.add_source_sink(R.SourceSink((6.0, y), 0.6, R.MonotonicCrowd(10.0), hl, lp_b, [(9.0, y), (12.0, y)],
                                               False, 1.0))
        assert sid_o == sid_g
    dt = (0, 100_000_000)
    spawned_o, destroyed_o = [], []

    def drain(step_events):
        """The oracle's listener calls since the last drain.  Inside a step the reference's removal order is
        HashMap-random; the canonical order is ascending id.  Calls made by add_agents / remove_agents keep theirs."""
        s, sxy, d = o.poll_events()
        spawned_o.extend((int(a), (float(p[0]), float(p[1]))) for a, p in zip(s, sxy))
        destroyed_o.extend(sorted(int(v) for v in d) if step_events else [int(v) for v in d])

    drain(False)  # the lattice
    steps = 0
    for it in range(90):
        op = rng.choice(["step", "step", "step", "step", "add", "remove", "set_state", "burst"])
        live = o.read_state()["id"]
        if op == "step":
            if len(live):
                P.resync(g, o)
            g.step(R.Duration(*dt))
            o.step(*dt)
            steps += 1
            drain(True)
        elif op == "burst":  # several asynchronous steps, one sync: counts live on the device in between
            k = int(rng.integers(2, 6))
            for _ in range(k):
                g.step_async(R.Duration(*dt))
                o.step(*dt)
                drain(True)
            g.sync()
            g._dispatch_events()
            steps += k
        elif op == "add":
            k = int(rng.integers(1, 24))
            xy = np.stack([rng.uniform(12.0, 52.0, k), rng.uniform(44.0, 60.0, k)], axis=1)  # they walk < 6 m
            assert list(o.add_agents(xy, ohl_b, olp_b, 1.0)) == g.add_agents(xy, hl_b, lp_b, 1.0)
        elif op == "remove" and len(live) > 8:
            for a in rng.choice(live, size=int(rng.integers(1, 6)), replace=False):
                o.remove_agent(int(a))
                g.remove_agents(int(a))
        elif op == "set_state" and len(live) > 8:
            ids = np.sort(rng.choice(live, size=int(rng.integers(1, 10)), replace=False)).astype(np.uint64)
            so = o.read_state()
            sel = np.searchsorted(so["id"], ids)
            vx, vy = rng.uniform(-0.2, 0.2, len(ids)), rng.uniform(-0.2, 0.2, len(ids))
            o.set_state(ids, so["x"][sel], so["y"][sel], vx, vy)
            g.set_state(ids, so["x"][sel], so["y"][sel], vx, vy)
        drain(False)  # add / remove
        # (bursts free-run for a few steps: the Zanlungo lattice may differ in the last bits -- libm --, nothing else)
        _check(g, o, (seed, it, op))
    assert steps > 30 and len(spawned_o) > 10 and len(destroyed_o) > 0
    assert [a for a, _ in rec.added] == [a for a, _ in spawned_o]
    assert [p for _, p in rec.added] == [p for _, p in spawned_o]
    assert rec.removed == destroyed_o
