"""GPU: the reference's own tests, restated against the CUDA path through the C ABI
(lib.rs:422-453, location_hash_2d.rs:342-397, zanlungo.rs:224-236)."""
import numpy as np
import pytest

import rmf_crowdsim_b200 as R

pytestmark = pytest.mark.gpu


def _grid100(h):
    pts = {}
    i = 0
    for x in range(10):
        for y in range(10):
            p = (x + 0.5, y + 0.5)
            h.add_or_update(i, p)
            pts[i] = p
            i += 1
    return pts


def test_radius_search():
    h = R.LocationHash2D(10, 10, 0.5, (0, 0), capacity=128)
    pts = _grid100(h)
    gt = {i for i, p in pts.items() if np.sqrt((p[0] - 4) ** 2 + (p[1] - 4) ** 2) < 1.1}
    assert set(h.get_neighbours_in_radius(1.1, (4, 4))) == gt == {33, 34, 43, 44}


def test_update():
    h = R.LocationHash2D(2, 2, 1, (0, 0), capacity=8)
    h.add_or_update(1, (0, 0))
    assert h.get_neighbours_in_radius(1.0, (0, 0)) == [1]
    h.add_or_update(1, (1, 0))
    assert h.get_neighbours_in_radius(1.0, (0, 0)) == []


def test_remove():
    h = R.LocationHash2D(1, 1, 1, (0, 0), capacity=8)
    h.add_or_update(1, (0, 0))
    assert len(h.get_neighbours_in_radius(1.1, (0, 0))) == 1
    h.remove_agent(1)
    assert len(h.get_neighbours_in_radius(1.1, (0, 0))) == 0
    h.remove_agent(1)  # unknown id: no-op, as the reference


def test_add_out_of_bounds_is_the_reference_error():
    h = R.LocationHash2D(2, 2, 1, (0, 0), capacity=8)
    with pytest.raises(R.CrowdsimError) as e:
        h.add_or_update(3, (5.0, 0.0))
    assert str(e.value) == "Index out of bounds"


def test_step_integration():
    velocity = (1.0, 0.0)
    sim = R.Simulation(R.LocationHash2D(1000, 1000, 20, (-500, -500), capacity=16))
    assert len(sim.agents) == 0
    ids = sim.add_agents([(0.0, 0.0)], R.ConstantVelocityPlan(velocity), R.NoLocalPlan(), 100.0)
    assert ids == [0] and len(sim.agents) == 1
    sim.step(R.Duration(1, 0))
    agents = sim.agents
    assert len(agents) == 1
    p = agents[0].position
    assert np.hypot(p[0] - velocity[0], p[1] - velocity[1]) < 1e-5
    assert p == (1.0, 0.0) and agents[0].velocity == (1.0, 0.0)


def _tti_of_pair(rel_vel, rel_pos):
    """t_i of agent 0 (at rest at the origin) seeing agent 1 at rel_pos moving with rel_vel,
    Zanlungo::new(1, 10, 0, 5, 0.1, 4) as in zanlungo.rs:226."""
    sim = R.Simulation(R.LocationHash2D(100, 100, 20, (-50, -50), capacity=16))
    sim.add_agents([(0.0, 0.0), rel_pos], R.NoHighLevelPlan(), R.Zanlungo(1, 10, 0, 5, 0.1, 4), 30.0)
    sim.set_state([0, 1], vx=[0.0, rel_vel[0]], vy=[0.0, rel_vel[1]])
    sim.set_trace(True)
    sim.step(R.Duration(0, 1000))
    tr = sim.read_trace()
    assert list(tr["id"]) == [0, 1]
    assert list(tr["nb_ids"]) == [1, 0]
    return tr["t_i"][0]


def test_time_to_collision_head_on():
    assert _tti_of_pair((1.0, 0.0), (-10.0, 0.0)) == 6.0


def test_time_to_collision_never_collide():
    assert _tti_of_pair((1.0, 0.0), (10.0, 0.0)) == float("inf")


def test_out_of_bounds_step_is_not_committed():
    sim = R.Simulation(R.LocationHash2D(10, 10, 1, (0, 0), capacity=16))
    sim.add_agents([(9.5, 5.0), (1.0, 1.0)], R.ConstantVelocityPlan((1.0, 0.0)), R.NoLocalPlan(), 1.0)
    with pytest.raises(R.CrowdsimError) as e:
        sim.step(R.Duration(1, 0))
    assert str(e.value) == "Index out of bounds"
    a = sim.agents
    assert a[0].position == (9.5, 5.0) and a[1].position == (1.0, 1.0)
    assert sim.stats().first_oob_id == 0
    sim.step(R.Duration(0, 250_000_000))  # still usable afterwards
    assert sim.agents[0].position == (9.75, 5.0)
