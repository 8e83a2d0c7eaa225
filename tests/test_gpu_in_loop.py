"""GPU: rcs_step_in_loop -- the reference's IN-LOOP index semantic (lib.rs:299: the index is updated
inside the per-agent loop) computed as a fixed-point iteration of whole-crowd sweeps, against the oracle running the
literal sequential loop in the same iteration order.  Neighbour lists and t_i bit-exact, forces / state 1e-9."""
import numpy as np
import pytest

import oracle_ffi as O
import parity as P
import rmf_crowdsim_b200 as R
from rmf_crowdsim_b200 import scenes as SC

pytestmark = pytest.mark.gpu


def _oracle(scene, order):
    o = O.OracleSim(scene.width, scene.height, scene.cell, scene.offset, index_mode=O.IN_LOOP,
                    iter_order=O.ASCENDING_ID if order is None else O.CUSTOM)
    kind, v = scene.hl
    hl = o.hl_parity(v) if kind == "parity" else o.hl_constant(v)
    ids = o.add_agents(scene.xy, hl, o.lp_zanlungo(*scene.lp[1:]), scene.eyesight)
    o.set_state(ids, scene.xy[:, 0], scene.xy[:, 1], scene.vxy[:, 0], scene.vxy[:, 1])
    if order is not None:
        o.set_custom_order(order)
    o.enable_trace(True)
    return o


@pytest.mark.parametrize("order_kind", ["ascending", "descending", "random"])
@pytest.mark.parametrize("cell,eyesight", [(2.0, 2.0), (1.0, 2.5)])
def test_in_loop_semantic_matches_the_sequential_oracle(order_kind, cell, eyesight):
    rng = np.random.default_rng(31)
    scene = SC.uniform_crowd(40, "shuffled", cell=cell, eyesight=eyesight, margin=8.0, seed=13,
                             lp=("zanlungo", 0.05, 1.0, 0.0, 0.5, 200.0, 0.1))
    scene.vxy = scene.vxy + rng.uniform(-0.3, 0.3, size=scene.vxy.shape)
    # a long step so that many agents change cell and many neighbour sets depend on who has moved already
    dt = (0, 250_000_000)
    n = scene.n
    order = {"ascending": None, "descending": np.arange(n, dtype=np.uint64)[::-1].copy(),
             "random": rng.permutation(n).astype(np.uint64)}[order_kind]
    g = SC.build_simulation(scene)
    g.set_trace(True)
    o = _oracle(scene, order)
    d = SC.build_simulation(scene)  # the deferred contract, for comparison
    d.set_trace(True)
    differs = 0
    for _ in range(2):
        P.resync(g, o)
        P.resync(d, o)
        sweeps = g.step_in_loop(R.Duration(*dt), order=order)
        o.step(*dt)
        d.step(R.Duration(*dt))
        assert 2 <= sweeps <= 64
        tg, to = g.read_trace(), o.read_trace()
        r = P.compare_traces(tg, to)  # neighbour lists and t_i bit-exact
        assert r["force_rel_err"] <= P.REL_TOL and r["finite_tti"] > 0
        s = P.compare_states(g.read_state(), o.read_state())
        assert s["vel_rel_err"] <= P.REL_TOL and s["pos_rel_err"] <= P.REL_TOL
        td = d.read_trace()
        differs += int(np.sum(np.diff(td["nb_offsets"].astype(np.int64)) != np.diff(tg["nb_offsets"].astype(np.int64))))
    assert differs > 0  # the in-loop semantic really is a different neighbourhood for some agents


def test_in_loop_out_of_bounds_is_the_reference_error_and_nothing_is_committed():
    scene = SC.uniform_crowd(12, "lane", margin=4.0, seed=3)
    scene.hl = ("constant", (60.0, 0.0))
    g = SC.build_simulation(scene)
    before = g.read_state()
    with pytest.raises(R.CrowdsimError) as e:
        g.step_in_loop(R.Duration(1, 0))
    assert "Index out of bounds" in str(e.value)
    after = g.read_state()
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(before[k].view(np.uint64), after[k].view(np.uint64))
