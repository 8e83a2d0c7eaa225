"""GPU: seeded random configurations of the hot path against the oracle AND against the thread-per-agent kernel.
Every case draws its own hash grid (cell size, offset), crowd (density, extent, a hole, a dense blob), two or three
agent groups (eyesight, Zanlungo constants, high-level planner, one NoLocalPlan group) and state, so that one step
mixes all the routes an agent can take through the kernels: the three-slice cooperative path, the chunked one
(eyesight > cell, crowded columns), agents left of / below the grid origin, empty stencils.
Bar: neighbour lists and t_i bit-exact against the oracle, forces / velocities 1e-9; the two kernel forms agree
bit for bit."""
import numpy as np
import pytest

import oracle_ffi as O
import parity as P
import rmf_crowdsim_b200 as R
from rmf_crowdsim_b200 import _native as N

pytestmark = pytest.mark.gpu


def _draw_case(seed):
    rng = np.random.default_rng(1000 + seed)
    cell = float(rng.choice([0.5, 0.75, 1.0, 2.0, 3.0, 5.0]))
    span = float(rng.choice([24.0, 36.0, 48.0, 72.0]))
    off = (float(rng.uniform(-10, 10)), float(rng.uniform(-10, 10)))
    n_cells = int(np.ceil((span + 12.0) / cell))
    width = height = n_cells * cell          # square grids: the reference's index formula (location_hash_2d.rs:59)
    spacing = float(rng.choice([0.6, 0.8, 1.0, 1.5]))
    side = int(span / spacing)
    gi, gj = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    xy = np.stack([gi.reshape(-1), gj.reshape(-1)], axis=1) * spacing
    xy = xy + rng.uniform(-0.25, 0.25, size=xy.shape) * spacing
    # a hole and a denser blob (finer lattice, same minimum distance rule: jitter < spacing / 4)
    c = rng.uniform(0.3, 0.7, size=2) * span
    xy = xy[np.linalg.norm(xy - c, axis=1) > 0.12 * span]
    b = rng.uniform(0.2, 0.8, size=2) * span
    bi, bj = np.meshgrid(np.arange(10), np.arange(10), indexing="ij")
    blob = b + np.stack([bi.reshape(-1), bj.reshape(-1)], axis=1) * (spacing / 3) + rng.uniform(-0.02, 0.02, (100, 2))
    xy = np.concatenate([xy[np.linalg.norm(xy - (b + 1.5 * spacing), axis=1) > 3.0 * spacing], blob])
    # the crowd starts 3 m left of / below the grid origin: those agents are filed in row / column 0
    xy = xy + np.array(off) - 3.0
    xy = xy[rng.permutation(len(xy))]
    n = len(xy)
    cut = np.sort(rng.choice(np.arange(1, n), size=2, replace=False))
    groups = []
    for g in range(3):
        eyesight = float(rng.choice([0.9, 1.6, 2.0, 2.7, 4.1]))
        radius = float(rng.choice([0.02, 0.04]))
        mass = float(rng.choice([20.0, 80.0]))     # heavy agents: dense blobs stay sane for the two steps
        zan = (float(rng.uniform(0.02, 0.3)), 1.0, 0.0, float(rng.uniform(0.3, 1.0)), mass, radius)
        hl = ("parity", (float(rng.uniform(0.5, 1.5)), float(rng.uniform(-0.5, 0.5)))) if rng.random() < 0.6 else \
             ("constant", (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1))))
        groups.append({"eyesight": eyesight, "zan": zan if g < 2 else None, "hl": hl})
    vxy = rng.uniform(-1.2, 1.2, size=(n, 2))
    vxy[rng.random(n) < 0.1] = 0.0               # a == 0 pairs
    return {"cell": cell, "width": width, "height": height, "off": off, "xy": xy, "vxy": vxy, "cut": cut,
            "groups": groups, "dt": (0, int(rng.choice([5_000_000, 10_000_000])))}


def _build_gpu(case, kernel):
    idx = R.LocationHash2D(case["width"], case["height"], case["cell"], case["off"], capacity=len(case["xy"]))
    g = R.Simulation(idx)
    g.set_option(N.RCS_OPT_STEP_KERNEL, kernel)
    lo = 0
    for grp, hi in zip(case["groups"], list(case["cut"]) + [len(case["xy"])]):
        hl = R.ParityVelocityPlan(grp["hl"][1]) if grp["hl"][0] == "parity" else R.ConstantVelocityPlan(grp["hl"][1])
        lp = R.Zanlungo(*grp["zan"]) if grp["zan"] else R.NoLocalPlan()
        g.add_agents(case["xy"][lo:hi], hl, lp, grp["eyesight"])
        lo = hi
    ids = np.arange(len(case["xy"]), dtype=np.uint64)
    g.set_state(ids, case["xy"][:, 0], case["xy"][:, 1], case["vxy"][:, 0], case["vxy"][:, 1])
    g.set_trace(True)
    return g


def _build_oracle(case):
    o = O.OracleSim(case["width"], case["height"], case["cell"], case["off"])
    lo = 0
    for grp, hi in zip(case["groups"], list(case["cut"]) + [len(case["xy"])]):
        hl = o.hl_parity(grp["hl"][1]) if grp["hl"][0] == "parity" else o.hl_constant(grp["hl"][1])
        lp = o.lp_zanlungo(*grp["zan"]) if grp["zan"] else o.lp_none()
        o.add_agents(case["xy"][lo:hi], hl, lp, grp["eyesight"])
        lo = hi
    ids = np.arange(len(case["xy"]), dtype=np.uint64)
    o.set_state(ids, case["xy"][:, 0], case["xy"][:, 1], case["vxy"][:, 0], case["vxy"][:, 1])
    o.enable_trace(True)
    return o


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_matches_the_oracle_and_both_kernel_forms_agree(seed):
    case = _draw_case(seed)
    ga, gb, o = _build_gpu(case, 0), _build_gpu(case, 1), _build_oracle(case)
    zan_ids = np.arange(len(case["xy"])) < case["cut"][1]   # the third group has no local planner (no trace)
    finite = 0
    for _ in range(2):
        P.resync(ga, o)
        P.resync(gb, o)
        for g in (ga, gb):
            g.step(R.Duration(*case["dt"]))
        o.step(*case["dt"])
        ta, tb, to = ga.read_trace(), gb.read_trace(), o.read_trace()
        for k in ("id", "nb_offsets", "nb_ids"):
            assert np.array_equal(ta[k], tb[k]), k
        for k in ("t_i", "fx", "fy"):
            assert np.array_equal(ta[k].view(np.uint64), tb[k].view(np.uint64)), k
        keep = zan_ids[ta["id"].astype(np.int64)]
        keep_o = zan_ids[to["id"].astype(np.int64)]
        r = P.compare_traces(P.csr_subset(ta, keep), P.csr_subset(to, keep_o))
        assert r["force_rel_err"] <= P.REL_TOL
        finite += r["finite_tti"]
        sa, sb, so = ga.read_state(), gb.read_state(), o.read_state()
        for k in ("x", "y", "vx", "vy"):
            assert np.array_equal(sa[k].view(np.uint64), sb[k].view(np.uint64)), k
        s = P.compare_states(sa, so)
        assert s["vel_rel_err"] <= P.REL_TOL and s["pos_rel_err"] <= P.REL_TOL
    assert finite > 0
    assert ga.stats().candidate_total == gb.stats().candidate_total
