"""Shared helpers of the parity tests: build oracle and CUDA simulations from one Scene and
compare them step by step.  Tolerances are the ones BASELINE.json's north_star states:
cell assignments and neighbour sets bit-exact, t_i bit-exact (only IEEE +,-,*,/,sqrt feed it),
forces / velocities <= 1e-9 relative (CUDA's exp/asin/sin differ from glibc's by <= 2 ulp)."""
from __future__ import annotations

import numpy as np

import oracle_ffi as O
from rmf_crowdsim_b200 import scenes as SC

REL_TOL = 1e-9


def build_oracle(scene: SC.Scene, index_mode=O.DEFERRED, inject_velocity=True) -> O.OracleSim:
    o = O.OracleSim(scene.width, scene.height, scene.cell, scene.offset, index_mode=index_mode)
    kind, v = scene.hl
    hl = o.hl_parity(v) if kind == "parity" else o.hl_constant(v)
    lp = o.lp_none() if scene.lp[0] == "none" else o.lp_zanlungo(*scene.lp[1:])
    ids = o.add_agents(scene.xy, hl, lp, scene.eyesight)
    assert np.array_equal(ids, np.arange(scene.n, dtype=np.uint64))
    if inject_velocity and np.any(scene.vxy):
        o.set_state(ids, scene.xy[:, 0], scene.xy[:, 1], scene.vxy[:, 0], scene.vxy[:, 1])
    return o


def rel_err(a: np.ndarray, b: np.ndarray, scale=None) -> float:
    """max |a-b| / max(|a|,|b|,scale) with identical non-finite patterns required."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    assert np.array_equal(np.isinf(a), np.isinf(b)), "inf pattern differs"
    fin = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(a[~fin & ~np.isnan(a)], b[~fin & ~np.isnan(b)]), "inf signs differ"
    if not fin.any():
        return 0.0
    den = np.maximum(np.abs(a[fin]), np.abs(b[fin]))
    if scale is not None:
        den = np.maximum(den, np.broadcast_to(scale, a.shape)[fin] if np.ndim(scale) else scale)
    den = np.maximum(den, 1e-300)
    return float(np.max(np.abs(a[fin] - b[fin]) / den))


def compare_traces(tg: dict, to: dict) -> dict:
    """GPU trace vs oracle trace (both ascending id).  Returns summary numbers."""
    assert np.array_equal(tg["id"], to["id"])
    assert np.array_equal(tg["nb_offsets"], to["nb_offsets"]), "neighbour counts differ"
    assert np.array_equal(tg["nb_ids"], to["nb_ids"]), "neighbour lists differ (bit-exact, canonical order)"
    # t_i: bit-exact (compare the bit patterns so that inf == inf and -0 != +0 are both caught)
    assert np.array_equal(tg["t_i"].view(np.uint64), to["t_i"].view(np.uint64)), "t_i not bit-exact"
    fmag = np.sqrt(to["fx"] ** 2 + to["fy"] ** 2)
    fmag = np.where(np.isfinite(fmag), fmag, 0.0)
    ex = rel_err(tg["fx"], to["fx"], scale=fmag)
    ey = rel_err(tg["fy"], to["fy"], scale=fmag)
    return {"force_rel_err": max(ex, ey), "finite_tti": int(np.isfinite(to["t_i"]).sum()),
            "neighbours": int(to["nb_offsets"][-1])}


def csr_subset(t: dict, keep: np.ndarray) -> dict:
    """Restrict a trace (ascending id) to the agents selected by the boolean mask `keep`."""
    off = t["nb_offsets"].astype(np.int64)
    chunks = [t["nb_ids"][off[k]:off[k + 1]] for k in np.nonzero(keep)[0]]
    new_off = np.zeros(int(keep.sum()) + 1, dtype=np.uint64)
    if chunks:
        new_off[1:] = np.cumsum([len(c) for c in chunks])
    nb = np.concatenate(chunks) if chunks else t["nb_ids"][:0]
    return {"id": t["id"][keep], "t_i": t["t_i"][keep], "fx": t["fx"][keep], "fy": t["fy"][keep],
            "nb_offsets": new_off, "nb_ids": nb}


def compare_states(sg: dict, so: dict) -> dict:
    assert np.array_equal(sg["id"], so["id"])
    vmag = np.sqrt(so["vx"] ** 2 + so["vy"] ** 2)
    vmag = np.where(np.isfinite(vmag), vmag, 0.0)
    ev = max(rel_err(sg["vx"], so["vx"], scale=vmag), rel_err(sg["vy"], so["vy"], scale=vmag))
    ep = max(rel_err(sg["x"], so["x"]), rel_err(sg["y"], so["y"]))
    assert np.array_equal(sg["next_waypoint"].astype(np.uint64), so["next_waypoint"].astype(np.uint64))
    return {"vel_rel_err": ev, "pos_rel_err": ep}


def step_both(g, o, scene: SC.Scene):
    from rmf_crowdsim_b200 import Duration

    g.step(Duration(*scene.dt))
    o.step(*scene.dt)


def resync(g, o) -> None:
    """Make the CUDA simulation's state bit-identical to the oracle's (ascending-id arrays)."""
    so = o.read_state()
    if len(so["id"]):
        g.set_state(None, so["x"], so["y"], so["vx"], so["vy"])
