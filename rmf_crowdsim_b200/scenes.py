"""Synthetic crowds for the benchmark configurations of SURVEY.md section 8(d) (BASELINE.json configs).

"Uniform" crowds are jittered lattices: site (i+0.5, j+0.5)*s + U(-s/4, s/4)^2, so density is uniform but
no pair starts closer than s/2 (pure uniform-random placement starts with agents inside each other's
agent_radius, which the reference model turns into 1e15 forces / NaN at the second step).  The random
generator is numpy's PCG64 with the seed recorded in the scene.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

import numpy as np


@dataclass
class Scene:
    name: str
    # LocationHash2D::new arguments
    width: float
    height: float
    cell: float
    offset: Tuple[float, float]
    # agents in id order (row k = agent id k)
    xy: np.ndarray
    vxy: np.ndarray  # initial velocities injected after add_agents (the reference starts at 0)
    eyesight: float
    # planners
    hl: Tuple[str, Tuple[float, float]] = ("parity", (1.3, 0.0))
    lp: Tuple = ("zanlungo", 0.05, 1.0, 0.0, 0.5, 1.0, 0.2)
    dt: Tuple[int, int] = (0, 16_666_667)
    seed: int = 1
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return int(self.xy.shape[0])


def jittered_lattice(nx: int, ny: int, s: float, seed: int) -> np.ndarray:
    """(nx*ny, 2) positions, row-major over (i, j) with j fastest; site (i, j) -> row i*ny + j."""
    rng = np.random.Generator(np.random.PCG64(seed))
    i, j = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64), indexing="ij")
    jit = rng.uniform(-s / 4.0, s / 4.0, size=(nx, ny, 2))
    x = (i + 0.5) * s + jit[..., 0]
    y = (j + 0.5) * s + jit[..., 1]
    return np.stack([x.reshape(-1), y.reshape(-1)], axis=1)


def _assign_ids(nx: int, ny: int, variant: str, seed: int) -> np.ndarray:
    """site index -> agent id.  'shuffled': seeded permutation.  'lane': id parity == lattice row (j)
    parity, so every lane of constant y moves one way under the parity planner."""
    n = nx * ny
    if variant == "shuffled":
        rng = np.random.Generator(np.random.PCG64(seed + 0x9E3779B9))
        return rng.permutation(n).astype(np.uint64)
    if variant == "lane":
        if ny % 2:
            raise ValueError("lane ordering needs an even number of rows")
        j = np.tile(np.arange(ny), nx)
        ids = np.zeros(n, dtype=np.uint64)
        for par in (0, 1):
            sel = np.nonzero(j % 2 == par)[0]
            ids[sel] = 2 * np.arange(len(sel), dtype=np.uint64) + par
        return ids
    raise ValueError(variant)


def uniform_crowd(side: int, variant: str = "shuffled", s: float = 1.0, cell: float = 2.0, eyesight: float = 2.0,
                  margin: float = 64.0, speed: float = 1.3, seed: int = 1, lp=None, name: str = "") -> Scene:
    """side x side agents at spacing s in a square spawn box; square hash domain with a margin
    (the reference's index formula is only sane for n_x == n_y, location_hash_2d.rs:59)."""
    n = side * side
    site_xy = jittered_lattice(side, side, s, seed)
    ids = _assign_ids(side, side, variant, seed)
    xy = np.zeros((n, 2))
    xy[ids.astype(np.int64)] = site_xy
    par = (np.arange(n) % 2).astype(np.float64)  # even id -> -v, odd id -> +v (viz main.rs:26-29)
    vxy = np.zeros((n, 2))
    vxy[:, 0] = np.where(par == 0, -speed, speed)
    box = side * s
    dom = box + 2 * margin
    dom = float(np.ceil(dom / cell) * cell)
    return Scene(
        name=name or f"uniform_{n}_{variant}",
        width=dom, height=dom, cell=cell, offset=(-margin, -margin),
        xy=xy, vxy=vxy, eyesight=eyesight,
        hl=("parity", (speed, 0.0)),
        lp=lp or ("zanlungo", 0.05, 1.0, 0.0, 0.5, 1.0, 0.2),
        seed=seed,
        meta={"variant": variant, "spacing": s, "side": side, "speed": speed},
    )


def config_c1() -> Scene:
    """rmf_crowdsim_viz 'three's a crowd' scene (main.rs:64-94) at dt = 16_666_667 ns."""
    xy = np.array([[100.0, 100.0], [100.0, -100.0], [60.0, 100.0]])
    return Scene("c1_viz", 1000.0, 1000.0, 20.0, (-500.0, -500.0), xy, np.zeros((3, 2)), 100.0,
                 hl=("parity", (0.0, 10.0)), lp=("zanlungo", 1.0, 1.0, 0.0, 40.0, 2.0, 20.0))


def config_c2(variant: str = "shuffled", seed: int = 1) -> Scene:
    """10k agents, 100 m x 100 m spawn box, density 1/m^2."""
    return uniform_crowd(100, variant, margin=32.0, seed=seed, name=f"c2_10k_{variant}")


def config_c2_sparse(seed: int = 1) -> Scene:
    """C2-sparse: 5 m spacing, R = cell = 5 m, for the long drift run."""
    return uniform_crowd(100, "shuffled", s=5.0, cell=5.0, eyesight=5.0, margin=40.0, seed=seed,
                         lp=("zanlungo", 0.1, 1.0, 0.0, 0.4, 1.0, 0.2), name="c2_sparse_10k")


def config_c3(variant: str = "shuffled", seed: int = 1, lp=None) -> Scene:
    """2^20 agents, 1024 m x 1024 m spawn box, domain 1152^2."""
    return uniform_crowd(1024, variant, margin=64.0, seed=seed, lp=lp, name=f"c3_1m_{variant}")


def config_c4(variant: str = "shuffled", seed: int = 1) -> Scene:
    """2^24 agents, 4096 m x 4096 m, bidirectional +-x flow by id parity."""
    return uniform_crowd(4096, variant, margin=64.0, seed=seed, name=f"c4_16m_{variant}")


def build_simulation(scene: Scene, device: int = 0, capacity: int | None = None, inject_velocity: bool = True):
    """Scene -> rmf_crowdsim_b200.Simulation (GPU)."""
    from . import sim as S

    idx = S.LocationHash2D(scene.width, scene.height, scene.cell, scene.offset,
                           capacity=capacity or max(scene.n, 16), device=device)
    simu = S.Simulation(idx)
    kind, v = scene.hl
    hl = {"parity": S.ParityVelocityPlan, "constant": S.ConstantVelocityPlan}[kind](v)
    lp = S.NoLocalPlan() if scene.lp[0] == "none" else S.Zanlungo(*scene.lp[1:])
    # keep the planner objects alive with the simulation (handles are keyed by object identity)
    simu._scene_planners = (hl, lp)
    add_agents_bulk(simu, scene.xy, hl, lp, scene.eyesight)
    if inject_velocity and np.any(scene.vxy):
        simu.set_state(None, vx=scene.vxy[:, 0].copy(), vy=scene.vxy[:, 1].copy())
    return simu


def add_agents_bulk(simu, xy: np.ndarray, hl, lp, eyesight: float) -> np.ndarray:
    """Simulation::add_agents without building Python lists of ids (large crowds)."""
    import ctypes as C

    from . import _native as N

    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    ids = np.zeros(xy.shape[0], dtype=np.uint64)
    N.check(simu._h, simu._lib.rcs_add_agents(simu._h, xy.shape[0], xy.ctypes.data_as(N.c_f64p), simu._hl(hl),
                                              simu._lp(lp), float(eyesight), ids.ctypes.data_as(N.c_u64p)))
    return ids
