"""Generates the golden vectors in this directory from the CPU oracle (oracle/, the C++ restatement of the
reference; the Rust crate itself cannot be built in this environment -- DESIGN.md section 6).

    python tests/golden/make_golden.py

Files (all numpy .npz, a few hundred KB in total):
  c1_viz.npz            'three's a crowd' scene (rmf_crowdsim_viz/src/main.rs:64-94), dt = 16_666_667 ns:
                        full state at steps 1, 300, 301, 302, 600, 1000 and the traces of steps 301, 302
  crowd_576.npz         24 x 24 jittered lattice, shuffled ids, Zanlungo: input state, neighbour CSR, t_i,
                        force and output state of ONE step (rows A3-A10 of SURVEY.md section 8a)
  pair_table.npz        512 random (agent, other, t_i) triples -> compute_agent_force (zanlungo.rs:93-170) and
                        512 random (rel_vel, rel_pos) -> time_to_collision (zanlungo.rs:49-74)
  knn_radius_100.npz    the 10 x 10 point grid of location_hash_2d.rs:310-368: kNN and radius answers
  source_sink.npz       tests/event_listeners_test.rs scenario: agent count / spawned / destroyed per step
  in_loop_400.npz       20 x 20 crowd, ONE step of 0.25 s under the reference's in-loop index update (lib.rs:299) in a
                        fixed random iteration order: the order, neighbour CSR, t_i, force and output state
                        (SURVEY.md section 8f-4); differs from the deferred contract for about half of the agents
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_ffi as O  # noqa: E402
import parity as P  # noqa: E402
from rmf_crowdsim_b200 import scenes as SC  # noqa: E402


def c1():
    scene = SC.config_c1()
    o = P.build_oracle(scene)
    o.enable_trace(True)
    out = {}
    for step in range(1, 1001):
        o.step(*scene.dt)
        if step in (1, 300, 301, 302, 600, 1000):
            st = o.read_state()
            for k in ("x", "y", "vx", "vy"):
                out[f"s{step}_{k}"] = st[k]
        if step in (301, 302):
            tr = o.read_trace()
            for k in ("t_i", "fx", "fy", "nb_offsets", "nb_ids"):
                out[f"t{step}_{k}"] = tr[k]
    np.savez(os.path.join(HERE, "c1_viz.npz"), **out)


def crowd():
    scene = SC.uniform_crowd(24, "shuffled", margin=8.0, seed=17)
    o = P.build_oracle(scene)
    o.enable_trace(True)
    cells = o.cell_of(scene.xy)
    o.step(*scene.dt)
    tr, st = o.read_trace(), o.read_state()
    np.savez(os.path.join(HERE, "crowd_576.npz"), in_xy=scene.xy, in_vxy=scene.vxy, cells=cells,
             nb_offsets=tr["nb_offsets"], nb_ids=tr["nb_ids"], t_i=tr["t_i"], fx=tr["fx"], fy=tr["fy"],
             x=st["x"], y=st["y"], vx=st["vx"], vy=st["vy"])


def pair_table():
    import ctypes as C

    rng = np.random.default_rng(2024)
    L = O.lib()
    n = 512
    params = np.array([0.05, 1.0, 0.0, 0.5, 1.0, 0.2])
    a = np.concatenate([rng.uniform(-2, 2, (n, 2)), rng.uniform(-1.5, 1.5, (n, 2)), rng.uniform(-1.5, 1.5, (n, 2))], 1)
    ob = np.concatenate([a[:, :2] + rng.uniform(-2, 2, (n, 2)), rng.uniform(-1.5, 1.5, (n, 2)), np.zeros((n, 2))], 1)
    aid = rng.integers(1000, 2000, n).astype(np.uint64)
    delta = rng.integers(1, 900, n).astype(np.int64) * np.where(rng.random(n) < 0.5, 1, -1)
    oid = (aid.astype(np.int64) + delta).astype(np.uint64)  # half of the pairs have the higher id (weight 0)
    t_i = rng.uniform(0.05, 6.0, n)
    force = np.zeros((n, 2))
    f64p = C.POINTER(C.c_double)
    for k in range(n):
        out = np.zeros(2)
        L.orc_agent_force(params.ctypes.data_as(f64p), int(aid[k]), np.ascontiguousarray(a[k]).ctypes.data_as(f64p),
                          int(oid[k]), np.ascontiguousarray(ob[k]).ctypes.data_as(f64p), float(t_i[k]),
                          out.ctypes.data_as(f64p))
        force[k] = out
    rv = rng.uniform(-3, 3, (n, 2))
    rp = rng.uniform(-3, 3, (n, 2))
    aim = slice(n // 2, n)  # half of the table is aimed at the other agent (finite roots, both branches)
    rv[aim] = -rp[aim] * rng.uniform(-0.5, 2.0, (n // 2, 1)) + rng.normal(0, 0.05, (n // 2, 2))
    rp[-16:] *= 0.05          # already overlapping: roots of opposite sign -> 0.0
    rv[:8] = 0.0              # a == 0 -> NaN roots -> inf
    ttc = np.array([O.ttc(0.2, rv[k], rp[k]) for k in range(n)])
    np.savez(os.path.join(HERE, "pair_table.npz"), params=params, agent=a, other=ob, aid=aid, oid=oid, t_i=t_i,
             force=force, rel_vel=rv, rel_pos=rp, ttc=ttc, ttc_radius=np.array([0.2]))


def knn_radius():
    o = O.OracleSim(10.0, 10.0, 0.5, (0.0, 0.0))
    for x in range(10):
        for y in range(10):
            o.index_add_or_update(10 * x + y, (x + 0.5, y + 0.5))
    rng = np.random.default_rng(5)
    q = np.concatenate([np.array([[0.6, 0.6], [1.7, 1.6], [4.0, 4.0]]), rng.uniform(-1, 11, (61, 2))])
    knn4 = np.full((len(q), 4), np.iinfo(np.uint64).max, dtype=np.uint64)
    cnt = np.zeros(len(q), dtype=np.uint64)
    rad_off = np.zeros(len(q) + 1, dtype=np.uint64)
    rad = []
    for k, p in enumerate(q):
        r = o.query_knn(4, p)
        knn4[k, :len(r)] = r
        cnt[k] = len(r)
        rr = o.query_radius(1.1, p)
        rad.append(rr)
        rad_off[k + 1] = rad_off[k] + np.uint64(len(rr))
    np.savez(os.path.join(HERE, "knn_radius_100.npz"), q=q, knn4=knn4, knn4_count=cnt, radius=np.array([1.1]),
             rad_offsets=rad_off, rad_ids=np.concatenate(rad))


def source_sink():
    o = O.OracleSim(1000, 1000, 20, (-500, -500))
    o.add_source_sink((0, 0), 1.0, 1.0, o.hl_constant((1, 0)), o.lp_none(), [(20, 0)], False, 5.0)
    count, spawned, destroyed = [], [], []
    for _ in range(40):
        o.step(1, 0)
        s, _, d = o.poll_events()
        count.append(o.agent_count())
        spawned.append(len(s))
        destroyed.append(len(d))
    st = o.read_state()
    np.savez(os.path.join(HERE, "source_sink.npz"), count=np.array(count), spawned=np.array(spawned),
             destroyed=np.array(destroyed), final_id=st["id"], final_x=st["x"])


def in_loop_scene():
    rng = np.random.default_rng(77)
    scene = SC.uniform_crowd(20, "shuffled", margin=8.0, seed=23, lp=("zanlungo", 0.05, 1.0, 0.0, 0.5, 200.0, 0.1))
    scene.vxy = scene.vxy + rng.uniform(-0.3, 0.3, size=scene.vxy.shape)
    order = rng.permutation(scene.n).astype(np.uint64)
    return scene, order, (0, 250_000_000)


def in_loop():
    scene, order, dt = in_loop_scene()
    o = O.OracleSim(scene.width, scene.height, scene.cell, scene.offset, index_mode=O.IN_LOOP, iter_order=O.CUSTOM)
    ids = o.add_agents(scene.xy, o.hl_parity(scene.hl[1]), o.lp_zanlungo(*scene.lp[1:]), scene.eyesight)
    o.set_state(ids, scene.xy[:, 0], scene.xy[:, 1], scene.vxy[:, 0], scene.vxy[:, 1])
    o.set_custom_order(order)
    o.enable_trace(True)
    o.step(*dt)
    tr, st = o.read_trace(), o.read_state()
    d = P.build_oracle(scene)  # the deferred contract on the same input, for the record
    d.enable_trace(True)
    d.step(*dt)
    td = d.read_trace()
    differs = int(np.sum(np.diff(td["nb_offsets"].astype(np.int64)) != np.diff(tr["nb_offsets"].astype(np.int64))))
    np.savez(os.path.join(HERE, "in_loop_400.npz"), in_xy=scene.xy, in_vxy=scene.vxy, order=order,
             dt=np.array(dt, dtype=np.uint64), nb_offsets=tr["nb_offsets"], nb_ids=tr["nb_ids"], t_i=tr["t_i"],
             fx=tr["fx"], fy=tr["fy"], x=st["x"], y=st["y"], vx=st["vx"], vy=st["vy"],
             agents_with_other_neighbour_count_than_deferred=np.array([differs]))


if __name__ == "__main__":
    in_loop()
    c1()
    crowd()
    pair_table()
    knn_radius()
    source_sink()
    print("golden vectors written to", HERE)
