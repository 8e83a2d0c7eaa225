"""bench.py --gpus N (N > 1): one process per GPU (torchrun), spatial strips with NCCL halo exchange.

torch.distributed is plumbing only: it carries the 128-byte NCCL id from rank 0 to the others and the
max-over-ranks of the device-timed region.  The halo exchange itself is ncclSend/ncclRecv issued by
librcs.so on the simulation's own stream (rmf_crowdsim_b200/csrc/rcs_host_dist.inl).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np


def strip_scene(workload: str, variant: str, rank: int, world: int, lp_none: bool):
    """The agents of this rank's strip only (the crowd generator is counter-based, scenes.py): returns
    (scene-without-agents, ids, xy, vxy) for lattice columns that can fall into the strip."""
    from . import scenes as SC

    side = {"c4": 4096, "c3": 1024, "c2": 100}.get(workload)
    if side is None and workload.startswith("side"):
        side = int(workload[4:])
    margin = 32.0 if workload == "c2" else 64.0
    s, cell, speed, seed, eyesight = 1.0, 2.0, 1.3, 1, 2.0
    zan = ("zanlungo", 0.05, 1.0, 0.0, 0.5, 1.0, 0.2)
    if workload.startswith("sparse"):
        # SURVEY.md 8d "C2-sparse": 5 m spacing, R = cell = 5 m, agent_scale 0.1, force_distance 0.4 -- a crowd whose
        # force pass is busy and that stays finite for some tens of committed steps
        side = int(workload[6:])
        s, cell, eyesight, margin = 5.0, 5.0, 5.0, 40.0
        zan = ("zanlungo", 0.1, 1.0, 0.0, 0.4, 1.0, 0.2)
    dom = float(np.ceil((side * s + 2 * margin) / cell) * cell)
    ncols = int(dom / cell)
    # strips balanced by agent count: the crowd fills [0, side * s), the margins are empty
    bounds = [0] + [int(round((k * side * s / world + margin) / cell)) for k in range(1, world)] + [ncols]
    c0, c1 = bounds[rank], bounds[rank + 1]
    # lattice column i sits at x in (i*s, (i+1)*s): cell column floor((x + margin) / cell)
    i0 = max(0, int(np.floor((c0 * cell - margin) / s)) - 1)
    i1 = min(side, int(np.ceil((c1 * cell - margin) / s)) + 1)
    xy = SC.jittered_lattice(side, side, s, seed, i0, i1)
    ids = SC.site_ids(side, side, variant, seed, i0, i1)
    par = (ids % np.uint64(2)).astype(np.float64)
    vxy = np.zeros_like(xy)
    vxy[:, 0] = np.where(par == 0, -speed, speed)
    scene = SC.Scene(name=f"{workload}_{variant}_strip{rank}", width=dom, height=dom, cell=cell,
                     offset=(-margin, -margin), xy=np.zeros((0, 2)), vxy=np.zeros((0, 2)), eyesight=eyesight,
                     hl=("parity", (speed, 0.0)), lp=("none",) if lp_none else zan, seed=seed)
    scene.meta["bounds"] = bounds
    return scene, side * side, ids, xy, vxy


def run(args) -> None:
    import torch
    import torch.distributed as dist

    from . import _native as N
    from . import sim as S
    from .strips import StripSimulation, owned_mask

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # Before anything is timed: the multi-process exchange must reproduce one handle's bits on a crowd whose force pass
    # is busy and whose agents migrate between the ranks (committed steps).
    dist_check = None
    if not getattr(args, "skip_verify", False):
        w, v, d, k = DIST_VERIFY
        dist_check = verify_core(dist, torch, rank, world, local, w, v, S.Duration(*d), k)
    nccl_id = fresh_nccl_id(dist, torch, rank)

    workload = args.workload or "c4"
    frozen = not (args.variant == "lane" and not args.no_local_plan)
    scene, n_total, ids, xy, vxy = strip_scene(workload, args.variant, rank, world, args.no_local_plan)
    bounds = scene.meta["bounds"]
    c0, c1 = bounds[rank], bounds[rank + 1]
    m = owned_mask(xy[:, 0], scene.offset[0], scene.cell, c0, c1)
    n_own = int(m.sum())
    # room for the owned agents, three ghost columns per side and the churn of a long committed run
    per_col = int(round(n_total ** 0.5)) * scene.cell  # agents per cell column at spacing 1 m
    halo_cap = int(3.3 * per_col) + 2048  # three columns per side (W = ring + reach) and 10 % slack
    cap = int(n_own * 1.05) + 2 * halo_cap + 4096
    idx = S.LocationHash2D(scene.width, scene.height, scene.cell, scene.offset, capacity=cap, device=local)
    sim = StripSimulation(idx, rank, world, nccl_id, halo_capacity=halo_cap, boundaries=bounds,
                          peer_gather=peer_gather(dist, torch))
    assert (sim.c0, sim.c1) == (c0, c1)
    sim.add_scene_agents(scene, ids, xy, vxy)
    lib, h = sim._lib, sim._h
    dt = S.Duration(*scene.dt)
    flags = N.RCS_STEP_NO_COMMIT if frozen else N.RCS_STEP_DEFAULT

    def one_step():
        N.check(h, lib.rcs_step_async(h, dt.secs, dt.nanos, flags))

    from bench import ClockSampler  # the launcher module (repo root is on sys.path)

    for _ in range(max(args.warmup, 3)):
        one_step()
    sim.sync()
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(4):  # a little load before the timed region so that the clock samples are meaningful
        for _ in range(8):
            one_step()
        sim.sync()
    # After a sync a strip handle compacts its arrays (new buffer roles and counts), so its first steps run kernel by
    # kernel until their launch sequence repeats and is captured as a CUDA graph: eight more untimed steps, then only
    # the device is synchronised (rcs_sync would compact again) and the timed steps replay the graphs.
    for _ in range(8):
        one_step()
    launches0 = sim.launch_count()
    g0 = sim.graph_stats()
    K = args.steps
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    sim.event_record(0)
    for _ in range(K):
        one_step()
    sim.event_record(1)
    sim.sync()
    torch.cuda.synchronize()
    dist.barrier()
    wall1 = time.perf_counter()
    ms = torch.tensor([sim.event_elapsed_ms(0, 1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = sim.launch_count() - launches0
    g1 = sim.graph_stats()
    # the dominant kernel, bracketed by its own event pair: a second short pass (such steps run kernel by kernel)
    N.check(h, lib.rcs_kernel_timing(h, 1))
    for _ in range(min(K, 10)):
        one_step()
    sim.sync()
    kt_ms, kt_n = C.c_double(), C.c_uint64()
    N.check(h, lib.rcs_kernel_time_ms(h, C.byref(kt_ms), C.byref(kt_n)))
    N.check(h, lib.rcs_kernel_timing(h, 0))
    st = sim.stats()
    clk = clocks.stop()

    # per-rank statistics -> global
    agg = torch.tensor([st.neighbour_total, st.candidate_total, st.finite_tti_count, st.nonfinite_count,
                        st.oob_count, sim.agent_count(), launches], dtype=torch.float64, device="cuda")
    dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    kmax = torch.tensor([kt_ms.value / max(kt_n.value, 1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(kmax, op=dist.ReduceOp.MAX)

    e2e = None
    if not args.skip_e2e:
        sim.spatial_index.close()
        scene.hl = ("host", scene.hl[1])
        idx2 = S.LocationHash2D(scene.width, scene.height, scene.cell, scene.offset, capacity=cap, device=local)
        sim2 = StripSimulation(idx2, rank, world, fresh_nccl_id(dist, torch, rank), halo_capacity=halo_cap,
                               boundaries=bounds, peer_gather=peer_gather(dist, torch))
        sim2.add_scene_agents(scene, ids, xy, vxy)
        e2e = run_e2e(sim2, scene, min(K, 10), 3, dist, torch)

    if rank == 0:
        from bench import ALGO_BYTES_NOLOCALPLAN, ALGO_BYTES_ZANLUNGO, kernel_counts, measured_peaks, workload_name

        peaks, how = measured_peaks()
        algo = ALGO_BYTES_NOLOCALPLAN if args.no_local_plan else ALGO_BYTES_ZANLUNGO
        n_live = int(agg[5].item())
        value = n_live * K / (total_ms * 1e-3)
        k_ms = float(kmax.item())
        achieved = algo * (n_live / world) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        # DRAM bytes of one rank's kernel launch: the single-GPU capture of the same workload, per agent, times the
        # rank's agents (the strip form of the kernel moves 8 B per agent more: cell read, keep flag written)
        counts = kernel_counts(f"{workload}_{args.variant}" + ("_nolp" if args.no_local_plan else ""))
        traffic = None
        if counts and counts.get("kernel_dram_bytes") and counts.get("agents"):
            traffic = (counts["kernel_dram_bytes"] / counts["agents"] + 8.0) * (n_live / world)
        line = {
            "metric": "agent-steps/sec (query+Zanlungo+integrate)", "value": value, "unit": "agent-steps/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": total_ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(workload, args.variant) + (", NoLocalPlan" if args.no_local_plan else ""),
                "agents": n_live, "agents_per_gpu": n_live / world,
                "strip_boundaries": "balanced by agent count (rcs_dist_set_boundaries)",
                "parallelism": f"{world} spatial strips along x, halo of 3 cell columns per side by {sim.transport} "
                               + ("(the pack pass stores into the neighbour's memory over NVLink, a release / "
                                  "acquire round number orders it; the whole step replays as one CUDA graph per rank)"
                                  if sim.transport != "nccl" else
                                  "(ncclSend / ncclRecv issued eagerly; rebuild + step kernels replay as a CUDA graph)")
                               + ", ring agents advanced redundantly (no migration message)",
                "mode": "frozen snapshot (RCS_STEP_NO_COMMIT)" if frozen else "committed steps",
                "l2": "per-rank working set (two state buffer sets + index) larger than L2; no flush between steps",
                "seed": scene.seed, "dt_ns": scene.dt[1],
                "mean_neighbours": agg[0].item() / max(n_live, 1), "candidates_per_agent": agg[1].item() / max(n_live, 1),
                "finite_tti_fraction": agg[2].item() / max(n_live, 1), "nonfinite": int(agg[3].item()),
                "oob": int(agg[4].item()),
            },
            "e2e": e2e, "gpu_launches": int(agg[6].item()), "graph_steps_rank0": g1[0] - g0[0],
            "dist_verified": None if dist_check is None else dist_check["ok"], "dist_verify": dist_check,
            "roofline": {
                "bound": "hbm", "kernel": "step_tile_kernel (+ step_aside_kernel) per rank, slowest rank",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": (achieved / peaks["hbm_gbs"]) if achieved else None, "peak_source": how, "traffic": traffic,
                "traffic_source": "profiles/kernel_counts.json (one-GPU ncu capture of this workload) per agent x agents "
                                  "per rank + 8 B per agent of strip bookkeeping; not captured on this run",
                "algorithmic_bytes_per_agent_step": algo, "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms * K / total_ms,
                "note": "the Zanlungo kernel is FP64-pipe / latency bound, not HBM-bound (DESIGN.md)",
            },
            "cpu_baseline": None, "clocks": clk, "wall_s_timed_region": wall1 - wall0,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def _gather_states(st, dist, torch, rank: int, world: int):
    """(id, x, y, vx, vy) of every rank on all ranks (padded), as raw 64-bit words."""
    n1 = len(st["id"])
    counts = torch.zeros(world, dtype=torch.int64, device="cuda")
    counts[rank] = n1
    dist.all_reduce(counts)
    nmax = max(int(counts.max().item()), 1)
    mine = torch.zeros((5, nmax), dtype=torch.int64, device="cuda")
    for r, k in enumerate(("id", "x", "y", "vx", "vy")):
        mine[r, :n1] = torch.from_numpy(st[k].view(np.int64).copy()).cuda()
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return parts, counts


def verify_stream(dist, torch, rank: int, world: int, local: int, K: int = 120) -> dict:
    """SourceSink spawn / despawn on strips over the NCCL transport (BASELINE config 5 in small): 4 x world source
    sinks, two per strip quarter, whose agents walk through the strip boundaries to sinks in other strips; ids come
    from the per-step spawn bitmap summed over the ranks (ncclAllReduce).  K committed steps, then every rank's agents
    against the same steps on one handle, bit for bit."""
    from . import sim as S
    from .strips import StripSimulation

    cell, ncols = 2.0, 64 * world
    dom = cell * ncols
    za = (0.05, 1.0, 0.0, 0.5, 1.0, 0.2)

    def sources():
        out = []
        for q in range(world):
            x0 = (q + 0.5) * dom / world  # the middle of strip q
            for k, (dx, y) in enumerate([(0.45 * dom, 10.0), (-0.45 * dom, 20.0), (0.3 * dom, 30.0), (0.3 * dom, 30.15)]):
                y += 32.0 * q  # streams of different sources stay out of each other's eyesight
                x1 = min(max(x0 + dx, 4.0), dom - 4.0)
                sp = 1.5 if k == 2 else 1.0  # the faster lane of the overtaking pair has the lower ids
                v = (sp if x1 > x0 else -sp, 0.0)
                wps = [((x0 + x1) / 2, y), (x1, y)]
                hl = S.ConstantVelocityPlan(v)
                out.append(S.SourceSink((x0, y), 0.6, S.MonotonicCrowd(2.0), hl, S.Zanlungo(*za), [wps[-1]], False,
                                        2.0))
        return out

    cap = 4 * world * 300
    sim = StripSimulation(S.LocationHash2D(dom, dom, cell, (0.0, 0.0), capacity=cap, device=local), rank, world,
                          fresh_nccl_id(dist, torch, rank), halo_capacity=2048, peer_gather=peer_gather(dist, torch))
    keep = sources()
    for ss in keep:
        sim.add_source_sink(ss)
    dt = S.Duration(0, 500_000_000)
    for k in range(K):
        sim.step_async(dt)
        if k % 40 == 39:
            sim.sync()
            sim._dispatch_events()
    sim.sync()
    st = sim.read_state()
    parts, counts = _gather_states(st, dist, torch, rank, world)
    ok, detail, n_ref = True, "", 0
    if rank == 0:
        got = np.concatenate([parts[r][:, : int(counts[r].item())].cpu().numpy() for r in range(world)], axis=1)
        got = got[:, np.argsort(got[0].view(np.uint64), kind="stable")]
        single = S.Simulation(S.LocationHash2D(dom, dom, cell, (0.0, 0.0), capacity=cap, device=local))
        keep2 = sources()
        for ss in keep2:
            single.add_source_sink(ss)
        for k in range(K):
            single.step_async(dt)
            if k % 40 == 39:
                single.sync()
                single._dispatch_events()
        single.sync()
        ref = single.read_state()
        n_ref = len(ref["id"])
        want = np.stack([ref[k].view(np.int64) for k in ("id", "x", "y", "vx", "vy")])
        ok = got.shape == want.shape and bool(np.array_equal(got, want))
        if not ok:
            detail = f"shapes {got.shape} vs {want.shape}"
        single.spatial_index.close()
    sim.spatial_index.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return {"ok": bool(int(flag.item())), "n_gpus": world, "steps": K, "source_sinks": 4 * world, "live_agents": n_ref,
            "transport": f"halo by {sim.transport}, spawn bitmap by ncclAllReduce", "workload": "SourceSink stream on strips",
            "detail": detail}


def verify_core(dist, torch, rank: int, world: int, local: int, workload: str, variant: str, dt, K: int) -> dict:
    """K COMMITTED steps of `workload` on `world` ranks over the NCCL transport; every rank's agents are gathered on
    rank 0 and compared bit for bit (ids, x, y, vx, vy) with the same steps on one handle.  Returns the verdict on
    every rank.  (The single-process transport is checked the same way by tests/test_gpu_strips.py; this is the
    check of the real multi-process exchange: ghosts, redundant ring, migration between processes.)"""
    from . import sim as S
    from .strips import StripSimulation, _planners, add_agents_with_ids

    nccl_id = fresh_nccl_id(dist, torch, rank)
    scene, n_total, ids, xy, vxy = strip_scene(workload, variant, rank, world, False)
    per_col = int(round(n_total ** 0.5)) * scene.cell / (5.0 if workload.startswith("sparse") else 1.0)
    w_cols = 1 + int(np.floor(scene.eyesight / scene.cell)) + 1  # ring + stencil reach
    halo_cap = int(1.2 * w_cols * per_col) + 2048
    idx = S.LocationHash2D(scene.width, scene.height, scene.cell, scene.offset, capacity=n_total, device=local)
    sim = StripSimulation(idx, rank, world, nccl_id, halo_capacity=halo_cap, boundaries=scene.meta["bounds"],
                          peer_gather=peer_gather(dist, torch))
    n0 = sim.add_scene_agents(scene, ids, xy, vxy)
    for _ in range(K):
        sim.step_async(dt)
    sim.sync()
    stats = sim.stats()
    st = sim.read_state()
    n1 = len(st["id"])
    parts, counts = _gather_states(st, dist, torch, rank, world)
    mig = torch.tensor([abs(n1 - n0), stats.finite_tti_count], dtype=torch.int64, device="cuda")
    dist.all_reduce(mig)
    ok, detail = True, ""
    if rank == 0:
        got = np.concatenate([parts[r][:, : int(counts[r].item())].cpu().numpy() for r in range(world)], axis=1)
        got = got[:, np.argsort(got[0].view(np.uint64), kind="stable")]
        full_scene, _, fids, fxy, fvxy = strip_scene(workload, variant, 0, 1, False)
        single = S.Simulation(S.LocationHash2D(scene.width, scene.height, scene.cell, scene.offset, capacity=n_total,
                                               device=local))
        hl, lp = _planners(full_scene)
        single._keep = (hl, lp)
        add_agents_with_ids(single, fids, fxy, fvxy, hl, lp, full_scene.eyesight)
        for _ in range(K):
            single.step_async(dt)
        single.sync()
        ref = single.read_state()
        want = np.stack([ref[k].view(np.int64) for k in ("id", "x", "y", "vx", "vy")])
        ok = got.shape == want.shape and bool(np.array_equal(got, want))
        if not ok:
            detail = f"shapes {got.shape} vs {want.shape}"
            if got.shape == want.shape:
                bad = np.nonzero((got != want).any(axis=0))[0]
                detail = f"{len(bad)} agents differ, first id {int(want[0, bad[0]])}"
        single.spatial_index.close()
    sim.spatial_index.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return {"ok": bool(int(flag.item())), "n_gpus": world, "agents": int(n_total), "steps": K,
            "transport": f"halo by {sim.transport}, one process per GPU",
            "workload": f"{workload}, ids {variant}, committed steps of {dt.secs + dt.nanos / 1e9:.4f} s",
            "net_migration_sum_over_ranks": int(mig[0].item()),
            "agents_with_finite_t_i_last_step": int(mig[1].item()), "detail": detail}


# The force-active crowd of the pre-timing check: 65 536 agents at 5 m spacing, R = cell = 5 m (SURVEY.md 8d
# "C2-sparse"), 30 committed steps of 1/30 s: ~6 % of the agents run the force pass every step and ~10 000 change
# their cell column (i.e. migrate when the column is a strip boundary); finite for > 40 steps in the oracle.
DIST_VERIFY = ("sparse256", "shuffled", (0, 33_333_333), 30)


def run_stream(args) -> None:
    """bench.py --gpus N --workload c5: the SourceSink stream of BASELINE config 5 on N strips (stream_bench.build):
    8192 source sinks, ~2.3 M live agents, spawns and despawns every step, agents migrating between the ranks."""
    import torch
    import torch.distributed as dist

    from . import _native as N
    from . import sim as S
    from . import stream_bench

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lp_none = not args.c5_zanlungo
    check = None if getattr(args, "skip_verify", False) else verify_stream(dist, torch, rank, world, local)
    sim, n_src, dom = stream_bench.build(lp_none, device=local, strip=(rank, world, fresh_nccl_id(dist, torch, rank), peer_gather(dist, torch)))
    lib, h = sim._lib, sim._h
    dt = S.Duration(0, 100_000_000)

    def steps(k):
        for _ in range(k):
            N.check(h, lib.rcs_step_async(h, dt.secs, dt.nanos, N.RCS_STEP_DEFAULT))

    done = 0
    while done < 1400:  # fill the building; events are drained so the buffers never overflow
        steps(100)
        sim.sync()
        sim._dispatch_events()
        done += 100
    n_before = sim.agent_count()
    steps(8)  # after the sync: the launch sequence repeats again and is captured as a graph
    launches0 = sim.launch_count()
    K = args.steps
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sim.event_record(0)
    steps(K)
    sim.event_record(1)
    sim.sync()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([sim.event_elapsed_ms(0, 1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    n_after = sim.agent_count()
    st = sim.stats()
    agg = torch.tensor([0.5 * (n_before + n_after), n_after, sim.launch_count() - launches0, st.nonfinite_count],
                       dtype=torch.float64, device="cuda")
    dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    sim._dispatch_events()
    if rank == 0:
        total_ms = float(ms.item())
        print(json.dumps({
            "metric": "agent-steps/sec (query+Zanlungo+integrate)", "value": agg[0].item() * K / (total_ms * 1e-3),
            "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": total_ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5: SourceSink stream, {n_src} sources, device-side route follower, "
                                   f"{'NoLocalPlan' if lp_none else 'Zanlungo'}, committed steps, {world} strips",
                       "agents_live": int(agg[1].item()), "domain_m": dom, "nonfinite": int(agg[3].item())},
            "e2e": None, "gpu_launches": int(agg[2].item()), "dist_verified": None if check is None else check["ok"],
            "dist_verify": check, "roofline": None, "cpu_baseline": None}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def verify(args) -> None:
    """bench.py --gpus N --verify-dist: the NCCL transport against one handle, bit for bit: first the lane-ordered
    crowd (long strides, many migrations, force pass idle), then the force-active sparse crowd."""
    import torch
    import torch.distributed as dist

    from . import sim as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for workload, variant, dt, K in ((args.workload or "side512", "lane", (0, 200_000_000), args.steps), DIST_VERIFY):
        r = verify_core(dist, torch, rank, world, local, workload, variant, S.Duration(*dt), K)
        ok = ok and r["ok"]
        if rank == 0:
            r["verify_dist"] = "ok" if r["ok"] else "FAILED"
            print(json.dumps(r), flush=True)
    r = verify_stream(dist, torch, rank, world, local)
    ok = ok and r["ok"]
    if rank == 0:
        r["verify_dist"] = "ok" if r["ok"] else "FAILED"
        print(json.dumps(r), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(3)


def peer_gather(dist, torch):
    """The halo transport of the multi-process runs: peer stores over NVLink (default), or ncclSend / ncclRecv with
    RCS_HALO=nccl in the environment."""
    from .strips import torch_peer_gather

    return None if os.environ.get("RCS_HALO", "peer") == "nccl" else torch_peer_gather(dist, torch)


def fresh_nccl_id(dist, torch, rank: int) -> bytes:
    from .strips import nccl_unique_id

    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(ident, 0)
    return bytes(ident.cpu().tolist())


def bind_to_gpu_numa_node(local: int) -> dict:
    """Runs this rank on the cores of the NUMA node its GPU hangs off (what `numactl --cpunodebind` would do), so that
    the pinned host buffers it allocates from here on are local to the GPU's PCIe root: with eight ranks moving 100 MB
    per step each, buffers on the far socket put every transfer on the socket interconnect.  RCS_NUMA=0 skips it."""
    info = {"node": None, "cpus": None}
    if os.environ.get("RCS_NUMA", "1") == "0":
        return info
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(vis.split(",")[local]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else local
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        with open(f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"node": node, "cpus": len(cpus)}
    except Exception as e:  # no sysfs entry, no NVML: stay where the launcher put us
        info["error"] = str(e)[:120]
    return info


def run_e2e(sim, scene, steps: int, warmup: int, dist, torch) -> dict:
    """End to end per rank with HOST buffers, every step: upload the preferred velocities of a host-side
    HighLevelPlanner for the rank's agents (pinned -> device), run the step, read x,y,vx,vy of the rank's
    agents back (device -> pinned).  Frozen snapshot, so every rank keeps its agent set."""
    from . import _native as N
    from .sim import Duration

    lib, h = sim._lib, sim._h
    n = sim.agent_count()
    ids = sim.read_state()["id"]
    numa = bind_to_gpu_numa_node(int(os.environ.get("LOCAL_RANK", "0")))

    def pinned(count):
        p = C.c_void_p()
        N.check(None, lib.rcs_host_alloc(max(count, 1) * 8, C.byref(p)))
        arr = np.frombuffer((C.c_char * (max(count, 1) * 8)).from_address(p.value), dtype=np.float64, count=count)
        return arr, p

    pref, pref_p = pinned(2 * n)
    outs = [pinned(n) for _ in range(4)]
    speed = scene.hl[1]
    par = (ids % np.uint64(2)) == 0
    pref[0::2] = np.where(par, -speed[0], speed[0])
    pref[1::2] = np.where(par, -speed[1], speed[1])
    d = Duration(*scene.dt)
    out_n = C.c_uint64()

    split = [0.0, 0.0, 0.0]  # host wall time per call (RCS_E2E_TRACE=1 prints it: where does the host block?)

    def one():
        # the download of this step's results runs on a second stream and overlaps the next step's upload + compute
        ta = time.perf_counter()
        N.check(h, lib.rcs_set_preferred_velocity(h, n, None, pref.ctypes.data_as(N.c_f64p)))
        tb = time.perf_counter()
        N.check(h, lib.rcs_step_async(h, d.secs, d.nanos, N.RCS_STEP_NO_COMMIT))
        tc = time.perf_counter()
        N.check(h, lib.rcs_read_agents_async(h, N.RCS_ORDER_ID, n, None, *[o[0].ctypes.data_as(N.c_f64p) for o in outs],
                                             C.byref(out_n)))
        td = time.perf_counter()
        split[0] += tb - ta
        split[1] += tc - tb
        split[2] += td - tc

    for _ in range(warmup):
        one()
    N.check(h, lib.rcs_read_wait(h))
    dist.barrier()
    t0 = time.perf_counter()
    split[:] = [0.0, 0.0, 0.0]
    for _ in range(steps):
        one()
    tw = time.perf_counter()
    N.check(h, lib.rcs_read_wait(h))
    dist.barrier()
    t1 = time.perf_counter()
    if os.environ.get("RCS_E2E_TRACE") and dist.get_rank() in (0, dist.get_world_size() // 2):
        print("e2e host ms/step: set_preferred_velocity %.3f, step_async %.3f, read_agents_async %.3f; final wait %.3f ms; "
              "total %.3f ms/step" % (1e3 * split[0] / steps, 1e3 * split[1] / steps, 1e3 * split[2] / steps,
                                      1e3 * (t1 - tw), 1e3 * (t1 - t0) / steps), file=sys.stderr, flush=True)
    t = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    for _, p in [(pref, pref_p)] + outs:
        lib.rcs_host_free(p)
    return {"value": tot.item() * steps / t.item(), "unit": "agent-steps/s",
            "h2d_bytes_per_step": int(16 * tot.item()), "d2h_bytes_per_step": int(32 * tot.item()), "steps": steps,
            "numa_rank0": numa,
            "path": "per rank: rcs_set_preferred_velocity(pinned host) + rcs_step_async + "
                    "rcs_read_agents_async(ORDER_ID, pinned host); max over ranks of the wall time"}
