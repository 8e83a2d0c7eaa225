"""CPU, world_size 2 and 3 over gloo: the host-side logic of the multi-GPU path (rank-local crowd generation,
strip ownership, the exchange of the 128-byte communicator id and of the peer-store handles, max-over-ranks timing reduction).  No CUDA."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rmf_crowdsim_b200 import dist_bench as DB
from rmf_crowdsim_b200 import scenes as SC
from rmf_crowdsim_b200.strips import column_of, neighbour_handles, owned_mask, torch_peer_gather


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, variant: str, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene, n_total, ids, xy, vxy = DB.strip_scene("side64", variant, rank, world, lp_none=False)
        ncols = int(scene.width / scene.cell)
        bounds = scene.meta["bounds"]  # strips balanced by agent count (rcs_dist_set_boundaries)
        assert bounds[0] == 0 and bounds[-1] == ncols and len(bounds) == world + 1
        c0, c1 = bounds[rank], bounds[rank + 1]
        m = owned_mask(xy[:, 0], scene.offset[0], scene.cell, c0, c1)
        # every agent of the global crowd is owned by exactly one rank
        t = torch.tensor([float(m.sum()), float(ids[m].astype(np.float64).sum())], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        assert int(t[0].item()) == n_total == 64 * 64
        assert t[1].item() == float(n_total * (n_total - 1) // 2)
        # the rank-local generator reproduces the global scene exactly (counter-based RNG)
        full = SC.uniform_crowd(64, variant, margin=64.0)
        sel = ids[m].astype(np.int64)
        assert np.array_equal(full.xy[sel], xy[m]) and np.array_equal(full.vxy[sel], vxy[m])
        assert (scene.width, scene.cell, scene.offset) == (full.width, full.cell, full.offset)
        # the communicator id travels from rank 0 (same code path as the NCCL run, gloo tensors on the CPU)
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            ident = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(ident, 0)
        assert bytes(ident.tolist()) == bytes(range(128))
        # the peer-store transport's handle exchange: every rank's 64-byte IPC handle to every rank, neighbours picked
        mine = bytes([rank + 1] * 64)
        left, right = neighbour_handles(torch_peer_gather(dist, torch)(mine), rank, world)
        assert left == (bytes([rank] * 64) if rank > 0 else None)
        assert right == (bytes([rank + 2] * 64) if rank + 1 < world else None)
        # device-time reduction used for `value`: max over ranks
        ms = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        assert ms.item() == 10.0 + world - 1
        q.put((rank, int(m.sum()), c0, c1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,variant", [(2, "shuffled"), (3, "lane")])
def test_rank_local_generation_and_ownership(world, variant):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, variant, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    assert got[0][2] == 0 and all(got[r][3] == got[r + 1][2] for r in range(world - 1))  # strips tile the columns
    assert sum(g[1] for g in got) == 64 * 64


def test_column_of_follows_the_saturating_cast():
    """location_hash_2d.rs:56: ((x - off) / res) as usize -- truncation, negatives and NaN -> 0."""
    x = np.array([-100.0, -64.0, -63.999, -62.0, 0.0, 1.999, 2.0, np.nan, 1e30])
    cx = column_of(x, -64.0, 2.0)
    assert list(cx[:8]) == [0, 0, 0, 1, 32, 32, 33, 0]
    assert cx[8] > 10**18


def test_transport_selection_and_numa_binding_degrade_gracefully(monkeypatch):
    """bench.py --gpus N: RCS_HALO=nccl selects the NCCL halo (no handle exchange at all); the NUMA binding of the e2e
    buffers is a no-op where the GPU's node cannot be found (no NVML / sysfs entry: this container) and never raises."""
    monkeypatch.setenv("RCS_HALO", "nccl")
    assert DB.peer_gather(dist, torch) is None
    monkeypatch.setenv("RCS_HALO", "peer")
    assert callable(DB.peer_gather(dist, torch))
    before = os.sched_getaffinity(0)
    info = DB.bind_to_gpu_numa_node(0)
    assert set(info) >= {"node", "cpus"}
    if info["node"] is None:
        assert os.sched_getaffinity(0) == before
    else:
        assert 0 < info["cpus"] <= len(before)
    os.sched_setaffinity(0, before)
    monkeypatch.setenv("RCS_NUMA", "0")
    assert DB.bind_to_gpu_numa_node(0) == {"node": None, "cpus": None}


def test_neighbour_handles_at_the_ends_of_the_strip_row():
    hs = [bytes([r + 1] * 64) for r in range(4)]
    assert neighbour_handles(hs, 0, 4) == (None, hs[1])
    assert neighbour_handles(hs, 3, 4) == (hs[2], None)
    assert neighbour_handles(hs, 2, 4) == (hs[1], hs[3])
    with pytest.raises(AssertionError):
        neighbour_handles(hs[:3], 0, 4)
