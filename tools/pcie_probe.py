"""Diagnostic (not part of the product): host <-> device copy bandwidth per rank, alone and with all ranks at once,
plus the NUMA picture, to explain the end-to-end numbers of bench.py --gpus N.
  python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py
"""
import ctypes
import glob
import os
import subprocess
import time

import torch
import torch.distributed as dist


def numa_of_gpu(i):
    try:
        bus = torch.cuda.get_device_properties(i).pci_bus_id
        dom = torch.cuda.get_device_properties(i).pci_domain_id
        dev = torch.cuda.get_device_properties(i).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        return int(open(path).read())
    except Exception as e:  # noqa: BLE001
        return f"? ({e})"


def bw(nbytes, h2d, d2h, both, reps=5):
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    t0 = time.perf_counter()
    for _ in range(reps):
        if both or h2d is not None and d2h is None:
            with torch.cuda.stream(s1):
                h2d[1].copy_(h2d[0], non_blocking=True)
        if both or d2h is not None and h2d is None:
            with torch.cuda.stream(s2):
                d2h[1].copy_(d2h[0], non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
        for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
            print(p, open(p).read().strip(), flush=True)
        for lib in ("libnuma.so.1", "libnuma.so"):
            try:
                ctypes.CDLL(lib)
                print("libnuma:", lib, flush=True)
                break
            except OSError:
                pass
    dist.barrier()
    n = 128 << 20
    hp_up = torch.empty(n, dtype=torch.uint8).pin_memory()
    hp_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_dn = torch.empty(n, dtype=torch.uint8, device="cuda")
    info = f"rank {rank}: gpu numa {numa_of_gpu(local)}, cpus {sorted(os.sched_getaffinity(0))}"
    solo = None
    for r in range(world):
        dist.barrier()
        if r == rank:
            bw(n, (hp_up, d_up), None, False, 2)
            solo = (bw(n, (hp_up, d_up), None, False), bw(n, None, (d_dn, hp_dn), False),
                    bw(2 * n, (hp_up, d_up), (d_dn, hp_dn), True))
    dist.barrier()
    allr = (bw(n, (hp_up, d_up), None, False), )
    dist.barrier()
    allr += (bw(n, None, (d_dn, hp_dn), False), )
    dist.barrier()
    allr += (bw(2 * n, (hp_up, d_up), (d_dn, hp_dn), True), )
    dist.barrier()
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(f"{info}\n   alone  h2d {solo[0]:.1f} d2h {solo[1]:.1f} both {solo[2]:.1f} GB/s | all ranks at once "
                  f"h2d {allr[0]:.1f} d2h {allr[1]:.1f} both {allr[2]:.1f} GB/s", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
