"""GPU, two or more devices: the multi-process strip path (one process per GPU) against one handle, bit for bit, with
both halo transports -- peer stores over NVLink (CUDA IPC mappings, rcs_dist_peer_*) and ncclSend / ncclRecv.  Runs
`bench.py --gpus 2 --verify-dist` the way the driver launches the multi-GPU bench (torch.distributed.run): the
lane-ordered migration run, the force-active sparse crowd and the SourceSink stream.  Skipped on a single-GPU box,
where tests/test_gpu_strips.py covers the same kernels through the single-process transport."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_two_processes_match_one_handle(transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, RCS_HALO=transport)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", "2",
           "--verify-dist"]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 3, out.stdout[-2000:]
    want = "peer stores" if transport == "peer" else "nccl"
    for ln in lines:
        assert ln["ok"] and ln["n_gpus"] == 2, ln
        assert want in ln["transport"], ln
