"""Force-kernel time of C4 on ONE GPU: a plain handle against a strip group of one rank (same crowd, same kernels; the
strip form adds ownership roles and keep flags).
python tools/strip_overhead.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rmf_crowdsim_b200 import _native as N  # noqa: E402
from rmf_crowdsim_b200 import Duration  # noqa: E402
from rmf_crowdsim_b200 import scenes as SC  # noqa: E402
from rmf_crowdsim_b200 import sim as S  # noqa: E402
from rmf_crowdsim_b200.strips import LocalStripGroup, _planners  # noqa: E402


def kernel_ms(sm, step, K=8):
    lib, h = sm._lib, sm._h
    for _ in range(3):
        step()
    sm.sync()
    N.check(h, lib.rcs_kernel_timing(h, 1))
    for _ in range(K):
        step()
    sm.sync()
    ms, n = C.c_double(), C.c_uint64()
    N.check(h, lib.rcs_kernel_time_ms(h, C.byref(ms), C.byref(n)))
    N.check(h, lib.rcs_kernel_timing(h, 0))
    return ms.value / max(n.value, 1)


def group_kernel_ms(grp, dt, K=8):
    lib = grp._lib
    for _ in range(3):
        grp.step(dt, no_commit=True, sync=False)
    for sm in grp.sims:
        sm.sync()
        N.check(sm._h, lib.rcs_kernel_timing(sm._h, 1))
    for _ in range(K):
        grp.step(dt, no_commit=True, sync=False)
    out = []
    for sm in grp.sims:
        sm.sync()
        ms, n = C.c_double(), C.c_uint64()
        N.check(sm._h, lib.rcs_kernel_time_ms(sm._h, C.byref(ms), C.byref(n)))
        N.check(sm._h, lib.rcs_kernel_timing(sm._h, 0))
        out.append(round(ms.value / max(n.value, 1), 4))
    return out


scene = SC.config_c4("shuffled")
dt = Duration(*scene.dt)
plain = SC.build_simulation(scene)
print("plain handle, 2^24 agents: kernel ms", round(kernel_ms(plain, lambda: plain.step_async(dt, no_commit=True)), 4), flush=True)
plain.spatial_index.close()
del plain
one = LocalStripGroup(scene, 1, capacity=scene.n + 4096, halo_capacity=1024)
print("strip group of one rank:   kernel ms", group_kernel_ms(one, dt), flush=True)
for sm in one.sims:
    sm.spatial_index.close()
del one
