"""Launch-list helper: all ranks of a strip-partitioned C4 crowd in ONE process on one GPU (single-process
transport), so that `ncu --metrics gpu__time_duration.sum` sees every kernel of the strip pipeline (everything
but the NCCL transfer).  python tools/strip_profile.py [world] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rmf_crowdsim_b200 import Duration  # noqa: E402
from rmf_crowdsim_b200 import scenes as SC  # noqa: E402
from rmf_crowdsim_b200.strips import LocalStripGroup  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene = SC.config_c4("shuffled")
per_col = 4096 * scene.cell
grp = LocalStripGroup(scene, world, capacity=int(scene.n / world * 1.05) + 8 * int(per_col * 1.5) + 8192,
                      halo_capacity=int(4 * per_col * 1.5) + 4096)
dt = Duration(*scene.dt)
for _ in range(steps):
    grp.step(dt, no_commit=True, sync=False)
for sm in grp.sims:
    sm.sync()
print("agents per rank", grp.agent_counts())
