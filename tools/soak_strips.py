"""Soak check (not a test): committed steps of a lane-ordered crowd on several strips of one GPU stay bit-identical
to one handle.  python tools/soak_strips.py [world] [steps] [side] [margin]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import rmf_crowdsim_b200 as R  # noqa: E402
from rmf_crowdsim_b200 import scenes as SC  # noqa: E402
from rmf_crowdsim_b200.strips import LocalStripGroup  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
side = int(sys.argv[3]) if len(sys.argv) > 3 else 64
margin = float(sys.argv[4]) if len(sys.argv) > 4 else 40.0
scene = SC.uniform_crowd(side, "lane", margin=margin, seed=21)
single = SC.build_simulation(scene)
grp = LocalStripGroup(scene, world)
dt = R.Duration(0, 50_000_000)
t = time.time()


def bits(st):
    return {k: (v.view(np.uint64) if v.dtype == np.float64 else v) for k, v in st.items()}


for k in range(steps):
    try:
        single.step(dt)
        grp.step(dt)
    except Exception as e:  # noqa: BLE001
        print("step", k, "failed:", e, flush=True)
        for r, sm in enumerate(grp.sims):
            try:
                st = sm.stats()
                print(" rank", r, "n", st.n_agents, "oob", st.oob_count, "nonfinite", st.nonfinite_count, flush=True)
            except Exception as e2:  # noqa: BLE001
                print(" rank", r, "stats:", e2, flush=True)
        a = single.read_state()
        print(" single x range", a["x"].min(), a["x"].max(), "y range", a["y"].min(), a["y"].max(),
              "grid", scene.offset, scene.width, flush=True)
        raise SystemExit(1)
    if k % 50 == 49:
        a, b = bits(single.read_state()), bits(grp.read_state())
        ok = all(np.array_equal(a[key], b[key]) for key in a)
        print(k + 1, "identical" if ok else "DIFFERENT", grp.agent_counts(), round(time.time() - t, 1), flush=True)
        if not ok:
            raise SystemExit(2)
print("soak ok")
