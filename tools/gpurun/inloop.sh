# rcs_step_in_loop at scale: sweeps and wall time for one step of C3 (2^20 agents), ascending-id and random order
timeout 250 python - <<'PY'
import time, numpy as np
import rmf_crowdsim_b200 as R
from rmf_crowdsim_b200 import scenes as SC
for variant in ("shuffled", "lane"):
    scene = SC.config_c3(variant)
    for order_kind in ("ascending", "random"):
        g = SC.build_simulation(scene)
        order = None if order_kind == "ascending" else np.random.default_rng(1).permutation(scene.n).astype(np.uint64)
        g.step_in_loop(R.Duration(*scene.dt), order=order)   # warm-up (allocations)
        g2 = SC.build_simulation(scene)
        g2.step_in_loop(R.Duration(*scene.dt), order=order)
        t0 = time.perf_counter(); sweeps = g2.step_in_loop(R.Duration(*scene.dt), order=order); t1 = time.perf_counter()
        d = SC.build_simulation(scene); d.step(R.Duration(*scene.dt)); d.step(R.Duration(*scene.dt))
        t2 = time.perf_counter(); d.step(R.Duration(*scene.dt)); t3 = time.perf_counter()
        print(f"C3 {variant} {order_kind}: {sweeps} sweeps, {1e3*(t1-t0):.2f} ms per in-loop step; deferred step {1e3*(t3-t2):.2f} ms (host wall, synchronous)", flush=True)
        del g, g2, d
PY
