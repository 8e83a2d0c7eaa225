"""One-off check (not a test): creating, using and destroying many handles returns all device memory.
python tools/leak_check.py"""
import ctypes as C
import gc
import sys

import numpy as np

sys.path.insert(0, ".")
import rmf_crowdsim_b200 as R  # noqa: E402
from rmf_crowdsim_b200 import scenes as SC  # noqa: E402
from rmf_crowdsim_b200.strips import LocalStripGroup  # noqa: E402

rt = C.CDLL("libcudart.so")


def free_bytes():
    f, t = C.c_size_t(), C.c_size_t()
    assert rt.cudaMemGetInfo(C.byref(f), C.byref(t)) == 0
    return f.value


def use_everything():
    scene = SC.uniform_crowd(48, "lane", cell=1.0, eyesight=2.0, margin=12.0, seed=3)
    g = SC.build_simulation(scene)
    g.set_trace(True)
    g.step(R.Duration(0, 10_000_000))
    g.read_trace()
    g.step_in_loop(R.Duration(0, 10_000_000))
    g.spatial_index.query_knn(scene.xy[:64], 5)
    g.spatial_index.query_radius(scene.xy[:64], np.full(64, 2.0))
    ss = R.SourceSink((5.0, 5.0), 0.5, R.MonotonicCrowd(10.0), R.ConstantVelocityPlan((1.0, 0.0)), R.NoLocalPlan(),
                      [(9.0, 5.0)], False, 1.0)
    h = SC.build_simulation(scene, capacity=scene.n + 256)
    h._keep = ss
    h.add_source_sink(ss)
    for _ in range(5):
        h.step(R.Duration(0, 100_000_000))
    grp = LocalStripGroup(scene, 3)
    for _ in range(3):
        grp.step(R.Duration(0, 100_000_000))
    grp.read_state()
    for sm in grp.sims:
        sm.spatial_index.close()
    g.spatial_index.close()
    h.spatial_index.close()


use_everything()  # warm-up: context, module load, allocator pools
gc.collect()
base = free_bytes()
for k in range(30):
    use_everything()
    gc.collect()
after = free_bytes()
print("free before", base, "after", after, "difference (bytes)", base - after)
assert base - after < (8 << 20), "device memory leaked"
print("leak check ok")
