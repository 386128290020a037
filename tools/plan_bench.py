"""Host planner timing on this machine: native (csrc/host_plan.cpp) vs the Python plan_isprs_batch, batch 64, C=4."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drs_b200 import host  # noqa: E402

rs = np.random.RandomState(0)
B = 64
data = [rs.rand(600, 700, 4)]
labels = [np.zeros((600, 700), np.uint8)]
inst = np.zeros((B, 4), dtype=np.int64)
inst[:, 1], inst[:, 2], inst[:, 3] = rs.randint(0, 500, B), rs.randint(0, 600, B), rs.randint(0, 360, B)
hw = np.asarray([(600, 700)], dtype=np.int32)
print("cores", os.cpu_count())
for crop in (25, 37, 49):
    host.rotation_table(crop)
    for thr in (1, 2, 4, 8):
        pl = host.NativePlanner(threads=thr)
        slot = host.PlanSlot(B, crop, 4)
        np.random.seed(1)
        pl.plan(hw, inst, crop, 4, slot)
        t = time.perf_counter()
        for _ in range(50):
            pl.plan(hw, inst, crop, 4, slot)
        print("crop %d threads %d native %.3f ms" % (crop, thr, (time.perf_counter() - t) / 50 * 1e3))
        pl.close()
    np.random.seed(1)
    t = time.perf_counter()
    for _ in range(5):
        host.plan_isprs_batch(data, labels, inst, crop, True, True)
    print("crop %d python %.3f ms" % (crop, (time.perf_counter() - t) / 5 * 1e3))
