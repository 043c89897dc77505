"""Single-frame latency breakdown on one GPU: wall-clock p50 of a device-resident batch-1 match, the per-kernel device
times of the same call (CUDA events around every launch) and the launch count.
    python scripts/latency_breakdown.py [cfg1|cfg2|cfg3|cfg4|cfg5]"""
import ctypes as C
import json
import os
import statistics
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from fastest_image_pattern_matching_b200 import TemplateMatcher  # noqa: E402
from fastest_image_pattern_matching_b200 import _lib as L  # noqa: E402


def main():
    wl_name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    wl = bench.WORKLOADS[wl_name]
    tpl, frames = bench.make_frames(1, 4242, wl_name)
    d = torch.from_numpy(frames[0]).cuda()
    H, W = frames[0].shape
    m = TemplateMatcher(0, result_capacity=256)
    bench.configure(m, wl)
    assert m.learnPattern(tpl)
    cap = m.result_capacity
    res = (L.fpm_result * cap)()
    n = (C.c_int * 1)()
    lat = []
    for i in range(10 + 200):
        t0 = time.perf_counter()
        m.matchBatchRaw(d.data_ptr(), 1, W, H, W, H * W, True, res, n)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = lat[10:]
    l0 = m.launchCount()
    m.setProfile(True); m.profileReset()
    for i in range(20):
        m.matchBatchRaw(d.data_ptr(), 1, W, H, W, H * W, True, res, n)
    prof = {k: (round(v[0] / 20 * 1e3, 1), v[1] // 20) for k, v in m.profile().items() if v[1]}
    m.setProfile(False)
    print(json.dumps({"workload": wl_name, "targets": n[0], "p50_ms": statistics.median(lat), "p10_ms": sorted(lat)[20],
                      "launches_per_match": (m.launchCount() - l0) / 20, "kernel_us_and_launches": prof,
                      "kernel_sum_us": round(sum(v[0] for v in prof.values()), 1)}))


if __name__ == "__main__":
    main()
