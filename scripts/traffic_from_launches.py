"""Derive per-kernel DRAM traffic and time shares of ONE step from an ncu launch list.

Input: the CSV written by
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file launches.csv python bench.py --device-steps-only --steps 1 --warmup 1 [--workload cfgN]
(two identical steps in the capture; the last one is taken: the longest suffix of the kernel-name sequence that
repeats immediately before itself).  Output: JSON like profiles/r01_traffic.json.
"""
import csv, json, re, sys

path, workload, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
launches = {}
order = []
for r in rows:
    lid = int(r[0])
    if lid not in launches:
        # "void fpm_pyrdown_kernel<(bool)1>(Pd2Args, ...)" -> "fpm_pyrdown_kernel"
        launches[lid] = {"name": re.sub(r"[<(].*", "", re.sub(r"^void\s+", "", r[4])), "grid": r[8]}
        order.append(lid)
    launches[lid][r[12]] = float(r[14]) * ({"us": 1.0, "ms": 1000.0, "ns": 0.001, "s": 1e6}.get(r[13], 1.0) if "time" in r[12]
                                              else {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[13], 1.0))
names = [launches[i]["name"] for i in order]
n = 0
for cand in range(len(names) // 2, 0, -1):
    if names[-cand:] == names[-2 * cand:-cand]:
        n = cand
        break
assert n > 0, "no repeated step found"
step = [launches[i] for i in order[-n:]]
# ROI warp vs top warp: the first fpm_warp_kernel launch of a step is the top-layer sweep
seen_warp = False
for l in step:
    if l["name"] == "fpm_warp_kernel":
        l["name"] = "fpm_warp_kernel(roi)" if seen_warp else "fpm_warp_kernel(top)"
        seen_warp = True
total_us = sum(l["gpu__time_duration.sum"] for l in step)
kern = {}
for l in step:
    k = kern.setdefault(l["name"], {"launches_per_step": 0, "dram": 0.0, "us": 0.0})
    k["launches_per_step"] += 1
    k["dram"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    k["us"] += l["gpu__time_duration.sum"]
out = {"source": "%s: ncu launch list of bench.py --device-steps-only (second of two identical steps, %d launches)" % (path, n),
       "config": {"workload": workload, "batch_per_gpu": batch}, "ncu_step_time_us": total_us,
       "kernels": {name: {"launches_per_step": k["launches_per_step"], "dram_bytes_per_launch": k["dram"] / k["launches_per_step"],
                          "ncu_time_us_per_launch": k["us"] / k["launches_per_step"], "share_of_step": k["us"] / total_us}
                   for name, k in sorted(kern.items(), key=lambda kv: -kv[1]["us"])}}
json.dump(out, sys.stdout, indent=1)
