#!/usr/bin/env python
"""bench.py -- throughput of the B200 NCC matcher on BASELINE.json's headline configuration.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload = "cfg1"): 4024x3036 u8 source (seeded synthetic stand-in for the missing
Src7.bmp: blurred noise + Dst7 at three jittered README poses), the real 762x521 Dst7 template,
TargetNum 3, Score 0.8, ToleranceAngle 180, MinReducedArea 256.  One "step" = one pass of
TemplateMatcher::match over a batch of `--batch` frames per GPU.

value  : images/sec, whole job, frames resident in HBM when the timed region starts.
e2e    : same metric through the C ABI with HOST (pinned) frames, H2D + D2H inside the timed region.
roofline / cpu_baseline: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec (4024x3036 src, 762x521 tpl, +-180deg)"
UNIT = "images/s"
WORKLOADS = {
    # BASELINE.json configs; cfg1 is the one the metric is quoted on (the default), the others are informational
    "cfg1": dict(w=4024, h=3036, tpl="762x521 (Dst7.bmp)", max_pos=3, score=0.8, tol=180.0, mra=256, overlap=0.0, batch=64,
                 expect=3, metric="images/sec (4024x3036 src, 762x521 tpl, +-180deg)"),
    "cfg2": dict(w=3648, h=3648, tpl="54x54 (Dst10.jpg)", max_pos=200, score=0.7, tol=0.0, mra=256, overlap=0.0, batch=32,
                 expect=205, metric="images/sec (3648x3648 src, 54x54 tpl x576, angle 0, TargetNum 200)"),
    "cfg3": dict(w=4096, h=3000, tpl="848x446 (Dst6.bmp)", max_pos=15, score=0.8, tol=180.0, mra=256, overlap=0.0, batch=32,
                 expect=15, metric="images/sec (Src6.jpg 4096x3000, Dst6 848x446, +-180deg, TargetNum 15)"),
    "cfg4": dict(w=4096, h=3072, tpl="512x512 synthetic", max_pos=4, score=0.8, tol=180.0, mra=256, overlap=0.0, batch=32,
                 expect=4, metric="images/sec (4096x3072 synthetic, 512x512 tpl, +-180deg)"),
    "cfg5": dict(w=8192, h=8192, tpl="1024x1024 synthetic", max_pos=4, score=0.8, tol=180.0, mra=256, overlap=0.0, batch=8,
                 expect=4, metric="images/sec (8192x8192 synthetic, 1024x1024 tpl, +-180deg)"),
}


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
def make_frames(n_distinct: int, seed0: int, workload: str = "cfg1"):
    """(template, [n, H, W] frames) of a BASELINE.json config: seeded synthetic sources, real templates"""
    import numpy as np
    from fastest_image_pattern_matching_b200 import synth
    if workload == "cfg1":
        tpl = synth.load_fixture("Dst7")
        frames = [synth.cfg1_source(seed=seed0 + i, tpl=tpl, jitter=True) for i in range(n_distinct)]
    elif workload == "cfg2":
        tpl = synth.load_fixture("Dst10")
        frames = [synth.cfg2_source(seed=seed0 + i, tpl=tpl) for i in range(min(n_distinct, 2))]
    elif workload == "cfg3":
        tpl = synth.load_fixture("Dst6")
        frames = [synth.load_fixture("Src6")]
    elif workload == "cfg4":
        tpl = synth.synth_template(512, 4)
        frames = [synth.synth_frame(4096, 3072, tpl, seed0 + i, 4) for i in range(n_distinct)]
    elif workload == "cfg5":
        tpl = synth.synth_template(1024, 4)
        frames = [synth.synth_frame(8192, 8192, tpl, seed0 + i, 4) for i in range(min(n_distinct, 2))]
    else:
        raise ValueError(workload)
    return tpl, np.stack(frames)


def configure(m, wl):
    m.setMaxPositions(wl["max_pos"]); m.setScore(wl["score"]); m.setToleranceAngle(wl["tol"])
    m.setMinReduceArea(wl["mra"]); m.setMaxOverlap(wl["overlap"])


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle (Python/cv2 restatement + SSE2 numerator = the reference's CPU path)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n_images, wl, workload = args
    import cv2
    import numpy as np  # noqa: F401
    cv2.setNumThreads(1)
    from oracle.oracle import OracleMatcher
    tpl, frames = make_frames(min(n_images, 2), seed, workload)
    m = OracleMatcher()
    m.max_pos, m.score, m.tolerance_angle, m.min_reduce_area, m.max_overlap = wl["max_pos"], wl["score"], wl["tol"], wl["mra"], wl["overlap"]
    m.learn_pattern(tpl)
    m.match(frames[0])                                  # warm-up (page-in, cv2 init)
    t0 = time.perf_counter()
    found = 0
    for i in range(n_images):
        found += len(m.match(frames[i % len(frames)]))
    return time.perf_counter() - t0, found


def cpu_throughput(wl, workers: int, images_per_worker: int, workload: str = "cfg1"):
    """images/sec of the CPU oracle with `workers` processes (one single-threaded matcher each)."""
    import multiprocessing as mp
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "libncc_rowdot.so"], check=True, capture_output=True)
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(100 + 10 * r, images_per_worker, wl, workload) for r in range(workers)])
        wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    total = workers * images_per_worker
    return total / busy, busy, wall, sum(r[1] for r in res)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_numa(local_rank: int):
    """Best effort: run this rank (and first-touch its pinned frame buffers) on the NUMA node of its GPU, so that the
    H2D DMA does not cross the socket interconnect.  Returns a short description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return "numa node unknown"
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "numa node %d (%d cpus)" % (node, len(allowed))
        return "numa node %d (no allowed cpus)" % node
    except Exception as e:  # pragma: no cover
        return "not bound (%s)" % type(e).__name__


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def emit(line: dict):
    """the ONE JSON line on the real stdout (fd 1 is pointed at stderr while the bench runs, so that banners
    printed by NCCL / torchrun children / libraries cannot pollute it)"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (0 = workload default: 64 for cfg1)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-steps-only", action="store_true",
                    help="run only warm-up + timed device-resident steps and exit (clean launch list for ncu; prints no bench line)")
    ap.add_argument("--cpu-images", type=int, default=0, help="images per worker for the CPU baseline (0 = auto)")
    ap.add_argument("--h2d-chunk", type=int, default=0, help="frames per H2D chunk of the e2e path (0 = library default)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    METRIC = wl["metric"]
    if args.batch <= 0:
        args.batch = wl["batch"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ---------------- reference arm: the CPU implementation on all host cores -----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = host_cores()
        per_worker = max(1, args.cpu_images or 2)
        vals, times = [], []
        for _ in range(max(args.warmup, 0)):
            pass                                       # each worker warms itself up (one untimed match)
        for _ in range(max(args.steps, 1)):
            v, busy, wall, found = cpu_throughput(wl, cores, per_worker, args.workload)
            vals.append(v); times.append(busy)
            if sum(times) > 150:                       # bounded: the whole run must end within minutes
                break
        value = (cores * per_worker * len(vals)) / sum(times)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1000.0 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "src": "%dx%d" % (wl["w"], wl["h"]), "tpl": wl["tpl"],
                       "target_num": wl["max_pos"], "score": wl["score"], "tolerance_angle": wl["tol"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps x %d processes x %d frames; Python/cv2 4.13 oracle + SSE2 IM_Conv_SIMD restatement, "
                                       "one single-threaded matcher per core (the reference full match() needs OpenCV C++/Qt, unbuildable here)"
                                       % (len(vals), cores, per_worker)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(line)
        return 0

    # ---------------- CPU baseline first (before CUDA is initialised in this process) ----------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.device_steps_only:
        cores = host_cores()
        workers = cores
        per_worker = args.cpu_images or 8
        v, busy, wall, found = cpu_throughput(wl, workers, per_worker, args.workload)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
                        "sample": "%d processes x %d frames of the workload (%.1f s); Python/cv2 4.13 oracle + SSE2 IM_Conv_SIMD restatement, "
                                  "single-threaded per process like the reference's own loops" % (workers, per_worker, busy),
                        "single_core_ms_per_match": 1000.0 * busy / per_worker}

    import numpy as np
    import torch
    import ctypes as C
    from fastest_image_pattern_matching_b200 import TemplateMatcher, build
    from fastest_image_pattern_matching_b200 import _lib as L

    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; stdout must carry ONE JSON line
        if not os.environ.get("FPM_KEEP_NCCL_DEBUG"):
            os.environ.pop("NCCL_DEBUG", None)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else "single process, not bound"
    if rank == 0:
        build()
    if dist:
        dist.barrier()
    B, K, W = args.batch, args.steps, max(args.warmup, 3)

    # distinct frames: B frames per step; 12.2 MB each -> one step reads 195 MB (> 126 MB L2), and the
    # steps alternate between two such sets so nothing survives in L2 from one step to the next
    tpl, frames_np = make_frames(min(B, 8), 1000 * (rank + 1), args.workload)
    reps = (B + frames_np.shape[0] - 1) // frames_np.shape[0]
    H, Wd = frames_np.shape[1:]
    pitch = (Wd + 127) // 128 * 128
    dev_sets = []
    for s in range(2):
        d = torch.empty((B, H, pitch), dtype=torch.uint8, device="cuda")
        src = torch.from_numpy(np.roll(frames_np, s, axis=0))
        for b in range(B):
            d[b, :, :Wd].copy_(src[b % src.shape[0]])
        dev_sets.append(d)
    host_sets = []
    for s in range(2):
        hbuf = torch.empty((B, H, Wd), dtype=torch.uint8).pin_memory()
        src = torch.from_numpy(np.roll(frames_np, s, axis=0))
        for b in range(B):
            hbuf[b].copy_(src[b % src.shape[0]])
        host_sets.append(hbuf)
    del reps

    m = TemplateMatcher(local_rank, result_capacity=16 if wl["expect"] <= 16 else 256)
    configure(m, wl)
    assert m.learnPattern(tpl)
    if args.h2d_chunk:
        m.setH2DChunk(args.h2d_chunk)
    if os.environ.get("FPM_TC"):
        m.setTensorCores(int(os.environ["FPM_TC"]))
    if os.environ.get("FPM_SPLIT") is not None:
        m.setSplitBatch(int(os.environ["FPM_SPLIT"]))
    if os.environ.get("FPM_WS_MB"):
        m.setWorkspaceMB(float(os.environ["FPM_WS_MB"]))
    cap = m.result_capacity
    res = (L.fpm_result * (cap * B))()
    counts = (C.c_int * B)()

    def step_device(i):
        d = dev_sets[i & 1]
        m.matchBatchRaw(d.data_ptr(), B, Wd, H, pitch, H * pitch, True, res, counts)

    def step_host(i):
        hb = host_sets[i & 1]
        m.matchBatchRaw(hb.data_ptr(), B, Wd, H, Wd, H * Wd, False, res, counts)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn):
        for i in range(W):
            step_fn(i)
        barrier()
        l0 = m.launchCount()
        m.timerRecord(0)
        t0 = time.perf_counter()
        for i in range(K):
            step_fn(i)
        m.timerRecord(1)
        ms = m.timerElapsedMs()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        launches = m.launchCount() - l0
        barrier()
        ms = max(ms, 0.0)
        if dist:
            t = torch.tensor([ms, wall_ms, float(launches)], device="cuda", dtype=torch.float64)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tmax[0]), float(tmax[1]), int(tsum[2])
        return ms, wall_ms, launches

    if args.device_steps_only:
        for i in range(max(args.warmup, 1) + K):
            step_device(i)
        torch.cuda.synchronize()
        sys.stderr.write("device-steps-only: %d steps of %d frames done\n" % (max(args.warmup, 1) + K, B))
        if dist:
            dist.destroy_process_group()
        return 0

    # sanity: every frame must yield the 3 pasted targets
    step_device(0)
    found = [counts[b] for b in range(B)]
    ok_found = all(f == wl["expect"] for f in found)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, dev_wall_ms, launches = timed(step_device)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms, e2e_wall_ms, _ = timed(step_host)

    # raw pinned-host -> device copy rate of one step's frames (context for e2e: the PCIe bound)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch = torch.empty_like(host_sets[0], device="cuda")
    scratch.copy_(host_sets[0], non_blocking=True)
    ev0.record()
    for _ in range(3):
        scratch.copy_(host_sets[0], non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * B * H * Wd / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    del scratch

    # per-kernel device time over the same K steps, CUDA events around every launch on the launch stream
    # (whole batch on one handle here: with the two concurrent half-batches of the timed steps the event-bracketed
    #  durations of overlapping kernels would not add up)
    split_default = m.getSplitBatch()
    m.setSplitBatch(0)
    m.setProfile(True)
    m.profileReset()
    for i in range(K):
        step_device(i)
    prof = m.profile()
    m.setProfile(False)
    m.setSplitBatch(split_default)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    total_images = world * B * K
    value = total_images / (dev_ms / 1000.0)
    e2e_value = total_images / (e2e_ms / 1000.0)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    bf16_sust = peaks.get("bf16_tflops_sustained", 1400.0)

    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tj.get("config", {}).get("batch_per_gpu") == B and tj.get("config", {}).get("workload") == args.workload:
            traffic = {k: v["dram_bytes_per_launch"] for k, v in tj["kernels"].items()}
    except Exception:
        pass
    kern = {}
    step_ms = sum(v[0] for v in prof.values()) / max(K, 1)
    for name, (ms, n, work) in prof.items():
        if n == 0:
            continue
        kern[name] = {"ms_per_step": ms / K, "launches_per_step": n / K, "avg_launch_us": 1000.0 * ms / n,
                      "share": (ms / K) / step_ms if step_ms > 0 else None, "work_per_launch": work / n}
    dom = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
    roofline = None
    if dom:
        kd = kern[dom]
        per_launch_s = kd["avg_launch_us"] * 1e-6
        if "corr" in dom or "top_score" in dom:
            # integer MACs on the CUDA-core dp4a pipe; reported against the int8 tensor peak the
            # north_star names (not measured on this pool: 2 x measured sustained bf16 dense)
            peak = 2.0 * bf16_sust
            ach = 2.0 * kd["work_per_launch"] / per_launch_s / 1e12
            roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "traffic": traffic.get(dom), "peak_source": "2 x bf16_tflops_sustained (int8 dense not in MEASURED_PEAKS.json)",
                        "note": "u8xu8->s32 MACs via dp4a; algorithmic ops = 2*49*w*h per eval"}
        else:
            ach = kd["work_per_launch"] / per_launch_s / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": traffic.get(dom), "traffic_source": "ncu dram__bytes_read+write per launch, profiles/r01_traffic.json",
                        "peak_source": hbm_src}
    # the HBM-bound stage the north_star names first (pyramid): always reported beside the dominant kernel
    hbm_kernels = {}
    for name in ("fpm_pyrdown_kernel", "fpm_warp_kernel(roi)"):
        if name in kern:
            kd = kern[name]
            ach = kd["work_per_launch"] / (kd["avg_launch_us"] * 1e-6) / 1e9
            hbm_kernels[name] = {"achieved_GBps": ach, "frac_of_hbm_peak": ach / hbm_peak, "traffic": traffic.get(name)}
    # the tensor-core correlation: algorithmic int8 ops against the (unmeasured) int8 dense peak, and its DRAM side
    for name in ("fpm_corr_mma_kernel", "fpm_corr_fused_kernel"):
        if name in kern:
            kd = kern[name]
            tops = 2.0 * kd["work_per_launch"] / (kd["avg_launch_us"] * 1e-6) / 1e12
            hbm_kernels[name] = {"achieved_TOPS": tops, "frac_of_int8_peak": tops / (2.0 * bf16_sust),
                                 "int8_peak_assumed_TOPS": 2.0 * bf16_sust, "traffic": traffic.get(name),
                                 "note": "HBM-bound at this batch size: the ROI patches of a step exceed L2"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "p50_ms_per_match_batch1": None, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "src": "%dx%d" % (Wd, H), "tpl": wl["tpl"], "target_num": wl["max_pos"],
                   "score": wl["score"], "tolerance_angle": wl["tol"], "min_reduce_area": wl["mra"], "batch_per_gpu": B,
                   "global_batch": world * B, "sharding": "frames over ranks, no data-path collective",
                   "concurrent_half_batches": bool(split_default and B >= split_default and wl["tol"] > 0),
                   "l2": "step input %.0f MB > 126 MB L2; two alternating frame sets" % (B * H * Wd / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * H * Wd, "d2h_bytes_per_step": B * cap * 96 + B * 4,
                "ms_per_step": e2e_ms / K, "h2d_copy_only_GBps": h2d_gbps, "host_numa": numa,
                "pcie_bound_images_per_s": world * h2d_gbps * 1e9 / (H * Wd)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kern,
        "hbm_kernels": hbm_kernels,
        "cpu_baseline": cpu_baseline,
        "targets_found_per_frame_ok": ok_found,
        "wall_ms_per_step": dev_wall_ms / K,
    }
    # p50 latency of a single-frame match (batch 1, device resident)
    lat = []
    d1 = dev_sets[0]
    for i in range(3 + 20):
        t0 = time.perf_counter()
        m.matchBatchRaw(d1[i % B].data_ptr(), 1, Wd, H, pitch, H * pitch, True, res, counts)
        lat.append((time.perf_counter() - t0) * 1000.0)
    line["p50_ms_per_match_batch1"] = statistics.median(lat[3:])
    emit(line)
    m.close()
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
