"""Multi-GPU check, launched with torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_check.py
Verifies on real GPUs that (a) the angle-sharded latency mode returns on every rank exactly what a single GPU returns --
through the C++/NCCL path (fpm_match_sharded: two ncclAllGather on device buffers) AND through the stage-API specification
(dist.match_angle_sharded, torch.distributed allgathers) -- and (b) frame sharding + gather reproduces the single-GPU results.
(bench.py --mode latency measures the C++ path; this script is the correctness check.)"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from fastest_image_pattern_matching_b200 import TemplateMatcher  # noqa: E402
import fpm_workloads as synth  # noqa: E402
from fastest_image_pattern_matching_b200 import dist as D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    m = TemplateMatcher(local)
    m.setMaxPositions(15); m.setScore(0.8); m.setToleranceAngle(180); m.setMinReduceArea(256)
    tpl, src = synth.load_fixture("Dst6"), synth.load_fixture("Src6")
    assert m.learnPattern(tpl)
    single = D.results_to_rows(m.match(src))
    for _ in range(2):
        res = D.match_angle_sharded(m, src, dist, dev)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    res = D.match_angle_sharded(m, src, dist, dev)
    dt = (time.perf_counter() - t0) * 1e3
    rows = D.results_to_rows(res)
    assert rows.shape == single.shape and np.array_equal(rows, single), "rank %d: angle-sharded != single GPU" % rank
    D.init_sharded(m, dist)                                    # ncclCommInitRank through the C ABI
    for on_dev in (False, True):
        if on_dev:
            d = torch.from_numpy(src).to(dev)
            got = m.matchSharded(ptr=d.data_ptr(), shape=src.shape, stride=src.shape[1], on_device=True)
        else:
            got = m.matchSharded(src)
        rows = D.results_to_rows(got)
        assert rows.shape == single.shape and np.array_equal(rows, single), "rank %d: fpm_match_sharded != single GPU" % rank
    assert m.collectiveCount() >= 5
    t0 = time.perf_counter(); m.match(src); dt1 = (time.perf_counter() - t0) * 1e3
    frames = [src, synth.load_fixture("Src8"), src, synth.load_fixture("Src9"), src]
    m8 = TemplateMatcher(local)
    m8.setMaxPositions(15); m8.setScore(0.8); m8.setToleranceAngle(180)
    m8.learnPattern(tpl)
    # frames of different sizes: per-frame match on the owning rank, then one gather
    fr = D.match_frames_sharded(m8, frames, dist, dev, gather=True, batch=1)
    assert np.array_equal(fr[0], single) and np.array_equal(fr[2], single) and np.array_equal(fr[4], single)
    if rank == 0:
        print("multi_gpu_check ok: world %d, %d targets, angle-sharded %.2f ms vs single-GPU match %.2f ms" % (world, len(res), dt, dt1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
