"""Multi-GPU sharding of the matcher: one process per GPU, torch.distributed for the plumbing.

Two modes:

* throughput -- `match_frames_sharded`: frames are independent units, frame i goes to rank
  i % world; every rank runs the whole pipeline on its frames; NO data-path collective.  The small
  per-frame result lists can be gathered to every rank afterwards (one all_gather of a padded tensor).

* latency -- ONE frame, every rank holds (or uploads 1/N of) the same frame and builds the pyramid
  redundantly, which is cheaper than exchanging it.  The top-layer angle schedule is split
  contiguously over the ranks, the per-rank pick lists are exchanged with an allgather, every rank
  sorts the union identically (score descending, ties in (angle, pick) order like the oracle's stable
  sort), candidate k is refined by rank k % world, the refined records are exchanged with a second
  allgather and every rank runs the identical final filter/NMS.

  The product path is C++: `fpm_match_sharded` (csrc/fpm_host.cu) issues the two `ncclAllGather` calls on
  device buffers that the sort / NMS kernels consume in place.  `init_sharded` only distributes the
  128-byte ncclUniqueId through torch.distributed and `match_sharded` is a one-line call; nothing of the
  data path runs in Python.

  `match_angle_sharded` is the same schedule written against the stage API of `TemplateMatcher`
  (stageNumAngles, stageTop, stageSortCandidates, stageRefine, stageFinal) with torch.distributed
  all_gathers: the executable specification of the partitioning, testable on CPU with a gloo group and
  a stand-in engine (tests/test_dist_cpu.py), and the cross-check of the C++ path on real GPUs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """frame i -> rank i % world"""
    return list(range(rank, n, world))


def angle_range(n_angles: int, rank: int, world: int):
    """contiguous split of the top-layer angle schedule (same arithmetic as fpm_shard_angle_range)"""
    per = max(1, (n_angles + world - 1) // world)
    a0 = min(n_angles, rank * per)
    return a0, min(n_angles, a0 + per)


def init_sharded(matcher, dist=None):
    """Collective: gives `matcher` (a TemplateMatcher on this rank's GPU) an NCCL communicator over the ranks of
    the torch.distributed group.  Rank 0 creates the ncclUniqueId through the C ABI; torch.distributed only carries
    the 128 bytes."""
    from .matcher import comm_unique_id
    if dist is None or not dist.is_initialized():
        matcher.commInit(1, 0, comm_unique_id())
        return 1
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    matcher.commInit(world, rank, box[0])
    return world


def match_sharded(matcher, src=None, **kw):
    """Latency mode through the C++/NCCL path (fpm_match_sharded); identical result list on every rank."""
    return matcher.matchSharded(src, **kw)


def _all_gather_rows(rows: np.ndarray, ncols: int, dist, device) -> np.ndarray:
    """all_gather of a variable number of float64 rows: one count exchange + one padded gather.
    Returns the concatenation in rank order (identical on every rank)."""
    import torch
    world = dist.get_world_size()
    rows = np.ascontiguousarray(rows, np.float64).reshape(-1, ncols)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros((cap, ncols), dtype=torch.float64, device=device)
    if rows.shape[0]:
        buf[:rows.shape[0]] = torch.from_numpy(rows).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = [o[:c].cpu().numpy() for o, c in zip(out, counts)]
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, ncols))


def match_angle_sharded(engine, src: np.ndarray, dist=None, device="cpu"):
    """Latency mode for one frame; returns the identical result list on every rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
    n_ang = engine.stageNumAngles(src.shape[1], src.shape[0])
    a0, a1 = angle_range(n_ang, rank, world)
    picks = engine.stageTop(src, a0, a1)                      # rows {angle_index, x, y, score, angle}
    if world > 1:
        picks = _all_gather_rows(picks, 5, dist, device)      # rank order == global angle order
    cands = engine.stageSortCandidates(picks)                 # identical on every rank
    mine = cands[rank::world]
    refined = engine.stageRefine(mine)
    if world > 1:
        refined = _all_gather_rows(refined, 5, dist, device)
    # deterministic order for the final stage: candidate id ascending
    if len(refined):
        refined = refined[np.argsort(refined[:, 0], kind="stable")]
    return engine.stageFinal(refined)


def results_to_rows(results) -> np.ndarray:
    """list of SingleTargetMatch -> [n, 12] float64 (the fpm_result field order)"""
    rows = np.zeros((len(results), 12), np.float64)
    for i, r in enumerate(results):
        rows[i] = [r.dMatchScore, r.dMatchedAngle, r.ptCenter[0], r.ptCenter[1], r.ptLT[0], r.ptLT[1], r.ptRT[0], r.ptRT[1],
                   r.ptRB[0], r.ptRB[1], r.ptLB[0], r.ptLB[1]]
    return rows


def match_frames_sharded(engine, frames: Sequence[np.ndarray], dist=None, device="cpu", gather: bool = True,
                         batch: Optional[int] = None):
    """Throughput mode: this rank matches frames rank, rank+world, ...; no data-path collective.

    Returns {frame index: [n, 12] rows}; with gather=True every rank receives all frames' rows."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
    idx = shard_indices(len(frames), rank, world)
    local = {}
    if idx:
        if hasattr(engine, "matchBatch") and batch != 1:
            res = engine.matchBatch(np.stack([frames[i] for i in idx]))
        else:
            res = [engine.match(frames[i]) for i in idx]
        for i, r in zip(idx, res):
            local[i] = results_to_rows(r)
    if world == 1 or not gather:
        return local
    # one padded gather: rows prefixed with their frame index
    rows = [np.concatenate([np.full((v.shape[0], 1), float(k)), v], axis=1) for k, v in local.items() if v.shape[0]]
    flat = np.concatenate(rows, axis=0) if rows else np.zeros((0, 13))
    allrows = _all_gather_rows(flat, 13, dist, device)
    out = {i: np.zeros((0, 12)) for i in range(len(frames))}
    for i in range(len(frames)):
        sel = allrows[allrows[:, 0] == float(i)]
        out[i] = sel[:, 1:]
    return out
