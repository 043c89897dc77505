"""CPU tests of the N>1 host logic: world_size-2 gloo groups (spawned processes) run the angle-sharded
and the frame-sharded drivers of fastest_image_pattern_matching_b200.dist with a deterministic stand-in
engine and must reproduce the single-process answer exactly."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


class FakeEngine:
    """Deterministic stand-in with the stage API of TemplateMatcher (no GPU, no image work)."""
    N_ANGLES = 41

    def stageNumAngles(self, w, h):
        return self.N_ANGLES

    def stageTop(self, src, a0, a1):
        rows = []
        for a in range(a0, a1):
            rng = np.random.default_rng(1000 + a)
            for j in range(int(rng.integers(0, 4))):
                score = float(np.float32(rng.choice([0.5, 0.75, rng.uniform(0.4, 1.0)])))   # includes exact ties
                rows.append([a, float(rng.integers(0, 60)), float(rng.integers(0, 40)), score, 9.0 * a])
        return np.array(rows, np.float64).reshape(-1, 5)

    def stageSortCandidates(self, picks):
        p = np.asarray(picks, np.float64).reshape(-1, 5)
        order = np.argsort(-p[:, 3].astype(np.float32), kind="stable")
        out = p[order].copy()
        out[:, 0] = np.arange(len(out))
        return out

    def stageRefine(self, cands):
        c = np.asarray(cands, np.float64).reshape(-1, 5)
        keep = c[c[:, 3] > 0.45]
        out = keep.copy()
        out[:, 1:3] = keep[:, 1:3] * 64 + keep[:, 0:1] * 0.25
        out[:, 3] = keep[:, 3] * 0.99
        return out

    def stageFinal(self, refined):
        r = np.asarray(refined, np.float64).reshape(-1, 5)
        r = r[np.lexsort((r[:, 0], -r[:, 3]))]
        return [tuple(x) for x in r[r[:, 3] >= 0.6]]

    # throughput mode
    def match(self, frame):
        from fastest_image_pattern_matching_b200.matcher import SingleTargetMatch
        s = float(frame.sum() % 97) / 97.0
        n = int(frame[0, 0]) % 3
        return [SingleTargetMatch((i, s), (i + 1, s), (i + 1, s + 1), (i, s + 1), (i + .5, s + .5), 10.0 * i, 1.0 - 0.1 * i)
                for i in range(n)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from fastest_image_pattern_matching_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = FakeEngine()
    src = np.zeros((48, 64), np.uint8)
    res = D.match_angle_sharded(eng, src, dist, "cpu")
    frames = [np.full((4, 4), i, np.uint8) for i in range(7)]
    fr = D.match_frames_sharded(eng, frames, dist, "cpu", gather=True, batch=1)
    q.put((rank, res, {k: v.tolist() for k, v in fr.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_drivers_match_single_process(world):
    import torch.multiprocessing as mp
    from fastest_image_pattern_matching_b200 import dist as D
    eng = FakeEngine()
    src = np.zeros((48, 64), np.uint8)
    want = D.match_angle_sharded(eng, src, None)
    frames = [np.full((4, 4), i, np.uint8) for i in range(7)]
    want_fr = {k: v.tolist() for k, v in D.match_frames_sharded(eng, frames, None, batch=1).items()}
    assert len(want) > 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res, fr in got:
        assert res == want, "rank %d angle-sharded result differs" % rank
        assert fr == want_fr, "rank %d frame-sharded result differs" % rank


def test_partitions_cover_everything_once():
    from fastest_image_pattern_matching_b200 import dist as D
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 41, 53):
            seen = []
            for r in range(world):
                a0, a1 = D.angle_range(n, r, world)
                seen += list(range(a0, a1))
            assert seen == list(range(n))
            fr = sorted(sum((D.shard_indices(n, r, world) for r in range(world)), []))
            assert fr == list(range(n))


def test_c_abi_partition_equals_python_specification():
    """fpm_shard_angle_range (the arithmetic fpm_match_sharded uses) == dist.angle_range; chunks are equal-sized so the
    allgathered blocks line up with the global angle index"""
    from fastest_image_pattern_matching_b200 import dist as D
    from fastest_image_pattern_matching_b200.matcher import shard_angle_range
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 41, 47, 53, 64):
            chunk = max(1, -(-n // world))
            for r in range(world):
                assert shard_angle_range(n, world, r) == D.angle_range(n, r, world)
                a0, a1 = shard_angle_range(n, world, r)
                assert a0 == min(n, r * chunk) and a1 - a0 <= chunk
