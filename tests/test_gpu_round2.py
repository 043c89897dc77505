"""Round-2 parity tests (GPU, through the C ABI):
  * the top layer against the reference's own numerator (cv::matchTemplate, src/TemplateMatcher.cpp:514) -- final accepted sets
    of the GPU vs the cv-fed oracle on every golden case, cfg4, cfg5 and the random scenes;
  * the top-layer score map exactly as match() computes it (sub-threshold early-out);
  * sub-pixel estimation (src/TemplateMatcher.cpp:1002-1072) at the north_star tolerances;
  * stage-then-match, > 65535 candidates in one batch, malformed BMP headers (advisor findings);
  * the angle-sharded latency pipeline (fpm_match_sharded): virtual ranks on one GPU, NCCL with one rank."""
import numpy as np
import pytest

from tests.helpers import assert_results_match, configure, get_image
from oracle import oracle as O
from tests.test_gpu_parity import _cfg45, _random_scene

pytestmark = pytest.mark.gpu

GOLDEN = ["src8", "src9", "test4_src3", "cfg3_src6", "cfg1_synth", "src4", "cfg2_synth", "src9_subpix"]


def _cv_fed(tpl, src, params):
    om = configure(O.OracleMatcher(), params)
    om.top_numerator = "cv"                       # cv::matchTemplate(TM_CCORR) at the top layer, like the reference
    assert om.learn_pattern(tpl)
    return om.match(src)


def _compare_sets(got, want):
    """identical accepted set within the north_star tolerances; any difference is reported with its score margin.
    Order: score descending on both sides; where the cv-fed scores differ by less than the score tolerance (the DFT
    numerator's float noise decides their order in the reference) only the SET is compared."""
    near_ties = any(abs(a.score - b.score) <= 1e-4 for a, b in zip(want, want[1:]))
    try:
        assert_results_match(got, want, ordered=not near_ties)
    except AssertionError as e:
        gs = sorted((round(r.ptCenter[0]), round(r.ptCenter[1]), r.dMatchScore) for r in got)
        ws = sorted((round(r.ptCenter[0]), round(r.ptCenter[1]), r.score) for r in want)
        raise AssertionError("%s\nGPU  (cx, cy, score): %s\ncv-fed oracle: %s" % (e, gs, ws))


@pytest.mark.parametrize("case", GOLDEN)
def test_final_results_equal_cv_fed_oracle_on_golden_cases(matcher, golden_cases, case):
    c = golden_cases[case]
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    configure(matcher, c["params"])
    assert matcher.learnPattern(tpl)
    _compare_sets(matcher.match(src), _cv_fed(tpl, src, c["params"]))


@pytest.mark.parametrize("seed", list(range(16)))
def test_final_results_equal_cv_fed_oracle_on_random_scenes(matcher, seed):
    src, tpl, params = _random_scene(1000 + seed)
    configure(matcher, params)
    assert matcher.learnPattern(tpl)
    _compare_sets(matcher.match(src), _cv_fed(tpl, src, params))


@pytest.mark.parametrize("cfg", ["cfg4", "cfg5"])
def test_final_results_equal_cv_fed_oracle_on_cfg4_cfg5(matcher, cfg):
    W, H, T = (4096, 3072, 512) if cfg == "cfg4" else (8192, 8192, 1024)
    tpl, frames, om = _cfg45(matcher, W, H, T, [11])
    om.top_numerator = "cv"
    want = om.match(frames[0])
    assert len(want) == 4
    _compare_sets(matcher.match(frames[0]), want)


# ---------------- the production top-layer map ----------------
@pytest.mark.parametrize("case", ["src8", "cfg3_src6", "cfg1_synth", "test4_src3"])
def test_production_top_score_map(matcher, golden_cases, case):
    """match() stores scores that are certainly below reject_below = Score*0.9^top - 0.01 as float32 estimates.
    Every value >= reject_below must be bit-equal to the exact map, every other value must be < reject_below in both."""
    c = golden_cases[case]
    configure(matcher, c["params"])
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    assert matcher.learnPattern(tpl)
    om = configure(O.OracleMatcher(), c["params"])
    om.learn_pattern(tpl)
    top = len(om.td.pyramid) - 1
    lvl = O.build_pyramid(src, top)[top]
    exact = matcher.dbgTopScore(lvl)
    prod, rb = matcher.dbgTopScoreProduction(lvl)
    thresh = c["params"]["score"] * 0.9 ** top
    assert np.isfinite(rb) and abs(rb - (thresh - 0.01)) < 1e-6
    keep = exact >= rb
    assert keep.any()
    assert np.array_equal(prod[keep], exact[keep]), "a score at or above reject_below differs from the exact map"
    assert (prod[~keep] < rb).all(), "an estimate crossed reject_below"
    # nothing the peak search can observe (>= the layer threshold) is an estimate
    assert np.array_equal(prod >= thresh, exact >= thresh)


# ---------------- sub-pixel estimation at the north_star tolerances ----------------
def test_subpixel_matches_oracle_at_north_star_tolerance(matcher, golden_cases):
    c = golden_cases["src9_subpix"]
    configure(matcher, c["params"])
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    assert matcher.learnPattern(tpl)
    got = matcher.match(src)
    om = configure(O.OracleMatcher(), c["params"])
    om.learn_pattern(tpl)
    want = om.match(src)
    assert len(want) == 1
    assert_results_match(got, want, 1e-4, 0.05, 0.01)
    assert_results_match(got, c["results"], 1e-4, 0.05, 0.01)


@pytest.mark.parametrize("seed", list(range(10)))
def test_subpixel_random_scenes(matcher, seed):
    src, tpl, params = _random_scene(2000 + seed)
    params = dict(params, sub_pixel=True, tolerance_angle=max(params["tolerance_angle"], 15.0))
    configure(matcher, params)
    assert matcher.learnPattern(tpl)
    got = matcher.match(src)
    om = configure(O.OracleMatcher(), params)
    om.learn_pattern(tpl)
    want = om.match(src)
    ties = len({r.score for r in want}) != len(want)
    assert_results_match(got, want, 1e-4, 0.05, 0.01, ordered=not ties)


# ---------------- advisor findings ----------------
def test_stage_top_then_match_sweeps_the_whole_schedule(matcher, golden_cases):
    """a partial sweep through the stage API must not truncate the next match() on the same handle"""
    c = golden_cases["src8"]
    configure(matcher, c["params"])
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    assert matcher.learnPattern(tpl)
    fresh = matcher.match(src)
    assert len(fresh) == len(c["results"])
    for a0, a1 in [(0, 3), (0, 0), (5, 9)]:
        matcher.stageTop(src, a0, a1)
        assert_results_match(matcher.match(src), fresh, 0, 0, 0)
        assert_results_match(matcher.matchBatch(src[None])[0], fresh, 0, 0, 0)


def test_more_than_65535_candidates_in_one_batch(fpm_built):
    """gridDim.y of the ROI warp is one candidate per slot: a legal batch with more candidates than that must run in waves"""
    from fastest_image_pattern_matching_b200 import TemplateMatcher
    import fpm_workloads as synth
    m = TemplateMatcher(0, result_capacity=512)
    try:
        m.setSplitBatch(0)
        tpl = synth.background(40, 40, 5, 1.5)
        frames = np.stack([synth.background(400, 400, 100 + i, 1.5) for i in range(32)])
        m.setMaxPositions(200); m.setScore(0.02); m.setToleranceAngle(180); m.setMinReduceArea(256); m.setMaxOverlap(0.9)
        assert m.learnPattern(tpl)
        m.setTrace(True)
        m.match(frames[0])
        n0 = len(m.traceCandidates())
        m.setTrace(False)
        assert n0 * len(frames) > 65535, "test does not reach the limit: %d candidates per frame" % n0
        got = m.matchBatch(frames)
        for i in (0, 7, 31):
            assert_results_match(got[i], m.match(frames[i]), 0, 0, 0)
    finally:
        m.close()


def _bmp(w, h, bpp=8, off=None, size=None, payload=64):
    import struct
    pal = 1024 if bpp == 8 else 0
    off = 54 + pal if off is None else off
    hdr = b"BM" + struct.pack("<IHHI", size if size is not None else off + payload, 0, 0, off)
    dib = struct.pack("<IiiHHIIiiII", 40, w, h, 1, bpp, 0, 0, 0, 0, 0, 0)
    return hdr + dib + bytes(pal) + bytes(payload)


@pytest.mark.parametrize("w,h,bpp,off", [(0x55555556, 1, 24, None), (16, -2 ** 31, 8, None), (2 ** 31 - 1, 2 ** 31 - 1, 8, None),
                                         (70000, 70000, 8, None), (16, 16, 8, 20), (16, 16, 8, 2 ** 31), (4, 4, 24, None)])
def test_ingest_bmp_rejects_malformed_headers(matcher, w, h, bpp, off):
    from fastest_image_pattern_matching_b200.matcher import FpmError
    data = _bmp(w, h, bpp, off, payload=16)               # 4x4x24 needs 48 bytes of pixels: truncated as well
    with pytest.raises(FpmError):
        matcher.ingestBmp(data)
    # the handle (and the CUDA context) is still usable
    good = np.arange(64, dtype=np.uint8).reshape(8, 8)
    assert matcher.dbgPyrDown(good).shape == (4, 4)


# ---------------- angle-sharded latency mode (fpm_match_sharded) ----------------
def _handles(n, params, tpl):
    from fastest_image_pattern_matching_b200 import TemplateMatcher
    hs = [TemplateMatcher(0) for _ in range(n)]
    for m in hs:
        configure(m, params)
        assert m.learnPattern(tpl)
    return hs


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("case", ["src8", "cfg3_src6", "test4_src3"])
def test_sharded_pipeline_with_virtual_ranks_equals_match(matcher, golden_cases, case, world):
    """fpm_match_sharded_virtual: `world` handles of one GPU play the ranks; the two exchanges are block copies into the
    same fixed-size per-rank blocks ncclAllGather fills; every rank's list must be bit-identical to a single-GPU match()."""
    from fastest_image_pattern_matching_b200.matcher import match_sharded_virtual
    c = golden_cases[case]
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    configure(matcher, c["params"])
    assert matcher.learnPattern(tpl)
    want = matcher.match(src)
    assert len(want) == len(c["results"])
    hs = _handles(world, c["params"], tpl)
    try:
        for _ in range(2):                                 # second call: cached plan and buffers
            per_rank = match_sharded_virtual(hs, src)
            for r, got in enumerate(per_rank):
                assert_results_match(got, want, 0, 0, 0)
        # a plain match on a handle that has just played a rank is still the whole sweep
        assert_results_match(hs[-1].match(src), want, 0, 0, 0)
    finally:
        for m in hs:
            m.close()


def test_sharded_virtual_ranks_cfg5_full_size(matcher):
    from fastest_image_pattern_matching_b200.matcher import match_sharded_virtual
    tpl, frames, om = _cfg45(matcher, 8192, 8192, 1024, [11])
    want = matcher.match(frames[0])
    assert len(want) == 4
    params = dict(max_pos=4, score=0.8, tolerance_angle=180, min_reduce_area=256, max_overlap=0.0)
    hs = _handles(8, params, tpl)
    try:
        for got in match_sharded_virtual(hs, frames[0]):
            assert_results_match(got, want, 0, 0, 0)
    finally:
        for m in hs:
            m.close()


def test_sharded_edge_cases_with_virtual_ranks(matcher):
    """more ranks than angles (tolerance 0: one angle), nothing found, guards"""
    import fpm_workloads as synth
    from fastest_image_pattern_matching_b200.matcher import match_sharded_virtual
    tpl = synth.background(48, 40, 3, 2.0)
    src = synth.background(320, 240, 4, 2.0)
    synth.paste_rotated(src, tpl, 150.0, 120.0, 0.0)
    for params in (dict(max_pos=3, score=0.7, tolerance_angle=0.0), dict(max_pos=3, score=0.999, tolerance_angle=30.0)):
        configure(matcher, params)
        assert matcher.learnPattern(tpl)
        want = matcher.match(src)
        hs = _handles(4, params, tpl)
        try:
            for got in match_sharded_virtual(hs, src):
                assert_results_match(got, want, 0, 0, 0)
            assert all(r == [] for r in match_sharded_virtual(hs, np.zeros((20, 20), np.uint8)))     # template larger than source
        finally:
            for m in hs:
                m.close()


def test_sharded_nccl_single_rank(matcher, golden_cases):
    """the NCCL plumbing (dlopen, ncclGetUniqueId, ncclCommInitRank) with a one-rank communicator; host and device frames"""
    import torch
    from fastest_image_pattern_matching_b200.matcher import comm_available, comm_unique_id
    assert comm_available()
    c = golden_cases["src8"]
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    configure(matcher, c["params"])
    assert matcher.learnPattern(tpl)
    want = matcher.match(src)
    matcher.commInit(1, 0, comm_unique_id())
    try:
        assert_results_match(matcher.matchSharded(src), want, 0, 0, 0)
        d = torch.from_numpy(src).cuda()
        assert_results_match(matcher.matchSharded(ptr=d.data_ptr(), shape=src.shape, stride=src.shape[1], on_device=True), want, 0, 0, 0)
    finally:
        matcher.commDestroy()
    assert_results_match(matcher.match(src), want, 0, 0, 0)


# ---------------- ROI warp fused into the tensor-core correlation (fpm_corr_warp_kernel) ----------------
def _match_with_trace(m, src, mode):
    m.setTensorCores(mode)
    m.setTrace(True)
    try:
        res = m.match(src)
        # candidates of a layer are appended with atomicAdd: order by (candidate id, angle) before comparing
        ev = {}
        for l in range(len(m.templateLevels()) - 1):
            e = m.traceEvals(l)
            ev[l] = e[np.lexsort((e[:, 1], e[:, 0]))] if len(e) else e
    finally:
        m.setTrace(False)
        m.setTensorCores(1)
    return res, ev


@pytest.mark.parametrize("case", ["cfg1_synth", "cfg3_src6", "src8", "src4", "src9"])
def test_warp_fused_correlation_is_bit_identical(matcher, golden_cases, case):
    """the kernel that computes the rotated ROI rows inside the tcgen05 producer must reproduce the unfused path
    (fpm_warp_kernel -> fpm_corr_mma_kernel) bit for bit: same per-eval scores and argmax at every layer, same results"""
    c = golden_cases[case]
    configure(matcher, c["params"])
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    assert matcher.learnPattern(tpl)
    want, wev = _match_with_trace(matcher, src, 5)
    got, gev = _match_with_trace(matcher, src, 6)
    assert len(want) == len(c["results"])
    for l in wev:
        assert wev[l].shape == gev[l].shape, "layer %d: eval count" % l
        assert np.array_equal(wev[l], gev[l]), "layer %d: per-eval (angle, score, argmax) records differ" % l
    assert_results_match(got, want, 0, 0, 0)


def test_warp_fused_correlation_edge_cases(matcher):
    """ROI boxes that leave the image (border 0), a batch, unaligned device rows, odd template sizes"""
    import torch
    import fpm_workloads as synth
    rng = np.random.default_rng(3)
    for tw, th, W, H in [(96, 70, 640, 480), (131, 77, 701, 533), (200, 64, 1000, 300)]:
        tpl = synth.background(tw, th, 11, 2.0)
        tpl = np.ascontiguousarray(tpl)
        frames = []
        for i in range(3):
            src = synth.background(W, H, 50 + i, 2.5)
            synth.paste_rotated(src, tpl, tw * 0.45, th * 0.5, float(rng.uniform(-30, 30)))        # partly outside
            synth.paste_rotated(src, tpl, W - tw * 0.4, H - th * 0.45, float(rng.uniform(140, 200)))
            synth.paste_rotated(src, tpl, W * 0.5, H * 0.5, float(rng.uniform(-180, 180)))
            frames.append(src)
        params = dict(max_pos=6, score=0.5, tolerance_angle=180, min_reduce_area=256, max_overlap=0.3)
        configure(matcher, params)
        assert matcher.learnPattern(tpl)
        for mode_pair in [(5, 6)]:
            matcher.setTensorCores(mode_pair[0])
            want = matcher.matchBatch(np.stack(frames))
            matcher.setTensorCores(mode_pair[1])
            got = matcher.matchBatch(np.stack(frames))
            matcher.setTensorCores(1)
            assert sum(len(w) for w in want) >= 3
            for g, w in zip(got, want):
                assert_results_match(g, w, 0, 0, 0)


# ---------------- pyramid descent without host round trips ----------------
@pytest.mark.parametrize("case", ["cfg1_synth", "cfg3_src6", "src8", "test4_src3", "src9_subpix", "cfg2_synth"])
def test_async_descent_equals_synchronous_descent(matcher, golden_cases, case):
    """single-frame latency path: all layers enqueued at once with grids sized to the top-layer count and the live counts read on
    the device -- must return exactly what the per-layer read-back path returns (also through the opt-in fused kernel)"""
    c = golden_cases[case]
    configure(matcher, c["params"])
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    assert matcher.learnPattern(tpl)
    try:
        matcher.setAsyncDescent(0)
        want = matcher.match(src)
        assert len(want) == len(c["results"])
        for mode in (1, 6):
            matcher.setTensorCores(mode)
            matcher.setAsyncDescent(1)
            got = matcher.match(src)
            ties = len({r.dMatchScore for r in want}) != len(want)
            assert_results_match(got, want, 0, 0, 0, ordered=not ties)
            got_b = matcher.matchBatch(np.stack([src, src, src]))          # forced on for a batch as well
            for g in got_b:
                assert_results_match(g, want, 0, 0, 0, ordered=not ties)
    finally:
        matcher.setAsyncDescent(-1)
        matcher.setTensorCores(1)
