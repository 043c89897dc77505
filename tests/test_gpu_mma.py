"""GPU parity of the tcgen05 (tensor-core) correlation path against exact integer row sums and against
the dp4a path: bit-exact s32 sums, identical match() results."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import assert_results_match, configure, get_image

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tw,th,ne", [(96, 66, 3), (64, 8, 1), (70, 3, 5), (130, 7, 130), (191, 131, 2), (250, 40, 9), (33, 20, 4)])
def test_mma_row_sums_bit_exact(matcher, tw, th, ne):
    rng = np.random.default_rng(tw * 7 + th + ne)
    tpl = rng.integers(0, 256, (th, tw), dtype=np.uint8)
    rois = rng.integers(0, 256, (ne, th + 6, tw + 6), dtype=np.uint8)
    rowsum, rowS, rowQ = matcher.dbgCorrRowsMMA(rois, tpl)
    for e in sorted(set([0, ne // 2, ne - 1])):
        want = np.transpose(O.ccorr_exact_rows(rois[e], tpl), (2, 0, 1))        # [th, 7, 7]
        bad = np.argwhere(rowsum[e].astype(np.int64) != want)
        assert len(bad) == 0, "eval %d: %d mismatches, first at (tr,r,c)=%s got %d want %d" % (
            e, len(bad), bad[0].tolist(), rowsum[e][tuple(bad[0])], want[tuple(bad[0])])
        r64 = rois[e].astype(np.int64)
        for c in range(7):
            assert np.array_equal(rowS[e][:, c], r64[:, c:c + tw].sum(axis=1))
            assert np.array_equal(rowQ[e][:, c], (r64[:, c:c + tw] ** 2).sum(axis=1))


@pytest.mark.parametrize("tw,th,ne", [(96, 66, 3), (64, 8, 1), (70, 3, 5), (130, 7, 130), (191, 131, 2), (250, 40, 9), (33, 20, 4),
                                      (100, 100, 300), (122, 5, 2), (123, 6, 2), (128, 9, 1), (129, 12, 3), (256, 4, 2)])
def test_fused_chain_and_windows_bit_exact(matcher, tw, th, ne):
    """fused kernel: float32 accumulation in template-row order (MatchTemplate's SIMD loop) and exact window sums"""
    rng = np.random.default_rng(tw * 11 + th + ne)
    tpl = rng.integers(0, 256, (th, tw), dtype=np.uint8)
    rois = rng.integers(0, 256, (ne, th + 6, tw + 6), dtype=np.uint8)
    numer, winS, winQ = matcher.dbgCorrFused(rois, tpl)
    for e in sorted(set([0, 1 % ne, ne // 2, ne - 1])):
        rows = O.ccorr_exact_rows(rois[e], tpl)                                  # [7, 7, th] exact row dots
        want = np.zeros((7, 7), np.float32)
        for tr in range(th):
            want = (want + rows[:, :, tr].astype(np.float32)).astype(np.float32)
        assert np.array_equal(numer[e], want), "eval %d numerators differ" % e
        r64 = rois[e].astype(np.int64)
        for r in range(7):
            for c in range(7):
                win = r64[r:r + th, c:c + tw]
                assert winS[e, r, c] == win.sum(), (e, r, c)
                assert winQ[e, r, c] == (win ** 2).sum(), (e, r, c)


def test_window_statistics_stress(matcher):
    """Both tensor-core kernels compute the window sums from the A tiles in shared memory while the TMA pipeline
    refills them: repeat a two-K-chunk, two-tile shape and compare EVERY eval (a stage released before all of its
    reads had completed showed up here as a few wrong tail bytes in a fraction of the runs)."""
    tw, th, ne = 130, 7, 130
    rng = np.random.default_rng(5)
    for rep in range(12):
        tpl = rng.integers(0, 256, (th, tw), dtype=np.uint8)
        rois = rng.integers(0, 256, (ne, th + 6, tw + 6), dtype=np.uint8)
        r64 = rois.astype(np.int64)
        want = np.stack([r64[:, :, c:c + tw].sum(axis=2) for c in range(7)], axis=2)          # [ne, th+6, 7]
        want_q = np.stack([(r64[:, :, c:c + tw] ** 2).sum(axis=2) for c in range(7)], axis=2)
        _, rowS, rowQ = matcher.dbgCorrRowsMMA(rois, tpl)
        assert np.array_equal(rowS, want), "row-split kernel, rep %d" % rep
        assert np.array_equal(rowQ, want_q), "row-split kernel, rep %d" % rep
        _, winS, winQ = matcher.dbgCorrFused(rois, tpl)
        win = np.zeros((ne, 7, 7), np.int64)
        win_q = np.zeros((ne, 7, 7), np.int64)
        for r in range(7):
            win[:, r, :] = want[:, r:r + th, :].sum(axis=1)
            win_q[:, r, :] = want_q[:, r:r + th, :].sum(axis=1)
        assert np.array_equal(winS, win), "fused kernel, rep %d" % rep
        assert np.array_equal(winQ, win_q), "fused kernel, rep %d" % rep


def test_mma_saturated(matcher):
    tw, th, ne = 762, 16, 2
    tpl = np.full((th, tw), 255, np.uint8)
    rois = np.full((ne, th + 6, tw + 6), 255, np.uint8)
    rowsum, rowS, rowQ = matcher.dbgCorrRowsMMA(rois, tpl)
    assert (rowsum == 762 * 255 * 255).all()


@pytest.mark.parametrize("case", ["cfg1_synth", "cfg3_src6", "src8"])
def test_match_same_with_and_without_tensor_cores(matcher, golden_cases, case):
    c = golden_cases[case]
    configure(matcher, c["params"])
    matcher.learnPattern(get_image(c["tpl"]))
    src = get_image(c["src"])
    try:
        matcher.setTensorCores(0)
        a = matcher.match(src)
        matcher.setTensorCores(2)
        b = matcher.match(src)
        matcher.setTensorCores(3)
        b3 = matcher.match(src)
        matcher.setTensorCores(4)
        b4 = matcher.match(src)
    finally:
        matcher.setTensorCores(1)
    d = matcher.match(src)
    assert_results_match(b, a, 0, 0, 0)
    assert_results_match(b3, a, 0, 0, 0)
    assert_results_match(b4, a, 0, 0, 0)
    assert_results_match(d, a, 0, 0, 0)
    assert_results_match(a, c["results"])
