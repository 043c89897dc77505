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
    finally:
        matcher.setTensorCores(1)
    d = matcher.match(src)
    assert_results_match(b, a, 0, 0, 0)
    assert_results_match(d, a, 0, 0, 0)
    assert_results_match(a, c["results"])
