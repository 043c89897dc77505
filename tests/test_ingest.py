"""Image ingest: BMP file images and RGB32 camera frames decoded to grayscale on the device
(src/MatchToolDialog.cpp:314, :341 cv::imread(IMREAD_GRAYSCALE); :1557-1575 QImage -> Grayscale8)."""
import struct

import cv2
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import assert_results_match, configure, get_image


def _bmp8(idx, palette, top_down=False, clr_used=0):
    h, w = idx.shape
    stride = (w + 3) // 4 * 4
    order = range(h) if top_down else range(h - 1, -1, -1)
    rows = b"".join(bytes(idx[y]) + b"\0" * (stride - w) for y in order)
    n_pal = clr_used or 256
    off = 14 + 40 + 4 * n_pal
    hdr = struct.pack("<2sIHHI", b"BM", off + len(rows), 0, 0, off)
    hdr += struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, 8, 0, len(rows), 0, 0, clr_used, 0)
    return hdr + palette[:n_pal].tobytes() + rows


def _cases():
    rng = np.random.default_rng(7)
    pal = rng.integers(0, 256, (256, 4), dtype=np.uint8)
    pal[:, 3] = 0
    idx = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    col = rng.integers(0, 256, (41, 50, 3), dtype=np.uint8)
    col_odd = rng.integers(0, 256, (19, 33, 3), dtype=np.uint8)
    return {
        "gray8_cv2": bytes(cv2.imencode(".bmp", idx)[1]),
        "pal8_bottom_up": _bmp8(idx, pal),
        "pal8_top_down": _bmp8(idx, pal, top_down=True),
        "pal8_clr_used_64": _bmp8((idx % 64).astype(np.uint8), pal, clr_used=64),
        "bgr24": bytes(cv2.imencode(".bmp", col)[1]),
        "bgr24_odd_width": bytes(cv2.imencode(".bmp", col_odd)[1]),
    }


def test_oracle_bmp_decode_pins():
    """the formulas the device kernel uses, pinned against the OpenCV decoder"""
    rng = np.random.default_rng(3)
    col = rng.integers(0, 256, (23, 31, 3), dtype=np.uint8)
    b, g, r = [col[..., i].astype(np.int64) for i in range(3)]
    want = ((b * 1868 + g * 9617 + r * 4899 + 8192) >> 14).astype(np.uint8)
    assert np.array_equal(O.ingest_bmp(cv2.imencode(".bmp", col)[1]), want)
    pal = rng.integers(0, 256, (256, 4), dtype=np.uint8)
    pal[:, 3] = 0
    idx = rng.integers(0, 256, (17, 29), dtype=np.uint8)
    lut = ((pal[:, 0].astype(np.int64) * 1868 + pal[:, 1].astype(np.int64) * 9617 + pal[:, 2].astype(np.int64) * 4899 + 8192) >> 14)
    for td in (False, True):
        assert np.array_equal(O.ingest_bmp(_bmp8(idx, pal, top_down=td)), lut.astype(np.uint8)[idx])
    # the reference's own test images are 8-bit BMPs with a gray ramp palette: the decode is the identity on the pixels
    src8 = get_image("Src8")
    assert np.array_equal(O.ingest_bmp(cv2.imencode(".bmp", src8)[1]), src8)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_cases()))
def test_gpu_bmp_ingest_bit_exact(matcher, name):
    data = _cases()[name]
    want = O.ingest_bmp(data)
    w, h = matcher.ingestBmp(data)
    assert (h, w) == want.shape
    assert np.array_equal(matcher.ingestedPixels(), want)


@pytest.mark.gpu
def test_gpu_bmp_ingest_rejects_what_it_cannot_decode(matcher):
    from fastest_image_pattern_matching_b200 import FpmError
    good = _cases()["bgr24"]
    for bad in (b"", b"XX" + good[2:], good[:100],                                   # not a BMP, truncated
                bytes(cv2.imencode(".bmp", np.zeros((8, 8, 4), np.uint8))[1])):        # 32-bit bitfields
        with pytest.raises(FpmError):
            matcher.ingestBmp(bad)
    rle = bytearray(good)
    rle[30] = 1                                                                        # BI_RLE8
    with pytest.raises(FpmError):
        matcher.ingestBmp(bytes(rle))


@pytest.mark.gpu
def test_gpu_rgb32_ingest(matcher):
    rng = np.random.default_rng(11)
    px = rng.integers(0, 2 ** 32, (45, 67), dtype=np.uint32)
    matcher.ingestRgb32(px)
    assert np.array_equal(matcher.ingestedPixels(), O.ingest_rgb32(px))


@pytest.mark.gpu
def test_match_on_ingested_frames_equals_host_arrays(matcher, golden_cases):
    """file bytes -> device decode -> learn / match, without the pixels ever crossing the bus as an image"""
    c = golden_cases["src8"]
    configure(matcher, c["params"])
    tpl, src = get_image(c["tpl"]), get_image(c["src"])
    matcher.ingestBmp(bytes(cv2.imencode(".bmp", tpl)[1]))
    assert matcher.learnIngested()
    matcher.ingestBmp(bytes(cv2.imencode(".bmp", src)[1]))
    got = matcher.matchIngested()
    assert_results_match(got, c["results"])
    matcher.learnPattern(tpl)
    assert_results_match(got, matcher.match(src), 0, 0, 0)
