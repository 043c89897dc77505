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


# ---------------- JPEG: Huffman decoding on the host, dequantisation + ISLOW IDCT on the device ----------------
def _jpeg_cases():
    rng = np.random.default_rng(21)
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (203, 317), dtype=np.uint8), (0, 0), 2.5)
    noisy = rng.integers(0, 256, (64, 80), dtype=np.uint8)
    col = cv2.GaussianBlur(rng.integers(0, 256, (131, 157, 3), dtype=np.uint8), (0, 0), 1.5)
    extreme = np.zeros((40, 56), np.uint8)
    extreme[::2, ::3] = 255                                  # ringing drives the IDCT output past [0, 255]: range limit
    enc = lambda img, *p: bytes(cv2.imencode(".jpg", img, list(p))[1])
    Q, SS, RST = cv2.IMWRITE_JPEG_QUALITY, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_RST_INTERVAL
    return {
        "gray_q95": enc(smooth, Q, 95),
        "gray_q30_odd_size": enc(smooth[:201, :315], Q, 30),
        "gray_noise_q100": enc(noisy, Q, 100),
        "gray_extreme_q50": enc(extreme, Q, 50),
        "gray_restart_7": enc(smooth, Q, 80, RST, 7),
        "color_420": enc(col, Q, 90, SS, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420),
        "color_444": enc(col, Q, 85, SS, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444),
        "color_422_restart": enc(col[:130, :151], Q, 75, SS, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, RST, 3),
        "color_411": enc(col, Q, 60, SS, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411),
        "tiny_1x1": enc(np.full((1, 1), 77, np.uint8), Q, 90),
        "gray_restart_1": enc(smooth[:64, :200], Q, 70, RST, 1),          # a restart marker after every MCU
        "color_420_restart_2": enc(col, Q, 88, SS, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, RST, 2),
        "gray_flat_restart_3": enc(np.full((40, 120), 128, np.uint8), Q, 90, RST, 3),   # 4-bit blocks: intervals shorter than a byte of padding
    }


def _host_luma(data):
    """quantised luma coefficients through the library's host-side Huffman decoder (no device involved)"""
    import ctypes as C
    from fastest_image_pattern_matching_b200 import _lib as L
    lib = L.load()
    buf = np.frombuffer(data, np.uint8)
    w, h, bw, bh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    quant = np.zeros(64, np.uint16)
    err = C.create_string_buffer(256)
    rc = lib.fpm_dbg_jpeg_luma(buf.ctypes.data, buf.size, C.byref(w), C.byref(h), C.byref(bw), C.byref(bh), quant.ctypes.data, None, 0, err, 256)
    if rc != 0:
        raise ValueError(err.value.decode())
    coef = np.zeros((bh.value * bw.value, 8, 8), np.int16)
    rc = lib.fpm_dbg_jpeg_luma(buf.ctypes.data, buf.size, None, None, None, None, None, coef.ctypes.data, coef.size, err, 256)
    assert rc == 0
    return w.value, h.value, bw.value, bh.value, quant.reshape(8, 8), coef


@pytest.mark.parametrize("name", sorted(_jpeg_cases()))
def test_jpeg_host_decoder_and_idct_model_equal_cv2(name):
    """CPU pin: the library's Huffman decoder + the numpy restatement of libjpeg's ISLOW IDCT reproduce cv2.imdecode
    (= the reference's cv::imread(IMREAD_GRAYSCALE)) bit for bit; the device kernel is then checked against the same frames"""
    data = _jpeg_cases()[name]
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
    w, h, bw, bh, quant, coef = _host_luma(data)
    assert (h, w) == want.shape
    px = O.jpeg_idct_islow(coef, quant).reshape(bh, bw, 8, 8).transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)
    assert np.array_equal(px[:h, :w], want)


@pytest.mark.parametrize("name", ["Src6", "Dst10"])
def test_jpeg_host_decoder_on_the_references_own_jpeg_files(name):
    """the two real JPEG files among the reference's Test Images (committed as they are under tests/golden/jpeg by
    make_golden.py): 4096x3000 grayscale baseline and 54x54 YCbCr 4:2:0"""
    import os
    data = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg", name + ".jpg"), "rb").read()
    w, h, bw, bh, quant, coef = _host_luma(data)
    px = O.jpeg_idct_islow(coef, quant).reshape(bh, bw, 8, 8).transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)[:h, :w]
    assert np.array_equal(px, cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE))
    assert np.array_equal(px, get_image(name))               # the decoded fixture every matching test uses


def _parallel_luma(data, shape):
    import ctypes as C
    from fastest_image_pattern_matching_b200 import _lib as L
    lib = L.load()
    buf = np.frombuffer(data, np.uint8)
    out = np.zeros(shape, np.int16)
    passes, err = C.c_int(), C.create_string_buffer(256)
    rc = lib.fpm_dbg_jpeg_luma_parallel(buf.ctypes.data, buf.size, out.ctypes.data, out.size, C.byref(passes), err, 256)
    if rc != 0:
        raise ValueError(err.value.decode())
    return out, passes.value


@pytest.mark.parametrize("name", sorted(_jpeg_cases()) + ["Src6", "Dst10"])
def test_parallel_huffman_decoder_equals_the_sequential_one(name):
    """the device Huffman decoder (self-synchronising sub-sequences, fpm_jpeg_par.cuh) run thread by thread on the CPU:
    the same quantised coefficients as the sequential host decoder, which is pinned against cv2 above"""
    import os
    data = _jpeg_cases()[name] if name in _jpeg_cases() else \
        open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg", name + ".jpg"), "rb").read()
    w, h, bw, bh, quant, coef = _host_luma(data)
    got, passes = _parallel_luma(data, coef.shape)
    assert np.array_equal(got, coef)
    assert 1 <= passes <= 128, passes                    # worst fixture: incompressible noise at quality 100


@pytest.mark.parametrize("seed", range(12))
def test_jpeg_decoders_on_random_encodings(seed):
    """random size / content / quality / chroma subsampling / restart interval / optimised Huffman tables: the sequential host
    decoder + the IDCT model equal cv2.imdecode, and the parallel decoder equals the sequential one"""
    rng = np.random.default_rng(1000 + seed)
    h, w = int(rng.integers(1, 200)), int(rng.integers(1, 260))
    color = bool(rng.integers(0, 2))
    img = rng.integers(0, 256, (h, w, 3) if color else (h, w), dtype=np.uint8)
    if rng.integers(0, 2):
        img = cv2.GaussianBlur(img, (0, 0), float(rng.uniform(0.6, 4.0)))
    params = [cv2.IMWRITE_JPEG_QUALITY, int(rng.integers(5, 101))]
    if color:
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, int(rng.choice([cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
                                                                     cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
                                                                     cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]))]
    if rng.integers(0, 2):
        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(rng.integers(1, 9))]
    if rng.integers(0, 2):
        params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
    data = bytes(cv2.imencode(".jpg", img, params)[1])
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
    gw, gh, bw, bh, quant, coef = _host_luma(data)
    px = O.jpeg_idct_islow(coef, quant).reshape(bh, bw, 8, 8).transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)[:gh, :gw]
    assert np.array_equal(px, want), params
    got, passes = _parallel_luma(data, coef.shape)
    assert np.array_equal(got, coef), params


def test_jpeg_parser_and_decoders_survive_damaged_files():
    """header bytes overwritten, files cut off, scan bytes overwritten: the parser rejects or both decoders (the sequential one
    and the CPU-emulated parallel one) run to the end inside their bounds -- 6 000 such files were run once, 400 stay here"""
    import ctypes as C
    from fastest_image_pattern_matching_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(99)
    cases = _jpeg_cases()
    names = sorted(cases)
    rejected = decoded = 0
    for it in range(400):
        data = bytearray(cases[names[it % len(names)]])
        if it % 3 == 0:
            for pos in rng.integers(2, min(len(data), 700), int(rng.integers(1, 6))):
                data[pos] = int(rng.integers(0, 256))
        elif it % 3 == 1:
            data = data[:int(rng.integers(2, len(data)))]
        else:
            for pos in rng.integers(2, len(data), int(rng.integers(1, 20))):
                data[pos] = int(rng.integers(0, 256))
        buf = np.frombuffer(bytes(data), np.uint8)
        w, h, bw, bh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        quant, err = np.zeros(64, np.uint16), C.create_string_buffer(256)
        rc = lib.fpm_dbg_jpeg_luma(buf.ctypes.data, buf.size, C.byref(w), C.byref(h), C.byref(bw), C.byref(bh), quant.ctypes.data, None, 0, err, 256)
        if rc != 0:
            rejected += 1
            assert err.value                                   # every rejection says why
            continue
        n = bw.value * bh.value * 64
        if n > 20_000_000:
            continue
        coef, passes = np.zeros(n, np.int16), C.c_int()
        lib.fpm_dbg_jpeg_luma_parallel(buf.ctypes.data, buf.size, coef.ctypes.data, coef.size, C.byref(passes), err, 256)
        decoded += 1
    assert rejected > 50 and decoded > 50, (rejected, decoded)


def test_jpeg_host_decoder_rejects_what_it_cannot_decode():
    rng = np.random.default_rng(5)
    img = cv2.GaussianBlur(rng.integers(0, 256, (48, 64), dtype=np.uint8), (0, 0), 2)
    prog = bytes(cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1])
    good = bytes(cv2.imencode(".jpg", img)[1])
    for bad, word in ((prog, "progressive"), (good[:200], "truncated"), (b"\x89PNG", "not a JPEG")):
        with pytest.raises(ValueError) as e:
            _host_luma(bad)
        assert word in str(e.value), (word, str(e.value))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_jpeg_cases()))
def test_gpu_jpeg_ingest_bit_exact(matcher, name):
    data = _jpeg_cases()[name]
    want = O.ingest_image(data)
    w, h = matcher.ingestJpeg(data)
    assert (h, w) == want.shape
    assert np.array_equal(matcher.ingestedPixels(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gray_q95", "gray_noise_q100", "color_420", "color_411", "gray_restart_7", "gray_restart_1",
                                  "color_420_restart_2", "gray_flat_restart_3"])
def test_gpu_jpeg_huffman_on_the_device_and_on_the_host_agree(matcher, name):
    data = _jpeg_cases()[name]
    want = O.ingest_image(data)
    matcher.setJpegDeviceHuffman(True)
    matcher.ingestJpeg(data)
    assert np.array_equal(matcher.ingestedPixels(), want)
    assert matcher.getJpegPasses() > 0                                     # decoded on the device, restart intervals too
    matcher.setJpegDeviceHuffman(False)
    matcher.ingestJpeg(data)
    assert np.array_equal(matcher.ingestedPixels(), want)
    assert matcher.getJpegPasses() == 0
    matcher.setJpegDeviceHuffman(True)


@pytest.mark.gpu
def test_gpu_jpeg_ingest_survives_damaged_scans(matcher):
    """bytes of the entropy-coded segment overwritten or cut off: the device decoder must neither hang nor fault (an out-of-step
    thread only ever produces garbage coefficients inside its bounds); the call returns a frame of the right size or an error,
    and the next good file decodes bit-exactly"""
    from fastest_image_pattern_matching_b200 import FpmError
    rng = np.random.default_rng(17)
    for name in ("gray_q95", "color_420", "gray_restart_7"):
        good = _jpeg_cases()[name]
        want = O.ingest_image(good)
        for trial in range(6):
            bad = bytearray(good)
            lo = len(bad) // 2
            if trial < 4:
                for pos in rng.integers(lo, len(bad) - 2, 12):
                    bad[pos] = int(rng.integers(0, 255))                 # never 0xFF: no accidental markers
            else:
                bad = bad[:lo + int(rng.integers(0, len(bad) - lo - 2))]     # truncated, no EOI
            try:
                w, h = matcher.ingestJpeg(bytes(bad))
                assert (h, w) == want.shape and matcher.ingestedPixels().shape == want.shape
            except FpmError:
                pass
        matcher.ingestJpeg(good)
        assert np.array_equal(matcher.ingestedPixels(), want)


@pytest.mark.gpu
def test_gpu_ingest_image_sniffs_the_signature_and_rejects_progressive(matcher):
    from fastest_image_pattern_matching_b200 import FpmError
    for data in (_jpeg_cases()["color_420"], _cases()["bgr24"]):       # cv::imread picks the decoder by signature, not by suffix
        matcher.ingestImage(data)
        assert np.array_equal(matcher.ingestedPixels(), O.ingest_image(data))
    img = np.zeros((16, 16), np.uint8)
    with pytest.raises(FpmError):
        matcher.ingestJpeg(bytes(cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1]))
    with pytest.raises(FpmError):
        matcher.ingestImage(bytes(cv2.imencode(".png", img)[1]))


@pytest.mark.gpu
def test_gpu_jpeg_ingest_of_the_references_own_jpeg_files(matcher, golden_cases):
    """Src6.jpg (4096x3000 grayscale baseline) and Dst10.jpg (54x54 YCbCr 4:2:0) of the reference's Test Images: the decoded
    fixtures under tests/golden/images came from cv2.imread of the same files (make_golden.py)"""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg")
    for name in ("Src6", "Dst10"):
        path = os.path.join(here, name + ".jpg")
        if not os.path.exists(path):
            pytest.skip("JPEG fixture not present")
        data = open(path, "rb").read()
        w, h = matcher.ingestImage(data)
        got = matcher.ingestedPixels()
        assert np.array_equal(got, O.ingest_image(data))
        assert np.array_equal(got, get_image(name))


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
