"""CPU emulation of the pyramid kernel (fpm_pyrdown.cuh): the kernel's __host__ __device__ phase functions run thread by
thread between the barriers (tests/pyrdown_emulate.cu) and must reproduce cv2.pyrDown bit for bit -- one level, and the two
levels of one launch -- for every staging granularity (16 / 8 / 4-byte cp.async, byte loads).  This is the check the
GPU-less container can run; the device path (shuffles instead of halo loads) is covered by tests/test_gpu_parity.py."""
import ctypes
import os
import shutil
import subprocess

import cv2
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pyrdown_emulate.cu")
HDR = os.path.join(HERE, "..", "fastest_image_pattern_matching_b200", "csrc", "fpm_pyrdown.cuh")
LIB = os.path.join(HERE, "_bin", "libpyrdown_emulate.so")


@pytest.fixture(scope="module")
def emu():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                        "-o", LIB, SRC], check=True, capture_output=True)
    lib = ctypes.CDLL(LIB)
    vp, i, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    lib.pd_emulate.argtypes = [vp, i, i, i, ll, i, i, vp, i, ll, vp, i, ll]

    def run(img, vec=16, two=True):
        img = np.ascontiguousarray(img)
        if img.ndim == 2:
            img = img[None]
        batch, h, w = img.shape
        mis = {16: 0, 8: 8, 4: 4, 1: 3}[vec]
        pitch = (w + 127) // 128 * 128 + mis
        stride = pitch * h
        raw = np.full(batch * stride + 256, 0xEE, np.uint8)
        base = (-raw.ctypes.data) % 128 + mis
        buf = raw[base:base + batch * stride].reshape(batch, h, pitch)
        buf[:, :, :w] = img
        w1, h1 = (w + 1) // 2, (h + 1) // 2
        w2, h2 = (w1 + 1) // 2, (h1 + 1) // 2
        p1, p2 = (w1 + 127) // 128 * 128, (w2 + 127) // 128 * 128
        d1 = np.full((batch, h1, p1), 0x77, np.uint8)
        d2 = np.full((batch, h2, p2), 0x77, np.uint8)
        lib.pd_emulate(buf.ctypes.data, w, h, pitch, stride, batch, vec, d1.ctypes.data, p1, h1 * p1,
                       d2.ctypes.data if two else None, p2, h2 * p2)
        assert (d1[:, :, w1:] == 0x77).all() and (d2[:, :, w2:] == 0x77).all(), "row padding was written"
        return d1[:, :, :w1], (d2[:, :, :w2] if two else None)

    return run


SHAPES = [(7, 9), (1, 1), (1, 5), (2, 2), (5, 3), (3, 300), (101, 333), (64, 47), (259, 517), (128, 256), (129, 257), (130, 258),
          (131, 261), (257, 1030)]


@pytest.mark.parametrize("shape", SHAPES)
def test_emulated_kernel_equals_cv2(emu, shape):
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    want1 = cv2.pyrDown(img)
    want2 = cv2.pyrDown(want1)
    for vec in (16, 8, 4, 1):
        got1, got2 = emu(img, vec, True)
        assert np.array_equal(got1[0], want1), (shape, vec, "level 1 of the two-level launch")
        assert np.array_equal(got2[0], want2), (shape, vec, "level 2 of the two-level launch")
        one, _ = emu(img, vec, False)
        assert np.array_equal(one[0], want1), (shape, vec, "one-level launch")


def test_emulated_kernel_batch_and_saturated_pixels(emu):
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (3, 70, 150), dtype=np.uint8)
    imgs[1] = 255                                         # largest packed 16-bit sums: 16 * 4080 + 128 must not carry
    imgs[2, ::2] = 255
    got1, got2 = emu(imgs, 16, True)
    for b in range(3):
        w1 = cv2.pyrDown(imgs[b])
        assert np.array_equal(got1[b], w1) and np.array_equal(got2[b], cv2.pyrDown(w1))
