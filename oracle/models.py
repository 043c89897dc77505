"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Plain numpy models of the OpenCV routines the CUDA kernels restate (the arithmetic lives in the
un-vendored third-party dependency OpenCV).  Each model is pinned
bit-exactly against cv2 4.13 in tests/test_oracle.py, so a kernel that matches cv2 on the GPU box
and a model that matches cv2 here describe the same arithmetic.
"""
import math

import numpy as np


def _reflect101(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * n - 2 - i
    return i


def pyrdown(img):
    """cv::pyrDown u8: 5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8, out ((w+1)/2, (h+1)/2)."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int64)
    rows = np.array([[_reflect101(2 * y + d, h) for d in range(-2, 3)] for y in range(oh)])
    cols = np.array([[_reflect101(2 * x + d, w) for d in range(-2, 3)] for x in range(ow)])
    a = img.astype(np.int64)
    tmp = (a[:, cols] * k).sum(axis=2)                    # h x ow
    out = (tmp[rows, :] * k[None, :, None]).sum(axis=1)   # oh x ow
    return ((out + 128) >> 8).astype(np.uint8)


def rotation_matrix(cx, cy, angle_deg):
    """cv::getRotationMatrix2D(center, angle, 1)."""
    a = angle_deg * (math.pi / 180)
    alpha, beta = math.cos(a), math.sin(a)
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]])


def _cvround(v):
    return np.rint(v).astype(np.int64)                     # round-half-to-even like cvRound/lrint


def warp_affine(img, m, dsize, border):
    """cv::warpAffine u8 C1 INTER_LINEAR BORDER_CONSTANT: fixed-point path (AB_BITS 10, INTER_BITS 5)."""
    M = np.array(m, np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    dw, dh = dsize
    h, w = img.shape
    xs = np.arange(dw, dtype=np.float64)
    ys = np.arange(dh, dtype=np.float64)
    adelta = _cvround(M[0] * xs * 1024)
    bdelta = _cvround(M[3] * xs * 1024)
    X0 = _cvround((M[1] * ys + M[2]) * 1024) + 16
    Y0 = _cvround((M[4] * ys + M[5]) * 1024) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy, ax, ay = X >> 5, Y >> 5, X & 31, Y & 31
    src = img.astype(np.int64)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return np.where(ok, v, border)

    v = ((32 - ax) * (32 - ay) * 32 * tap(sy, sx) + ax * (32 - ay) * 32 * tap(sy, sx + 1) +
         (32 - ax) * ay * 32 * tap(sy + 1, sx) + ax * ay * 32 * tap(sy + 1, sx + 1))
    return ((v + 16384) >> 15).astype(np.uint8)


def mean_stddev(img):
    """cv::meanStdDev u8 single channel."""
    a = img.astype(np.int64)
    n = a.size
    S, Q = int(a.sum()), int((a * a).sum())
    scale = 1.0 / n
    mean = S * scale
    var = max(Q * scale - mean * mean, 0.0)
    return mean, math.sqrt(var)


def paint(mat, x, y, w, h, value=-1.0):
    """cv::rectangle(mat, Rect(x,y,w,h), value, FILLED): [x, x+w) x [y, y+h) clipped; empty Rect paints nothing."""
    if w <= 0 or h <= 0:
        return
    x0, y0 = max(x, 0), max(y, 0)
    x1, y1 = min(x + w, mat.shape[1]), min(y + h, mat.shape[0])
    if x1 > x0 and y1 > y0:
        mat[y0:y1, x0:x1] = value
