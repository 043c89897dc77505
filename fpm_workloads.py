"""Seeded synthetic workloads of the BASELINE.json configs for bench.py, tests/ and smoke().

Bench / test infrastructure (it reads the fixtures under tests/golden/): deliberately NOT part of the
fastest_image_pattern_matching_b200 package, which holds only what the matching path needs."""
from __future__ import annotations

import math
import os

import numpy as np

GOLDEN_IMAGES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "images")


def load_fixture(name: str) -> np.ndarray:
    """Grayscale fixture decoded once from the reference's Test Images (tests/golden/make_golden.py)."""
    import cv2
    p = os.path.join(GOLDEN_IMAGES, name + ".png")
    img = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(p)
    return img


def background(w: int, h: int, seed: int, sigma: float = 3.0) -> np.ndarray:
    import cv2
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    return cv2.GaussianBlur(noise, (0, 0), sigma)


def paste_rotated(dst: np.ndarray, tpl: np.ndarray, cx: float, cy: float, angle_deg: float, fast: bool = False) -> None:
    """Paste `tpl` rotated by angle_deg (OpenCV sign) with its centre at (cx, cy), in place.
    fast: warp only the bounding box of the rotated template (bench frames; same picture up to fixed-point rounding --
    the golden fixtures were made with the full-frame warp and keep it)."""
    import cv2
    th, tw = tpl.shape
    if fast:
        r = int(math.ceil(math.hypot(tw, th) / 2)) + 2
        x0, y0 = max(0, int(cx) - r), max(0, int(cy) - r)
        x1, y1 = min(dst.shape[1], int(cx) + r + 1), min(dst.shape[0], int(cy) + r + 1)
        paste_rotated(dst[y0:y1, x0:x1], tpl, cx - x0, cy - y0, angle_deg)
        return
    m = cv2.getRotationMatrix2D(((tw - 1) / 2.0, (th - 1) / 2.0), angle_deg, 1.0)
    m[0, 2] += cx - (tw - 1) / 2.0
    m[1, 2] += cy - (th - 1) / 2.0
    size = (dst.shape[1], dst.shape[0])
    warped = cv2.warpAffine(tpl, m, size, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    mask = cv2.warpAffine(np.full_like(tpl, 255), m, size, flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT,
                          borderValue=0)
    dst[mask > 0] = warped[mask > 0]


# README.md:47-49 poses of the Cognex comparison case (centre x, centre y, angle in OpenCV sign)
CFG1_POSES = [(1725.857, 1045.433, -0.046), (2662.869, 1537.446, 119.979), (1768.936, 2098.494, -120.150)]


def cfg1_source(seed: int = 7, tpl: np.ndarray | None = None, jitter: bool = False, fast: bool = False) -> np.ndarray:
    """4024x3036 synthetic stand-in for the missing Src7.bmp: blurred noise + Dst7 at the README poses."""
    if tpl is None:
        tpl = load_fixture("Dst7")
    img = background(4024, 3036, seed, 3.0)
    rng = np.random.default_rng(seed + 1000)
    for (cx, cy, a) in CFG1_POSES:
        if jitter:
            cx += float(rng.uniform(-40, 40)); cy += float(rng.uniform(-40, 40)); a += float(rng.uniform(-25, 25))
        paste_rotated(img, tpl, cx, cy, a, fast)
    return img


def cfg2_source(seed: int = 10, tpl: np.ndarray | None = None) -> np.ndarray:
    """3648x3648 synthetic stand-in for Src10.bmp: Dst10 tiled on a 150 px grid (576 copies)."""
    if tpl is None:
        tpl = load_fixture("Dst10")
    img = background(3648, 3648, seed, 2.0)
    th, tw = tpl.shape
    for gy in range(24):
        for gx in range(24):
            y, x = 40 + gy * 150, 40 + gx * 150
            img[y:y + th, x:x + tw] = tpl
    return img


def synth_template(size: int, seed: int = 4) -> np.ndarray:
    """cfg4/cfg5 template: normalised blur-noise + one filled circle + one filled rectangle."""
    import cv2
    t = background(size, size, seed, 4.0).astype(np.float32)
    t = (t - t.min()) / max(float(t.max() - t.min()), 1.0) * 255.0
    t = t.astype(np.uint8)
    cv2.circle(t, (int(size * 0.3), int(size * 0.35)), int(size * 0.18), 255, -1)
    cv2.rectangle(t, (int(size * 0.55), int(size * 0.5)), (int(size * 0.9), int(size * 0.8)), 0, -1)
    return t


def synth_frame(w: int, h: int, tpl: np.ndarray, seed: int, k: int = 4, fast: bool = False) -> np.ndarray:
    """cfg4/cfg5 frame: blurred noise + k non-overlapping rotated instances of tpl."""
    img = background(w, h, seed, 3.0)
    rng = np.random.default_rng(seed + 5000)
    diag = math.hypot(*tpl.shape)
    centres = []
    tries = 0
    while len(centres) < k and tries < 10000:
        tries += 1
        cx = float(rng.uniform(diag / 2 + 2, w - diag / 2 - 2))
        cy = float(rng.uniform(diag / 2 + 2, h - diag / 2 - 2))
        if all(math.hypot(cx - x, cy - y) >= diag for x, y, _ in centres):
            centres.append((cx, cy, float(rng.uniform(-180, 180))))
    for cx, cy, a in centres:
        paste_rotated(img, tpl, cx, cy, a, fast)
    return img
