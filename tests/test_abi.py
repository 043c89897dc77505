"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/fpm_b200.h declares; the host-side geometry (shared host/device code) agrees with cv2; the
product path fails loudly without a GPU instead of falling back to a CPU implementation."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fpm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fpm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(fpm_built):
    import ctypes
    from fastest_image_pattern_matching_b200 import _lib
    lib = ctypes.CDLL(fpm_built)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in _lib.load().fpm_version()


def test_sass_is_sm100a(fpm_built):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", fpm_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(fpm_built):
    import torch
    from fastest_image_pattern_matching_b200 import TemplateMatcher, FpmError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(FpmError):
        TemplateMatcher(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fastest_image_pattern_matching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


def test_rotated_rect_geometry_matches_cv2(fpm_built):
    """fpm_geometry.cuh (host build of the device NMS code) vs cv2 on reference-style rectangles."""
    import cv2
    from fastest_image_pattern_matching_b200.matcher import rrect_from3_host, rrect_overlap_host
    from oracle.oracle import MatchParameter, OracleMatcher
    rng = np.random.default_rng(11)
    f32 = np.float32
    bad_type = bad_dec = 0
    for it in range(3000):
        w, h = (int(rng.integers(20, 800)), int(rng.integers(20, 600))) if it % 2 else (762, 521)
        a1 = float(rng.uniform(-180, 180))
        a2 = a1 + float(rng.normal(0, 30)) if it % 3 else float(rng.uniform(-180, 180))
        p1 = (float(f32(rng.uniform(0, 3000))), float(f32(rng.uniform(0, 3000))))
        d = rng.uniform(0, 1.2) * np.hypot(w, h)
        th = rng.uniform(0, 2 * np.pi)
        p2 = (float(f32(p1[0] + d * np.cos(th))), float(f32(p1[1] + d * np.sin(th))))
        rr = []
        for p, a in ((p1, a1), (p2, a2)):
            lt, rt, lb, rb = OracleMatcher._corners(MatchParameter(pt=p, angle=a), w, h)
            mine = rrect_from3_host(lt, rt, rb)
            try:
                c = cv2.RotatedRect((float(lt[0]), float(lt[1])), (float(rt[0]), float(rt[1])), (float(rb[0]), float(rb[1])))
                ref = (c.center[0], c.center[1], c.size[0], c.size[1], c.angle)
                assert np.array_equal(np.array(mine, f32), np.array(ref, f32))
            except cv2.error:
                pass
            rr.append(mine)
        t1 = ((rr[0][0], rr[0][1]), (rr[0][2], rr[0][3]), rr[0][4])
        t2 = ((rr[1][0], rr[1][1]), (rr[1][2], rr[1][3]), rr[1][4])
        typ, inter = cv2.rotatedRectangleIntersection(t1, t2)
        for mo in (0.0, 0.1, 0.5, 0.8):
            dec, mt, ratio = rrect_overlap_host(rr[0], rr[1], mo)
            if typ == 0:
                od = 0
            elif typ == 2:
                od = 1
            elif inter is None or len(inter) < 3:
                od = 0
            else:
                pts = OracleMatcher._sort_pt_with_center([(f32(p[0][0]), f32(p[0][1])) for p in inter])
                area = cv2.contourArea(np.array(pts, f32).reshape(-1, 1, 2))
                od = 1 if area / float(f32(rr[0][2]) * f32(rr[0][3])) > mo else 0
            bad_type += mt != typ
            bad_dec += od != dec
    assert bad_type == 0 and bad_dec == 0


def test_rotated_rect_geometry_degenerate_contacts(fpm_built):
    """axis-aligned / corner- and edge-touching / identical / contained rectangles (all-ties score maps
    produce exactly these): the delete decision must agree with cv2 for every pair."""
    import cv2
    from fastest_image_pattern_matching_b200.matcher import rrect_from3_host, rrect_overlap_host
    from oracle.oracle import MatchParameter, OracleMatcher
    f32 = np.float32

    def rect_of(pt, angle, w, h):
        lt, rt, lb, rb = OracleMatcher._corners(MatchParameter(pt=pt, angle=angle), w, h)
        return rrect_from3_host(lt, rt, rb)

    def decisions(r1, r2):
        t1 = ((r1[0], r1[1]), (r1[2], r1[3]), r1[4])
        t2 = ((r2[0], r2[1]), (r2[2], r2[3]), r2[4])
        typ, inter = cv2.rotatedRectangleIntersection(t1, t2)
        out = []
        for mo in (0.0, 0.3, 0.8):
            dec, mt, _ = rrect_overlap_host(r1, r2, mo)
            if typ == 0:
                od = 0
            elif typ == 2:
                od = 1
            elif inter is None or len(inter) < 3:
                od = 0
            else:
                pts = OracleMatcher._sort_pt_with_center([(f32(p[0][0]), f32(p[0][1])) for p in inter])
                area = cv2.contourArea(np.array(pts, f32).reshape(-1, 1, 2))
                od = 1 if area / float(f32(r1[2]) * f32(r1[3])) > mo else 0
            out.append(od == dec)
        return all(out)

    w, h = 24, 20
    angles = [0.0, 90.0, 180.0, -90.0, 270.0, 45.0, 0.5, -180.0, 30.0]
    bad = 0
    for a1 in angles:
        for a2 in angles:
            for dx in [-48, -24, -12, 0, 12, 24, 48, 23, 25, 1, 0.5]:
                for dy in [-40, -20, -10, 0, 10, 20, 40, 19, 21, 1]:
                    bad += not decisions(rect_of((50.0, 60.0), a1, w, h), rect_of((50.0 + dx, 60.0 + dy), a2, w, h))
    for s in [0.5, 0.9, 1.0, 1.1, 2.0]:
        for a in angles:
            r1 = rect_of((100.0, 100.0), a, 100, 80)
            r2 = (r1[0], r1[1], r1[2] * s, r1[3] * s, r1[4])
            bad += not decisions(r1, r2)
            bad += not decisions(r2, r1)
    assert bad == 0


def test_synthetic_generators_are_deterministic(golden_cases):
    import hashlib
    import fpm_workloads as synth
    a = synth.cfg1_source()
    assert a.shape == (3036, 4024)
    assert hashlib.sha256(a.tobytes()).hexdigest()[:16] == golden_cases["cfg1_synth"]["src_sha"]
    t = synth.synth_template(64, 4)
    f1, f2 = synth.synth_frame(512, 384, t, 3, 2), synth.synth_frame(512, 384, t, 3, 2)
    assert np.array_equal(f1, f2)
