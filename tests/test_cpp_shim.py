"""The header-only C++ `TemplateMatcher` shim (include/fpm_template_matcher.hpp) compiles and links
against the C-ABI library; on a GPU box the compiled program must find the 3 Src8 targets."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PROG = r'''
#include "fpm_template_matcher.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>
int main(int argc, char** argv) {
    if (argc < 7) return 2;
    int tw = atoi(argv[2]), th = atoi(argv[3]), sw = atoi(argv[5]), sh = atoi(argv[6]);
    std::vector<unsigned char> t((size_t)tw * th), s((size_t)sw * sh);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(t.data(), 1, t.size(), f) != t.size()) return 3; fclose(f);
    f = fopen(argv[4], "rb"); if (!f || fread(s.data(), 1, s.size(), f) != s.size()) return 3; fclose(f);
    try {
        fpm::TemplateMatcher m(0);
        m.setMaxPositions(5); m.setScore(0.8); m.setToleranceAngle(180); m.setMaxOverlap(0.8);
        if (!m.learnPattern(t.data(), tw, th, tw)) return 4;
        auto r = m.match(s.data(), sw, sh, sw);
        printf("%d\n", (int)r.size());
        for (auto& x : r) printf("%.6f %.4f %.3f %.3f\n", x.dMatchScore, x.dMatchedAngle, x.ptCenter.x, x.ptCenter.y);
    } catch (const std::exception& e) { printf("EXC %s\n", e.what()); return 5; }
    return 0;
}
'''


def _build(tmp_path, fpm_built):
    src = tmp_path / "shim_test.cpp"
    src.write_text(PROG)
    exe = tmp_path / "shim_test"
    libdir = os.path.dirname(fpm_built)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libfpm_b200.so", "-Wl,-rpath," + libdir], check=True, capture_output=True)
    return exe


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(tmp_path, fpm_built):
    import torch
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = _build(tmp_path, fpm_built)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    (tmp_path / "a.raw").write_bytes(bytes(16))
    out = subprocess.run([str(exe), str(tmp_path / "a.raw"), "4", "4", str(tmp_path / "a.raw"), "4", "4"], capture_output=True, text=True)
    assert out.returncode == 5 and "no usable CUDA device" in out.stdout      # no CPU fallback


@pytest.mark.gpu
def test_cpp_shim_matches(tmp_path, fpm_built, golden_cases):
    import fpm_workloads as synth
    exe = _build(tmp_path, fpm_built)
    t, s = synth.load_fixture("Dst8"), synth.load_fixture("Src8")
    (tmp_path / "t.raw").write_bytes(t.tobytes())
    (tmp_path / "s.raw").write_bytes(s.tobytes())
    out = subprocess.run([str(exe), str(tmp_path / "t.raw"), str(t.shape[1]), str(t.shape[0]), str(tmp_path / "s.raw"),
                          str(s.shape[1]), str(s.shape[0])], capture_output=True, text=True, check=True).stdout.split("\n")
    assert int(out[0]) == 3
    want = golden_cases["src8"]["results"]
    for line, w in zip(out[1:4], want):
        sc, ang, cx, cy = [float(v) for v in line.split()]
        assert abs(sc - w["score"]) <= 1e-4 and abs(ang - w["angle"]) <= 0.01 and abs(cx - w["cx"]) <= 0.05 and abs(cy - w["cy"]) <= 0.05
