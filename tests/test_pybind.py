"""The pybind11 module (the reference README's "C++ .so with Pybind11 for Python") builds, imports, fails loudly
without a GPU and, on the GPU box, returns the golden Src8 targets."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def pymod(fpm_built):
    from fastest_image_pattern_matching_b200 import _build
    _build.build_pybind()
    sys.path.insert(0, os.path.join(ROOT, "fastest_image_pattern_matching_b200"))
    import fpm_b200_pybind
    return fpm_b200_pybind


def test_pybind_builds_and_has_the_reference_surface(pymod):
    import torch
    for name in ("setMaxPositions", "setMaxOverlap", "setScore", "setToleranceAngle", "setMinReduceArea", "setUseSIMD",
                 "setSubPixelEstimation", "learnPattern", "match", "getLastExecutionTime", "isPatternLearned", "clearPattern"):
        assert hasattr(pymod.TemplateMatcher, name)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            pymod.TemplateMatcher(0)


@pytest.mark.gpu
def test_pybind_match(pymod, golden_cases):
    import fpm_workloads as synth
    m = pymod.TemplateMatcher(0)
    m.setMaxPositions(5); m.setScore(0.8); m.setToleranceAngle(180); m.setMaxOverlap(0.8)
    assert m.learnPattern(synth.load_fixture("Dst8"))
    res = m.match(synth.load_fixture("Src8"))
    want = golden_cases["src8"]["results"]
    assert len(res) == len(want) == 3
    for r, w in zip(res, want):
        assert abs(r.dMatchScore - w["score"]) <= 1e-4 and abs(r.dMatchedAngle - w["angle"]) <= 0.01
        assert abs(r.ptCenter[0] - w["cx"]) <= 0.05 and abs(r.ptCenter[1] - w["cy"]) <= 0.05
    assert m.match(np.zeros((0, 0), np.uint8)) == []
