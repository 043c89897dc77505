"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle_cpu_match.so ("CPU Baseline A": the C++ restatement of TemplateMatcher::match built
with the reference's Release flags, oracle/cpu_match.cpp).  Used by tests/ (checked against the Python oracle) and by
bench.py's CPU arms; never by the product path."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
PARAMS = {"max_pos": 0, "max_overlap": 1, "score": 2, "tolerance_angle": 3, "min_reduce_area": 4, "use_simd": 5, "sub_pixel": 6}


def load(build: bool = True):
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle_cpu_match.so")
    if not os.path.exists(path) and build:
        subprocess.run(["make", "-C", _HERE, "liboracle_cpu_match.so"], check=True, capture_output=True)
    lib = C.CDLL(path)
    lib.cpum_create.restype = C.c_void_p
    lib.cpum_destroy.argtypes = [C.c_void_p]
    lib.cpum_set.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.cpum_learn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.cpum_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.cpum_last_ms.argtypes = [C.c_void_p]
    lib.cpum_last_ms.restype = C.c_double
    _LIB = lib
    return lib


class CpuMatcher:
    """same surface as oracle.OracleMatcher's basic use: attributes max_pos/score/..., learn_pattern, match -> [n, 12] rows
    (score, angle, cx, cy, lt, rt, rb, lb)"""

    def __init__(self):
        self.lib = load()
        self.h = self.lib.cpum_create()
        self.max_pos, self.max_overlap, self.score, self.tolerance_angle = 70, 0.0, 0.7, 0.0
        self.min_reduce_area, self.use_simd, self.sub_pixel = 256, True, False

    def __del__(self):
        try:
            self.lib.cpum_destroy(self.h)
        except Exception:
            pass

    def _push(self):
        for k, i in PARAMS.items():
            self.lib.cpum_set(self.h, i, float(getattr(self, k)))

    def learn_pattern(self, tpl):
        self._push()
        t = np.ascontiguousarray(tpl, np.uint8)
        self._tpl = t
        return self.lib.cpum_learn(self.h, t.ctypes.data, t.shape[1], t.shape[0]) == 0

    def match(self, src, cap=4096):
        self._push()
        s = np.ascontiguousarray(src, np.uint8)
        out = np.zeros((cap, 12), np.float64)
        n = self.lib.cpum_match(self.h, s.ctypes.data, s.shape[1], s.shape[0], out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("MinReduceArea changed after learnPattern")
        return out[:min(n, cap)]

    def last_ms(self):
        return self.lib.cpum_last_ms(self.h)
