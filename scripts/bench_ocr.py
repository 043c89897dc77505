"""Multi-template read of the M12 fixture (36 glyph templates): GPU handles overlapped on the device
(fpm_match_multi) vs one handle at a time.  Prints one JSON line (not the driver's bench contract; see bench.py for
that).  The text is checked against the committed golden answer (tests/golden/cases.json, written by the oracle in
make_golden.py); the CPU oracle itself is timed by tests/test_ocr.py, not here."""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastest_image_pattern_matching_b200 import GlyphReader  # noqa: E402
import fpm_workloads as synth  # noqa: E402
from fastest_image_pattern_matching_b200.matcher import OCR_LETTERS  # noqa: E402

case = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "cases.json")))["ocr_m12"]
src = synth.load_fixture(case["src"])
tpls = {ch: synth.load_fixture("M12/" + ch) for ch in OCR_LETTERS}
reader = GlyphReader(tpls, **case["params"])
for _ in range(5):
    text, _ = reader.read(src)
multi, serial = [], []
for _ in range(30):
    t0 = time.perf_counter(); reader.read(src); multi.append((time.perf_counter() - t0) * 1e3)
for _ in range(10):
    t0 = time.perf_counter()
    for m in reader.matchers:
        m.match(src)
    serial.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"workload": "ocr_m12: 36 glyph templates x one 925x448 image", "text_ok": text == case["text"],
                  "gpu_multi_p50_ms": statistics.median(multi), "gpu_one_handle_at_a_time_p50_ms": statistics.median(serial),
                  "reads_per_s": 1000.0 / statistics.median(multi)}))
