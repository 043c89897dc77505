"""Multi-template read of the M12 fixture (36 glyph templates): GPU concurrent handles vs one
handle at a time vs the CPU oracle.  Prints one JSON line (not the driver's bench contract; see bench.py for that)."""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastest_image_pattern_matching_b200 import GlyphReader, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (CPU baseline leg only)

case = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "cases.json")))["ocr_m12"]
src = synth.load_fixture(case["src"])
tpls = {ch: synth.load_fixture("M12/" + ch) for ch in O.OCR_LETTERS}
reader = GlyphReader(tpls, **case["params"])
for _ in range(5):
    text, _ = reader.read(src)
assert text == case["text"]
multi, serial = [], []
for _ in range(30):
    t0 = time.perf_counter(); reader.read(src); multi.append((time.perf_counter() - t0) * 1e3)
for _ in range(10):
    t0 = time.perf_counter()
    for m in reader.matchers:
        m.match(src)
    serial.append((time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter()
cpu_text, _ = O.ocr_read(src, tpls, case["params"])
cpu_ms = (time.perf_counter() - t0) * 1e3
print(json.dumps({"workload": "ocr_m12: 36 glyph templates x one 925x448 image", "text_ok": text == cpu_text == case["text"],
                  "gpu_multi_p50_ms": statistics.median(multi), "gpu_one_handle_at_a_time_p50_ms": statistics.median(serial),
                  "cpu_oracle_ms": cpu_ms, "reads_per_s": 1000.0 / statistics.median(multi)}))
