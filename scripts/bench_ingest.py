"""JPEG ingest of the reference's Src6.jpg (4096x3000 grayscale baseline, 1.5 MB): p50 of fpm_ingest_jpeg with the Huffman decoding
on the device / on the host, cv2.imdecode on the host beside it, and ingest + match (cfg3 parameters) from the file bytes.
    python scripts/bench_ingest.py"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import bench  # noqa: E402
import fpm_workloads as synth  # noqa: E402
from fastest_image_pattern_matching_b200 import TemplateMatcher  # noqa: E402


def p50(fn, n=60, skip=5):
    ts = []
    for _ in range(n + skip):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return round(statistics.median(ts[skip:]), 4)


def main():
    data = open(os.path.join(ROOT, "tests", "golden", "jpeg", "Src6.jpg"), "rb").read()
    buf = np.frombuffer(data, np.uint8)
    m = TemplateMatcher(0, result_capacity=256)
    wl = bench.WORKLOADS["cfg3"]
    bench.configure(m, wl)
    assert m.learnPattern(synth.load_fixture("Dst6"))
    out = {"file": "Src6.jpg", "file_bytes": len(data), "pixels": "4096x3000"}
    m.setJpegDeviceHuffman(True)
    out["ingest_device_huffman_p50_ms"] = p50(lambda: m.ingestJpeg(buf))
    out["sync_passes"] = m.getJpegPasses()
    assert np.array_equal(m.ingestedPixels(), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE))
    out["ingest_and_match_p50_ms"] = p50(lambda: (m.ingestJpeg(buf), m.matchIngested()))
    out["targets"] = len(m.matchIngested())
    m.setJpegDeviceHuffman(False)
    out["ingest_host_huffman_p50_ms"] = p50(lambda: m.ingestJpeg(buf), n=15, skip=2)
    out["cv2_imdecode_host_p50_ms"] = p50(lambda: cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE), n=15, skip=2)
    frame = cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)
    out["match_pageable_pixels_p50_ms"] = p50(lambda: m.match(frame))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
