"""Stress the fused tensor-core kernel's window statistics: repeat one shape, report every mismatching cell."""
import sys
import numpy as np
from fastest_image_pattern_matching_b200.matcher import TemplateMatcher

tw, th, ne, reps = [int(a) for a in sys.argv[1:5]]
m = TemplateMatcher()
rng = np.random.default_rng(1)
bad_total = 0
for rep in range(reps):
    tpl = rng.integers(0, 256, (th, tw), dtype=np.uint8)
    rois = rng.integers(0, 256, (ne, th + 6, tw + 6), dtype=np.uint8)
    numer, winS, winQ = m.dbgCorrFused(rois, tpl)
    r64 = rois.astype(np.int64)
    want = np.zeros((ne, 7, 7), np.int64)
    for r in range(7):
        for c in range(7):
            want[:, r, c] = r64[:, r:r + th, c:c + tw].sum(axis=(1, 2))
    bad = np.argwhere(winS != want)
    if len(bad):
        bad_total += 1
        es = sorted(set(bad[:, 0].tolist()))
        print("rep", rep, "bad cells", len(bad), "evals", es[:10])
        edge = m.last_edge_rows
        for e in es[:2]:
            for y in range(th + 6):
                if y < 6 or y >= th:
                    wantrow = [int(r64[e, y, c:c + tw].sum()) for c in range(7)]
                    got = edge[e, y].tolist()
                    if got != wantrow:
                        d = [g - w for g, w in zip(got, wantrow)]
                        inc = [d[c] - d[c - 1] for c in range(1, 7)]          # error of (tl[i] - hd[i]) per i
                        print("  eval", e, "row", y, "delta", d, "per-byte err", inc)
                        for dy in range(-6, 7):
                            yy = y + dy
                            if 0 <= yy < th + 6 and dy != 0:
                                alt_t = [int(r64[e, yy, tw + i]) - int(r64[e, y, tw + i]) for i in range(6)]
                                alt_h = [-(int(r64[e, yy, i]) - int(r64[e, y, i])) for i in range(6)]
                                for nm, alt in (("tail", alt_t), ("head", alt_h)):
                                    hit = sum(1 for i in range(6) if inc[i] != 0 and inc[i] == alt[i])
                                    nz = sum(1 for i in range(6) if inc[i] != 0)
                                    if nz and hit == nz:
                                        print("     explained by", nm, "bytes of row", yy, "(dy=%d)" % dy)
        for e in es[:0]:
            d = (winS[e] - want[e])
            print(" eval", e, "delta S (rows r, cols c):\n", d)
            # candidate bytes
            print("  head bytes rows0..:", rois[e][:, :6].tolist()[:3], " tail:", rois[e][:, tw:tw + 6].tolist()[:3])
print("reps with mismatches:", bad_total, "of", reps)
