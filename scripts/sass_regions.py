"""Instructions executed and stall samples per block of SASS lines of one launch in an .ncu-rep
(ncu -i REP --page source --print-source=sass --csv).  Usage: sass_regions.py REP [kernel-regex] [lines-per-block]"""
import csv, subprocess, sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass", "--launch-count", "1"]
if kre:
    cmd += ["-k", "regex:" + kre]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
name = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else "?"
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


data = []
seen = set()
i_addr = hdr.index("Address")
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    if r[i_addr] in seen:                              # the CSV repeats the listing
        break
    seen.add(r[i_addr])
    data.append(r)
tot, ts = sum(I(r[i_ex]) for r in data), sum(I(r[i_s]) for r in data)
print("%s: %d SASS lines, %.1f M warp instructions, %d stall samples" % (name[:60], len(data), tot / 1e6, ts))
for k in range(0, len(data), blk):
    seg = data[k:k + blk]
    ex, sm = sum(I(r[i_ex]) for r in seg), sum(I(r[i_s]) for r in seg)
    if ex < 0.005 * tot and sm < 0.005 * ts:
        continue
    ops = {}
    for r in seg:
        t = r[i_src].strip().split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        ops[op] = ops.get(op, 0) + I(r[i_ex])
    top = ", ".join("%s %.1fM" % (o, v / 1e6) for o, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print("  lines %4d-%4d: %5.1f %% of instructions, %5.1f %% of samples | %s" % (k, k + len(seg) - 1, 100.0 * ex / tot, 100.0 * sm / max(ts, 1), top))
