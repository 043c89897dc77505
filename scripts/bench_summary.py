import json,sys
for f in sys.argv[1:]:
    d=json.load(open(f))
    print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["p50_ms_per_match_batch1"],3), round(d["e2e"]["value"]), d["targets_found_per_frame_ok"])
    print("   ", {k: round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
