python -m pytest tests -q -m gpu 2>&1 | tail -4
for tc in 1 0; do
python - <<PY
import subprocess, json, sys
PY
FPM_TC=$tc python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch 64 > gpurun_out/b_tc$tc.json 2> gpurun_out/b_tc$tc.err
python - <<PY
import json
d=json.load(open("gpurun_out/b_tc$tc.json"))
print("TC=$tc value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"| e2e",round(d["e2e"]["value"]),"p50",round(d["p50_ms_per_match_batch1"],3),"ok",d["targets_found_per_frame_ok"])
for k,v in d["kernels"].items(): print("   %-32s %8.3f ms/step  %6.1f us/launch  share %.3f"%(k,v["ms_per_step"],v["avg_launch_us"],v["share"]))
print("   roofline",d["roofline"])
PY
done
