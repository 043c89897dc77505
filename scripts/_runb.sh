for cfg in "4 0" "8 0" "64 4" "64 8" "64 16"; do
set -- $cfg
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch $1 --h2d-chunk $2 > gpurun_out/b.json 2> gpurun_out/b.err
python - <<PY
import json
d=json.load(open("gpurun_out/b.json"))
print("batch $1 chunk $2: value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"| e2e",round(d["e2e"]["value"]),"e2e ms/step",round(d["e2e"]["ms_per_step"],3),"pcie bound",round(d["e2e"]["pcie_bound_images_per_s"]),"ok",d["targets_found_per_frame_ok"])
PY
done
