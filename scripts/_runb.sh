for ws in 4096 192 96 48; do
FPM_WS_MB=$ws python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch 64 > gpurun_out/b_ws$ws.json 2> gpurun_out/b_ws$ws.err
python - <<PY
import json
d=json.load(open("gpurun_out/b_ws$ws.json"))
print("WS=$ws value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"| e2e",round(d["e2e"]["value"]),"ok",d["targets_found_per_frame_ok"])
print("   "+"  ".join("%s %.3f"%(k.replace("fpm_","").replace("_kernel",""),v["ms_per_step"]) for k,v in d["kernels"].items()))
PY
done
