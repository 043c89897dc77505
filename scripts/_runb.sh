python -m pytest tests -q -m gpu 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err
python - <<PY
import json
d=json.load(open("gpurun_out/b.json"))
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"| e2e",round(d["e2e"]["value"]),"p50",round(d["p50_ms_per_match_batch1"],3),"ok",d["targets_found_per_frame_ok"])
print("   "+"  ".join("%s %.3f"%(k.replace("fpm_","").replace("_kernel",""),v["ms_per_step"]) for k,v in d["kernels"].items()))
print(d["hbm_kernels"]["fpm_pyrdown_kernel"])
PY
