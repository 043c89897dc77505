python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 1 > gpurun_out/b1.json 2> gpurun_out/b1.err
python - <<PY
import json
d=json.load(open("gpurun_out/b1.json"))
print("batch1 value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"wall",round(d["wall_ms_per_step"],3),"p50",round(d["p50_ms_per_match_batch1"],3), "launches/step", d["gpu_launches"]/20)
ks=d["kernels"]; print("kernel sum ms", round(sum(v["ms_per_step"] for v in ks.values()),3))
print("   "+"  ".join("%s %.3f"%(k.replace("fpm_","").replace("_kernel",""),v["ms_per_step"]) for k,v in ks.items()))
PY
