for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
python bench.py --steps 6 --warmup 3 --workload $wl --cpu-images 2 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$wl.json"))
    cb=d.get("cpu_baseline") or {}
    print("$wl batch",d["config"]["batch_per_gpu"],"value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"| e2e",round(d["e2e"]["value"],1),"pcie",round(d["e2e"]["pcie_bound_images_per_s"],1),"| p50",round(d["p50_ms_per_match_batch1"],3),"| cpu",round(cb.get("value",0),2),"img/s on",cb.get("cores"),"cores | ok",d["targets_found_per_frame_ok"])
    print("   "+"  ".join("%s %.3f"%(k.replace("fpm_","").replace("_kernel",""),v["ms_per_step"]) for k,v in d["kernels"].items()))
except Exception as e:
    print("$wl FAILED", e); print(open("gpurun_out/bench_$wl.err").read()[-1500:])
PY
done
