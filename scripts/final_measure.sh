#!/bin/bash
# One-GPU measurement campaign of a round (run on the GPU box through gpurun): parity tests, the bench line of every
# BASELINE.json configuration, the CPU arm, the latency mode, the side benches, the ncu launch list and full captures.
# Everything lands in gpurun_out/final/.
set -u
O=gpurun_out/final
mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
for w in cfg2 cfg3 cfg4 cfg5; do python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_cfg1_reference_arm.json 2> $O/bench_ref.err
python bench.py --mode latency > $O/lat_n1.json 2> $O/lat_n1.err
python scripts/bench_ocr.py > $O/ocr.json 2> $O/ocr.err
python scripts/bench_ingest.py > $O/ingest_src6_jpeg.json 2> $O/ingest.err
python scripts/latency_breakdown.py cfg1 > $O/latency_breakdown_cfg1.json 2> $O/lb.err
# ncu: launch list of two steps (half-batch split off: one stream, clean per-kernel list), then full captures
FPM_SPLIT=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file $O/launches_batch64.csv python bench.py --device-steps-only --steps 1 --warmup 1 > $O/ncu_list.log 2>&1
python scripts/traffic_from_launches.py $O/launches_batch64.csv cfg1 64 > $O/traffic.json 2> $O/traffic.err
FPM_SPLIT=0 ncu --set full --clock-control none --import-source on -k regex:fpm_pyrdown --launch-skip 6 -c 1 -o $O/ncu_pyrdown_l0 -f python bench.py --device-steps-only --steps 1 --warmup 1 > $O/ncu_pd.log 2>&1
ls -la $O | tail -30
