python -m pytest tests -q -m gpu 2>&1 | tail -2
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none -s 234 -c 40 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log | cut -c1-300
