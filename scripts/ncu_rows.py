"""Print selected raw metrics for every launch of an .ncu-rep file (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "?")[:40], {k: (d[k], units[hdr.index(k)]) for k in want if k in d})
