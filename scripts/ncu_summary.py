"""Key metrics of the launches in an .ncu-rep (ncu -i REP --page raw --csv), one block per launch.
Usage: ncu_summary.py REP [> profiles/rNN_ncu_<what>_summary.txt]"""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:70], d.get("Grid Size", "?"), d.get("Block Size", "?")))
    for k in WANT:
        if k in d:
            print("   %-72s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = sorted(((float(v), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                 if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
    print("   stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (k, v) for v, k in st[:7]))
