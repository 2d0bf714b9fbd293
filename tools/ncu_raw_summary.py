"""Key raw metrics of every launch in an .ncu-rep as a markdown table:  python tools/ncu_raw_summary.py rep [rep...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("\n### `%s` (%s, launch id %s)\n" % (r[hdr.index("Kernel Name")][:80], rep.split("/")[-1], r[hdr.index("ID")]))
        print("| metric | unit | value |\n|---|---|---|")
        for w in WANT:
            if w in hdr:
                print("| %s | %s | %s |" % (w, units[hdr.index(w)], r[hdr.index(w)]))
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    if float(r[i]) > 0.3:
                        print("| %s | inst | %s |" % (h.replace("smsp__average_warps_issue_stalled_", "stall: ").replace("_per_issue_active.ratio", ""), r[i]))
                except ValueError:
                    pass
