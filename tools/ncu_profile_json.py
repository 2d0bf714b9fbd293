"""The figures bench.py quotes from an ncu capture, as JSON under profiles/:

  python tools/ncu_profile_json.py <rep> <launch index in the report> <rays traced by that launch> "<command that made it>" <out.json>
"""
import csv
import json
import subprocess
import sys

rep, index, rays, source, out_path = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4], sys.argv[5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2 + index]


def val(name, default=None):
    if name not in hdr:
        return default
    try:
        return float(r[hdr.index(name)].replace(",", ""))
    except ValueError:
        return default


def scaled(name):
    """value in base units (ncu prints Kbyte / Mbyte / Gbyte, usecond / msecond ...)."""
    v = val(name)
    if v is None:
        return None
    u = units[hdr.index(name)].lower()
    for prefix, k in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("msecond", 1e-3), ("usecond", 1e-6), ("nsecond", 1e-9),
                      ("second", 1.0), ("sector", 32.0)):
        if u.startswith(prefix):
            return v * k
    return v * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


t = scaled("gpu__time_duration.sum")
dram = (scaled("dram__bytes_read.sum") or 0.0) + (scaled("dram__bytes_write.sum") or 0.0)
local = sum(scaled(h) or 0.0 for h in ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"))
stalls = {}
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
        try:
            if float(r[i]) > 0.3:
                stalls[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(r[i])
        except ValueError:
            pass
inst = val("smsp__inst_executed.sum")
lts = scaled("lts__t_bytes.sum") or scaled("lts__t_sectors.sum")
doc = {
    "kernel": r[hdr.index("Kernel Name")][:120], "source": source, "rays": rays, "gpu_time_us": t * 1e6 if t else None,
    "dram_bytes_read": scaled("dram__bytes_read.sum"), "dram_bytes_write": scaled("dram__bytes_write.sum"),
    "dram_bytes_per_ray": dram / rays, "dram_gbs": dram / t / 1e9 if t else None,
    "dram_pct_of_peak": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "lts_bytes": lts, "lts_gbs": lts / t / 1e9 if (lts and t) else None,
    "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
    "lane_efficiency": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "warp_instructions_per_ray": inst / rays if inst else None,
    "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers": val("launch__registers_per_thread"),
    "local_memory_bytes_per_ray": local / rays,
    "local_memory_l1_hit_pct": [val("l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"), val("l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct")],
    "stalls": stalls,
}
with open(out_path, "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps(doc))
