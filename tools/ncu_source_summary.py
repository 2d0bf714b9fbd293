"""Summarise the SASS source page of an .ncu-rep: per-opcode instruction counts, thread efficiency and stall samples.

  python tools/ncu_source_summary.py gpurun_out/prof.ncu-rep [kernel-index]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blk = '"Kernel Name"' + blocks[1 + which]
rows = list(csv.reader(io.StringIO(blk)))
print(rows[0][1][:100])
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
tot_inst = tot_thr = tot_samp = 0
by_op = defaultdict(lambda: [0, 0, 0])
lines = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ci["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0] + ("." + src.split()[0 if not src.startswith("@") else 1].split(".")[1] if op in ("I2F", "F2F", "MUFU") and "." in src.split()[0 if not src.startswith("@") else 1] else "")
    inst = int(r[ci["Instructions Executed"]]); thr = int(r[ci["Thread Instructions Executed"]]); samp = int(r[ci["# Samples"]])
    tot_inst += inst; tot_thr += thr; tot_samp += samp
    by_op[op][0] += inst; by_op[op][1] += thr; by_op[op][2] += samp
    lines.append((samp, inst, thr, src))
print("total warp-instructions %d  thread-efficiency %.2f/32  samples %d" % (tot_inst, tot_thr / max(1, tot_inst), tot_samp))
print("%-14s %12s %7s %7s %8s" % ("opcode", "warp-inst", "share", "thr/32", "samples"))
for op, (i, t, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:28]:
    print("%-14s %12d %6.1f%% %7.1f %7.1f%%" % (op, i, 100.0 * i / tot_inst, t / max(1, i), 100.0 * s / max(1, tot_samp)))
print("-- top 25 instructions by stall samples")
for samp, inst, thr, src in sorted(lines, key=lambda x: -x[0])[:25]:
    print("%6.2f%% inst %10d thr %5.1f  %s" % (100.0 * samp / max(1, tot_samp), inst, thr / max(1, inst), src[:90]))
