"""Per CUDA source line: warp-stall samples, instructions, shared wavefronts of one kernel of an .ncu-rep
(needs -lineinfo and --import-source on).  usage: ncu_lines.py report.ncu-rep [top] [kernel regex] [matching launches to skip]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = sys.argv[3] if len(sys.argv) > 3 else None           # regex of the kernel name; the first matching launch is shown
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]
if kern:
    cmd += ["--kernel-name", "regex:" + kern, "--launch-count", "1"]
    if len(sys.argv) > 4:                                    # skip that many matching launches first
        cmd += ["--launch-skip", sys.argv[4]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
c = {h: i for i, h in enumerate(hdr)}
samp, inst = c["# Samples"], c["Instructions Executed"]
wav, l1req = c.get("L1 Wavefronts Shared"), c.get("L1 Tag Requests Global")          # absent when the kernel uses no shared memory
lines = []
for r in rows[hi + 1:]:
    if len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":          # a CUDA line (its SASS rows follow)
        lines.append((int(r[0]), r[1].strip(), int(r[samp] or 0), int(r[inst] or 0), int(r[wav] or 0) if wav is not None else 0, int(r[l1req] or 0) if l1req is not None else 0))
ts, ti, tw, tg = (sum(x[i] for x in lines) for i in (2, 3, 4, 5))
print(f"{len(lines)} lines; samples {ts}, warp instructions {ti}, shared wavefronts {tw}, global L1 tag requests {tg}")
for ln, src, s, i, w, g in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{100 * s / max(ts, 1):5.1f}% smp {100 * i / max(ti, 1):5.1f}% ins {100 * w / max(tw, 1):5.1f}% shw {100 * g / max(tg, 1):5.1f}% gl  L{ln:<5d} {src[:90]}")
