"""Top SASS lines by warp-stall samples for one kernel of an .ncu-rep (needs -lineinfo / --import-source on).
usage: ncu_hotspots.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print(f"kernel {kern}: {len(body)} SASS lines, {tot} samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stalls}
print("stall mix:", ", ".join(f"{h[6:]} {100 * v / max(tot, 1):.0f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
body.sort(key=lambda r: -int(r[col["# Samples"]] or 0))
for idx, r in enumerate(body[:top]):
    s = int(r[col["# Samples"]] or 0)
    main = max(stalls, key=lambda h: int(r[col[h]] or 0))
    print(f"{100 * s / max(tot, 1):5.1f}%  {r[col['Source']][:70]:70s} exec={r[col['Instructions Executed']]:>9s} {main[6:]}")
