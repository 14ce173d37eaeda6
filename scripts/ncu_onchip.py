"""profiles/r2_onchip.json from an `ncu --set full` capture of bench.py (one launch of every hot kernel on the same
chunk of reads): the on-chip counters bench.py quotes in roofline.onchip and the measured DRAM bytes per read of the
certified phase-1 group (k_guess_bm + k_mma_meta + k_mma_bound + k_light; with cert_plan 3: k_classify_h + k_bound).
usage: ncu_onchip.py capture.ncu-rep out.json [metrics.csv]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = r[col[name]].replace(",", "")
    u = units[col[name]]
    x = float(v)
    return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0) if "byte" in u else x


def ms(r):
    v, u = float(r[col["gpu__time_duration.sum"]].replace(",", "")), units[col["gpu__time_duration.sum"]]
    return v / {"ns": 1e6, "us": 1e3, "ms": 1.0, "s": 1e-3}.get(u, 1e6)


def reads_of(r):
    name = r[col["Kernel Name"]]
    grid = int(r[col["launch__grid_size"]].replace(",", ""))
    if "k_classify_h" in name or "k_resolve" in name:
        return grid * 4
    if "k_guess" in name:
        return grid * 8
    if "k_bound" in name and "k_mma" not in name:
        return grid
    return None


tensor_cols = [h for h in hdr if "tensor" in h and "pct" in h and (".avg." in h) and "ops_path" not in h]


launch = []
for r in data:
    name = r[col["Kernel Name"]].split("(")[0]
    launch.append(dict(kernel=name, reads=reads_of(r), ms=ms(r), grid=int(r[col["launch__grid_size"]].replace(",", "")),
                       l1_data_pipe_pct=val(r, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
                       shared_wavefront_pct=val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                       issue_pct=val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active"),
                       warps_active_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                       l2_hit_pct=val(r, "lts__t_sector_hit_rate.pct"),
                       dram_bytes=val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                       tensor={h: val(r, h) for h in tensor_cols if r[col[h]].replace(",", "").replace(".", "").isdigit()}))
# the chunk: the k_guess_bm launch with the most reads, and the group's other kernels launched on the same reads after it
ib = max((i for i, l in enumerate(launch) if "k_guess_bm" in l["kernel"]), key=lambda i: launch[i]["reads"])
nreads = launch[ib]["reads"]
group = {}
for i in range(ib, min(len(launch), ib + 6)):
    l = launch[i]
    for key in ("k_guess_bm", "k_mma_meta", "k_mma_bound", "k_classify_h", "k_bound", "k_light", "k_resolve"):
        if l["kernel"].startswith(key) and key not in group:
            group[key] = l
phase1 = [group[k] for k in ("k_guess_bm", "k_mma_meta", "k_mma_bound", "k_classify_h", "k_bound", "k_light") if k in group]
res = {
    "capture": rep.split("/")[-1], "reads_in_chunk": nreads,
    "dram_bytes_per_read": sum(l["dram_bytes"] for l in phase1) / nreads,
    "phase1_ns_per_read_under_ncu": 1e6 * sum(l["ms"] for l in phase1) / nreads,
    "onchip": {k: {"counter": "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "value": round(l["l1_data_pipe_pct"], 2),
                   "shared_wavefronts_pct_of_peak": round(l["shared_wavefront_pct"], 2), "issue_slots_pct": round(l["issue_pct"], 2),
                   "warps_active_pct": round(l["warps_active_pct"], 2), "l2_hit_pct": round(l["l2_hit_pct"], 2), "ms": round(l["ms"], 4),
                   "ns_per_read": round(1e6 * l["ms"] / nreads, 2), "dram_bytes_per_read": round(l["dram_bytes"] / nreads, 1),
                   **({"tensor_pipe": {h: round(v, 2) for h, v in l["tensor"].items() if v > 0}} if k == "k_mma_bound" and l["tensor"] else {})}
               for k, l in group.items()},
    "group": [k for k in ("k_guess_bm", "k_mma_meta", "k_mma_bound", "k_classify_h", "k_bound", "k_light") if k in group],
    "note": "ncu --set full --clock-control none, one launch per kernel on the same chunk of reads of `bench.py --steps 2 --warmup 3 --legs none`; "
            "cold-cache, serialised timings (shares agree with the live CUDA-event numbers, absolutes do not)",
}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        w = csv.writer(f)
        keys = [k for k in launch[0].keys() if k != "tensor"]
        w.writerow(keys)
        for l in launch:
            w.writerow([l[k] for k in keys])
