"""File -> file throughput of the rdp_classifier executable on the bench workload (BASELINE configs[2] reads).
usage: cli_bench.py [reads] [gpus]   -> prints one JSON object (used by bench.py's secondary.cli leg and by hand)"""
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "pangea-plus_b200"))
from pangea_b200 import synth  # noqa: E402

BIN = os.path.join(HERE, "..", "pangea-plus_b200", "bin", "rdp_classifier")


def write_query(path, data, L, n, repeat):
    """n reads of L bases, written `repeat` times with running ids: '>r%08d:AB\\n' + bases + '\\n' per record"""
    arr = data.reshape(n, L)
    hdr = np.frombuffer(b">r00000000:AB\n", np.uint8)
    rec2 = np.empty((n, len(hdr) + L + 1), np.uint8)
    rec2[:, :len(hdr)] = hdr
    rec2[:, len(hdr):len(hdr) + L] = arr
    rec2[:, -1] = ord("\n")
    with open(path, "wb") as f:
        for r in range(repeat):
            ids = np.arange(r * n, (r + 1) * n)
            for d in range(8):
                rec2[:, 2 + 7 - d] = ord("0") + (ids // 10 ** d) % 10
            rec2.tofile(f)
    return repeat * n


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
    gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    tmp = os.environ.get("PG_CLI_TMP", "/dev/shm/pg_cli" if os.path.isdir("/dev/shm") else "/tmp/pg_cli")
    os.makedirs(tmp, exist_ok=True)
    tr = synth.synth16s(0x9178, 9178, 1219)
    names, anc = tr["node_names"], tr["anc"]
    with open(f"{tmp}/train.fa", "w") as f:
        for i, g in enumerate(tr["genus"]):
            s = tr["data"][tr["off"][i]:tr["off"][i + 1]].tobytes().decode()
            f.write(f">T{i:06d}\t" + ";".join(names[n] for n in anc[g]) + "\n" + s + "\n")
    n = 1 << 18
    data, off, src = synth.synth_reads(0x250, tr, n, paired=True)
    total = write_query(f"{tmp}/q.fa", data, 689, n, max(1, reads // n))
    subprocess.run([BIN, "--train", f"{tmp}/train.fa", "-t", f"{tmp}/m.pgm"], check=True, capture_output=True)
    out = {"reads": total, "gpus": gpus, "query_bytes": os.path.getsize(f"{tmp}/q.fa")}
    for fmt in ("allrank", "pangea"):
        t = time.time()
        r = subprocess.run([BIN, "-q", f"{tmp}/q.fa", "-o", f"{tmp}/o_{fmt}.txt", "-t", f"{tmp}/m.pgm", "-f", fmt, "--gpus", str(gpus)],
                           capture_output=True, text=True, env=dict(os.environ, PG_TIMING="1"))
        wall = time.time() - t
        if r.returncode != 0:
            out[fmt] = {"error": r.stderr[-400:]}
            continue
        m = re.search(r"pipeline ([0-9.]+) s for (\d+) reads = ([0-9.]+) M reads/s.*busy: read ([0-9.]+), gpu ([0-9.]+) \(ingest ([0-9.]+)\), format ([0-9.]+), write ([0-9.]+)", r.stderr)
        out[fmt] = {"timing": [l for l in r.stderr.splitlines() if l.startswith("[timing]")], "wall_s": wall, "wall_reads_per_s": total / wall, "output_bytes": os.path.getsize(f"{tmp}/o_{fmt}.txt")}
        if m:
            out[fmt].update(pipeline_s=float(m.group(1)), pipeline_reads_per_s=1e6 * float(m.group(3)),
                            busy_s={"read": float(m.group(4)), "gpu": float(m.group(5)), "ingest": float(m.group(6)), "format": float(m.group(7)), "write": float(m.group(8))})
        # sanity: as many lines as reads, the first one assigned down to a genus
        with open(f"{tmp}/o_{fmt}.txt", "rb") as f:
            first = f.readline()
        out[fmt]["first_line"] = first.decode()[:120]
    for fn in ("q.fa", "o_allrank.txt", "o_pangea.txt", "train.fa") if not os.environ.get("PG_CLI_KEEP") else ():
        try:
            os.remove(f"{tmp}/{fn}")
        except OSError:
            pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
