#!/usr/bin/env python
"""bench.py -- classified reads/s (8-mer naive Bayes + 100 bootstraps) on N B200s.

Workload (BASELINE.json configs[2], the one the metric is quoted on): synthetic
250 bp paired Illumina 16S reads joined the PANGEA way (mateA + N x 189 + mateB,
486 good words) against a 9178-sequence / 1219-genus model.  The reference's
rdp_download_9178seqs.fa is absent from its tree, so the training set is the
seeded substitute synth16s(0x9178, 9178, 1219).

A "step" = one pass of the hot path (word extraction + orientation, gather-sum,
100 bootstraps, argmax, vote) over one batch of reads per GPU.
  value : reads/s with the 2-bit packed reads already resident in HBM
  e2e   : reads/s through pg_classify() with HOST buffers (ASCII reads in pinned
          memory -> H2D -> pack -> classify -> vote -> 64-byte records D2H)
  --impl reference : the CPU restatement of RDP 2.5 (oracle/rdp_ref.c; the jar is
          not vendored and there is no JVM) on the host cores, same reads.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))

METRIC = "classified reads/sec (8-mer NB, 100 boots)"
TRAIN_SEED, TRAIN_SEQS, TRAIN_GENERA = 0x9178, 9178, 1219
READ_SEED = 0x250


# BASELINE configs[3] in its genus dimension: 10 000 genera (the ~3M-sequence training set is cut to 6 per
# genus -- the table, and so the classification cost, depends on G only)
RDP_SCALE = (0x3000000, 60000, 10000)
WORKLOAD = "illumina"


def make_workload(paired: bool, nreads: int, nbatches: int):
    from pangea_b200 import synth

    seed, seqs, genera = RDP_SCALE if WORKLOAD == "rdp_scale" else (TRAIN_SEED, TRAIN_SEQS, TRAIN_GENERA)
    tr = synth.synth16s(seed, seqs, genera)
    batches = [synth.synth_reads(READ_SEED + b, tr, nreads, paired=paired) for b in range(nbatches)]
    return tr, batches


def algorithmic_bytes_per_read(n_words: int, G: int, L: int) -> int:
    # every (word, genus) fp32 cell of the dense table once per read,
    # the 2-bit packed bases + validity mask, and the 64-byte result record.
    return n_words * G * 4 + math.ceil(L / 4) + 8 * math.ceil(L / 64) + 64


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread every 2 ms (the timed
    region of a default run is a fraction of a second -- `nvidia-smi -lms` would not deliver a single sample)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.h, self.nv = index, None, None
        self.sm, self.mask, self.stop_flag, self.t = [], 0, False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                import torch

                pr = torch.cuda.get_device_properties(index)
                bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None or self.h is None:
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.t.join(timeout=2)
        try:
            mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.sm)}


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(reads_per_launch: float):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed
    `ncu --set full` capture (profiles/classify_traffic.json), per launch: the capture was one
    launch of 65536 reads, so it is carried to this run's launch size per read."""
    p = REPO / "profiles" / "classify_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())["dram_bytes_per_read"] * reads_per_launch
        except Exception:
            return None
    return None


def cpu_oracle_rate(tr, data, off, seconds: float, threads: int):
    """reads/s of oracle/rdp_ref.c on a bounded sample of the same reads."""
    sys.path.insert(0, str(REPO / "tests"))
    import oracle_rdp as ora

    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    n = len(off) - 1
    probe = min(n, 4 * threads)
    t0 = time.perf_counter()
    om.classify(data[: off[probe]], off[: probe + 1], 0, threads)
    dt = time.perf_counter() - t0
    sample = int(max(probe, min(n, probe * seconds / max(dt, 1e-6))))
    t0 = time.perf_counter()
    om.classify(data[: off[sample]], off[: sample + 1], 0, threads)
    dt = time.perf_counter() - t0
    om.free()
    return sample / dt, sample, dt


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    paired = not args.single
    tr, batches = make_workload(paired, max(args.ref_reads, 64), 1)
    data, off, _ = batches[0]
    sys.path.insert(0, str(REPO / "tests"))
    import oracle_rdp as ora

    threads = os.cpu_count() or 1
    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    n = args.ref_reads
    for _ in range(args.warmup):
        om.classify(data[: off[min(n, 4 * threads)]], off[: min(n, 4 * threads) + 1], 0, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        om.classify(data[: off[n]], off[: n + 1], 0, threads)
    dt = time.perf_counter() - t0
    om.free()
    rate = n * args.steps / dt
    L = int(off[1] - off[0])
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(paired, n, L, tr["G"], "host cores only"),
        "cpu_baseline": {"value": rate, "unit": "reads/s", "cores": threads, "kind": "port",
                         "sample": f"{n} reads per step x {args.steps} steps of the same synthetic workload; "
                                   "oracle/rdp_ref.c (C restatement of RDP 2.5; the jar is not vendored and no JVM exists)"},
        "e2e": {"value": rate, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(paired, reads_per_step, L, G, l2_note):
    return {
        "workload": ("1M-class synthetic 250 bp paired Illumina 16S reads (mateA+N*189+mateB, 486 words)" if paired
                     else "synthetic 250 bp single-end 16S reads (243 words)")
                    + (f" vs synth16s 9178-seq/{G}-genus model [BASELINE configs[2]; 9178-seq file absent from the reference]"
                       if WORKLOAD == "illumina" else
                       f" vs synth16s {RDP_SCALE[1]}-seq/{G}-genus model [BASELINE configs[3] in its genus dimension]"),
        "reads_per_step_per_gpu": reads_per_step, "record_len": L, "genera": G, "bootstraps": 100,
        "l2": l2_note,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1 << 18, help="reads per step per GPU")
    ap.add_argument("--ref-reads", type=int, default=2048, help="reads per step of the CPU reference arm")
    ap.add_argument("--workload", default="illumina", choices=["illumina", "rdp_scale"],
                    help="illumina = BASELINE configs[2] (default, the metric's configuration); rdp_scale = 10 000 genera")
    ap.add_argument("--single", action="store_true", help="single-end 250 bp (243 words) instead of the joined pair")
    ap.add_argument("--mode", type=int, default=int(os.environ.get("PG_BENCH_MODE", "1")), help="1 certified (default: same results, half the shared-memory traffic), 0 strict")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    global WORKLOAD
    WORKLOAD = args.workload
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import pangea_b200 as pg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    paired = not args.single
    nb = 2
    # every rank draws its own reads (different seeds): read shards are independent
    global READ_SEED
    READ_SEED += 1000 * rank
    tr, batches = make_workload(paired, args.reads, nb)
    L = int(batches[0][1][1])
    G = tr["G"]

    ctx = pg.Context(local)
    # a dedicated (non-default) stream: the library, the NCCL ops and the timing events all
    # run on it, so the CUDA events bracket exactly the work that was launched
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    assert stream.cuda_stream != 0

    # ---- model: rank 0 trains on its GPU, the integer counts are NCCL-broadcast once,
    # every rank derives the fp32 table locally (identical bits, no table broadcast needed)
    if rank == 0:
        model = ctx.train(tr["data"], tr["off"], tr["genus"], G)
    else:
        model = ctx.model_create(G)
    if world > 1:
        from pangea_b200 import dist as pgdist

        ctx.sync()
        pgdist.broadcast_buffers([torch.as_tensor(pgdist.DeviceBuffer(ptr, nbytes), device=dev)
                                  for ptr, nbytes in model.buffers()], src=0)
        torch.cuda.synchronize()
        if rank != 0:
            model.commit()
    model.set_lineage(tr["anc"])

    # ---- device-resident packed reads for the `value` leg; pinned host copies for e2e
    packed, pinned = [], []
    for data, off, _ in batches:
        packed.append(ctx.pack(data, off))
        hb = torch.empty(data.size, dtype=torch.uint8, pin_memory=True)
        hb.numpy()[:] = data
        ho = torch.empty(off.size, dtype=torch.int64, pin_memory=True)
        ho.numpy()[:] = off
        pinned.append((hb, ho))
    res_dev = torch.empty(args.reads * 64, dtype=torch.uint8, device=dev)
    gather_list = [torch.empty_like(res_dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    res_host = torch.empty(args.reads * 64, dtype=torch.uint8, pin_memory=True)
    res_np = res_host.numpy().view(pg.RESULT_DTYPE)

    def step_resident(i):
        ctx.classify_packed(model, packed[i % nb], res_dev, None, mode=args.mode)
        if world > 1:
            dist.gather(res_dev, gather_list, dst=0)          # results back to rank 0 in rank order

    def step_e2e(i):
        hb, ho = pinned[i % nb]
        ctx.classify(model, hb.numpy(), ho.numpy(), mode=args.mode, out=res_np)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(steps):
            fn(i)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.kernel_time_reset()
    l0 = ctx.launch_count()
    ms = timed(step_resident, args.steps)
    launches = ctx.launch_count() - l0
    kms, klaunch = ctx.kernel_time()
    clocks = sampler.stop() if rank == 0 else None
    route = ctx.classify_stats()                 # routing of the last step's reads inside the library

    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

    total_reads = args.reads * args.steps * world
    value = total_reads / (ms / 1e3)
    e2e_value = total_reads / (ms_e2e / 1e3)

    # correctness guard: the timed path produced real assignments
    chk = np.frombuffer(res_dev.cpu().numpy().tobytes(), dtype=pg.RESULT_DTYPE)
    assert (chk["status"] == 0).all() and (chk["votes"][:, 0] == 100).all(), "bench produced invalid records"

    if rank == 0:
        peak, peak_src = peaks()
        n_words = int(np.median(chk["n_words"]))
        bpr = algorithmic_bytes_per_read(n_words, G, L)
        reads_per_launch = args.reads * args.steps / max(klaunch, 1)
        avg_launch_s = (kms / max(klaunch, 1)) / 1e3
        achieved = bpr * reads_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(paired, args.reads, L, G,
                                           f"inputs per step ({args.reads * L / 1e6:.0f} MB ASCII, "
                                           f"{args.reads * n_words * 2 / 1e6:.0f} MB word ids) exceed the 126 MB L2; batches alternate"),
                           mode="strict" if args.mode == 0 else "certified", parallelism=f"reads sharded x{world}",
                           median_words=n_words, routing_last_step=route),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(reads_per_launch), "peak_source": peak_src,
                         "kernel": ("certified phase 1 = k_guess_bm + k_classify_h (best part of the best block) + k_bound + k_light, timed as one group"
                                    if args.mode == 1 else "k_classify_strict"),
                         "kernel_ms_per_launch": kms / max(klaunch, 1),
                         "kernel_share_of_step": kms / ms if ms > 0 else None,
                         "algorithmic_bytes_per_read": bpr, "reads_per_launch": reads_per_launch},
            "e2e": {"value": e2e_value, "unit": "reads/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(args.reads * L + (args.reads + 1) * 8 + args.reads * 4),
                    "d2h_bytes_per_step": int(args.reads * 64 + args.reads * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            data, off, _ = batches[0]
            threads = os.cpu_count() or 1
            rate, sample, dt = cpu_oracle_rate(tr, data, off, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "reads/s", "cores": threads, "kind": "port",
                                    "sample": f"first {sample} reads of step 0 ({dt:.1f} s), oracle/rdp_ref.c with OpenMP over reads; "
                                              "stands in for the RDP 2.5 jar, which is not vendored and cannot run (no JVM)"}
        print(json.dumps(line), flush=True)

    for p in packed:
        p.free()
    model.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
