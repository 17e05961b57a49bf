#!/usr/bin/env python
"""bench.py -- taxa/sec on the initial NJ + TopHits build (BASELINE.json `metric`).

A "step" is one pass of the hot path over one synthetic alignment: NeighbourJoining ctor tail
(out-profile + initial out-distances) + fastNJ() with top-hits, i.e. what the reference reports
as "Initial topology".  Workload at N=1: BASELINE.json configs[1], 16 000 taxa x 200 nt columns,
fp32 (nt fp32 in the reference is its SSE3 path).

  value    taxa/s with the leaf codes already resident in HBM (device stopwatch, CUDA events on the
           context's stream, around ctor tail + fastNJ)
  e2e      taxa/s through the public C-ABI call vft_nj_build with HOST buffers: context creation,
           pinned staging + H2D of the alignment, every per-batch H2D/D2H and the tree read-back inside
  roofline the distance sweep (k_dist_pairs / k_one_vs_all / k_out_distance): algorithmic bytes
           (counted by the library per call) / its device time from CUDA events on the
           launching stream, measured in one extra profiled pass after the timed steps
  cpu_baseline / --impl reference: the UNMODIFIED reference binary built under oracle/_ref
           (-threads all host cores), timed on the same alignment: "Initial topology" stamp minus
           the "Identified unique sequences" stamp.

Multi-GPU (torchrun, one rank per GPU): each rank builds the tree of its own alignment (seed =
1+rank): replicas, weak scaling, no data-path collective.  NCCL is used
for the barrier and the max-over-ranks reduction only.
"""
from __future__ import annotations

import argparse
import json
import os

# One rank per GPU shares the host cores with the other ranks: the driver's host-thread regions must not
# oversubscribe them (spinning OpenMP workers of several ranks on the same cores cost 5x), and NCCL's banner must
# not land on stdout next to the JSON line.  Both are read when the libraries load, so they are set first.
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["OMP_WAIT_POLICY"] = "PASSIVE"
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from veryfasttree_b200 import api, synth  # noqa: E402

WORKLOAD = {"name": "16k-taxa x 200-col nucleotide, JC/%different distances, fp32 (BASELINE.json configs[1])",
            "n": 16000, "pos": 200, "kind": "nt", "precision": 32}
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "VeryFastTree")


def make_workload(seed: int, n: int):
    chars = synth.make_alignment(n, WORKLOAD["pos"], WORKLOAD["kind"], seed)
    chars = chars[synth.unique_rows(chars)]          # the reference's Uniquify; report nUnique
    return chars


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def best_thread_count(host_cores: int, seed: int = 99):
    """The reference's OpenMP build gets SLOWER with many threads on this path (measured: 128 threads
    are >10x slower than 1 on a 128-core host), so "all the host threads it can use" is calibrated:
    the thread count that is fastest on a 3000-taxon sample of the same workload is the one timed."""
    sample = make_workload(seed, 3000)
    best, best_t = 1, None
    for th in [1, 2, 4, 8, 16, 32, 64, 128]:
        if th > host_cores:
            break
        t, _ = run_reference(sample, th, timeout=120)
        if t is None:
            continue
        if best_t is None or t < best_t:
            best, best_t = th, t
        elif t > 2.0 * best_t:
            break
    return best


def run_reference(chars, threads: int, timeout: float = 900):
    """Times the unmodified reference on `chars`; returns (seconds of the NJ+TopHits phase, nUnique)."""
    if not os.path.exists(REF_BIN):
        return None, None
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "a.fa")
        synth.write_fasta(fa, chars)
        args = [REF_BIN, "-nt", "-threads", str(threads), "-noml", "-nni", "0", "-spr", "0", "-nosupport",
                "-log", os.path.join(td, "log"), fa]
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        try:
            p = subprocess.run(args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env, timeout=timeout)
        except subprocess.TimeoutExpired:
            return None, None
        if p.returncode != 0:
            return None, None
        log = open(os.path.join(td, "log")).read()
        m = re.search(r"Initial topology in ([0-9.]+) seconds", log)
        # stderr progress stamps: "   0.05 seconds: Identified unique sequences"
        u = re.search(r"([0-9.]+) seconds: Identified unique sequences", p.stderr.replace("\r", "\n"))
        if not m:
            return None, None
        t = float(m.group(1)) - (float(u.group(1)) if u else 0.0)
        return t, chars.shape[0]


def batch_regime(lib, codes, device, n_pairs=262144, n_joins=6000):
    """The same distance kernel in its bandwidth regime: one refresh-shaped request of n_pairs candidate pairs
    (lists of internal nodes) through vft_dist_pairs; device time from CUDA events on the context's stream."""
    n, L = codes.shape
    cfg = api.make_config(n, L, 4, WORKLOAD["precision"], device=device)
    cfg.reserved = 1
    rs = np.random.RandomState(3)
    with api.Context(lib, cfg) as ctx:
        ctx.upload_leaves(codes)
        ctx.outprofile_rebuild()
        ctx.out_distance_all(n, 0.0)
        active = list(range(n))
        n_joins = min(n_joins, n // 2 - 2)
        for k in range(n_joins):
            a = active.pop(rs.randint(len(active))); b = active.pop(rs.randint(len(active)))
            ctx.profile_average_update(n + k, a, b, n - k, -1.0, 0.001)
            active.append(n + k)
        internal = np.arange(n, n + n_joins)
        m = 128
        pi = np.repeat(internal[rs.randint(0, n_joins, size=m)], n_pairs // m)
        pj = internal[rs.randint(0, n_joins, size=n_pairs)]
        best = None
        for _ in range(4):
            c0 = ctx.counters()
            ctx.dist_pairs(pi, pj)
            c1 = ctx.counters()
            ms, by = c1.msDist - c0.msDist, c1.algoBytes - c0.algoBytes
            if best is None or ms < best[0]:
                best = (ms, by)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbps = best[1] / best[0] / 1e6
    return {"what": "one vft_dist_pairs request of %d internal-node pairs (%d lists), k_eval(batch)" % (n_pairs, 128),
            "ms": round(best[0], 3), "algorithmic_bytes": int(best[1]), "achieved": round(gbps, 1), "unit": "GB/s",
            "frac": round(gbps / peak, 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--taxa", type=int, default=WORKLOAD["n"], help="override the workload size (debug only)")
    args = ap.parse_args()

    # stdout carries exactly ONE line, the JSON: anything a library prints on fd 1 meanwhile (NCCL's version banner)
    # is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(json_fd, 1)
        print(json.dumps(line), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    host_cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        chars = make_workload(1, args.taxa)
        if not os.path.exists(REF_BIN):
            emit({"impl": "reference", "unavailable": "oracle/_ref/VeryFastTree was not built (no /root/reference at build time)"})
            return 0
        times = []
        threads = best_thread_count(host_cores)
        for it in range(args.warmup + args.steps):
            t, nu = run_reference(chars, threads)
            if t is None:
                emit({"impl": "reference", "unavailable": "reference binary failed to run"})
                return 0
            if it >= args.warmup:
                times.append(t)
        total = sum(times)
        value = nu * len(times) / total
        line = {"impl": "reference", "metric": "taxa/sec on initial NJ+TopHits build", "value": value, "unit": "taxa/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD["name"], "taxa": int(nu), "columns": WORKLOAD["pos"],
                           "reference_flags": "-nt -threads %d -noml -nni 0 -spr 0 -nosupport (AUTO ext = SSE3 for nt fp32)" % threads,
                           "host_cores": host_cores},
                "cpu_baseline": {"value": value, "unit": "taxa/s", "cores": threads, "kind": "reference",
                                 "sample": "the full workload, %d timed runs of the unmodified reference binary; threads = fastest of a 1..%d sweep on a 3000-taxon sample" % (len(times), host_cores)},
                "e2e": {"value": value, "unit": "taxa/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = api.load()
    # host threads of the driver's list-processing regions: this rank's share of the cores
    host_threads = max(1, min(16, host_cores // max(1, world)))

    chars = make_workload(1 + rank, args.taxa)
    codes = api.encode(chars, WORKLOAD["kind"])
    n_unique = codes.shape[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(profile=False):
        flush.fill_(1)                       # L2 flush between iterations
        torch.cuda.synchronize()
        return api.nj_build(codes, 4, WORKLOAD["precision"], lib=lib, device=local_rank, trace=False, profile=profile,
                            host_threads=host_threads)

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, e2e_s, launches, h2d, d2h = 0.0, 0.0, 0, 0, 0
    per_step = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        tr = one_step()
        dev_ms += tr.stats["deviceMsResident"]
        e2e_s += tr.stats["secondsEndToEnd"]
        per_step.append((round(tr.stats["deviceMsResident"], 1), round(tr.stats["secondsEndToEnd"] * 1e3, 1)))
        launches += tr.stats["counters"]["launches"]
        h2d += tr.stats["counters"]["h2dBytes"]
        d2h += tr.stats["counters"]["d2hBytes"] + 4 * 2 * n_unique * 4     # + tree read-back
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # max over ranks of the device-timed and end-to-end sums, sum of the units (veryfasttree_b200/dist.py)
    from veryfasttree_b200 import dist as vdist
    dev_ms_max, e2e_max, taxa_total, launches_total = vdist.aggregate_step_times(dev_ms, e2e_s, float(n_unique), float(launches), device="cuda")

    # one extra profiled pass: per-kernel-class device time from CUDA events on the launching stream
    ptree = one_step(profile=True)
    prof = ptree.stats["counters"]
    if rank == 0:
        st = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in tr.stats.items() if k not in ("counters",)}
        st["secondsHost"] = [round(x, 3) for x in tr.stats["secondsHost"]]
        print("[bench] last timed step:", json.dumps(st), file=sys.stderr)
        print("[bench] profiled pass counters:", json.dumps({k: v for k, v in prof.items() if k not in ("msKernel", "nKernel")}), file=sys.stderr)
        print("[bench] per step (device ms, end-to-end ms):", per_step, file=sys.stderr)
        for nm, ms, cnt in zip(api.KERNEL_NAMES, prof["msKernel"], prof["nKernel"]):
            if cnt:
                print("[bench] profiled pass  %-24s %7d launches %9.2f ms  %8.2f us/launch" % (nm, cnt, ms, 1e3 * ms / cnt), file=sys.stderr)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # per-kernel table of the profiled pass (CUDA events on the launching stream around every launch)
    kernels = {}
    for nm, ms, cnt, by in zip(api.KERNEL_NAMES, prof["msKernel"], prof["nKernel"], prof["bytesKernel"]):
        if cnt:
            kernels[nm] = {"launches": int(cnt), "ms": round(ms, 2), "us_per_launch": round(1e3 * ms / cnt, 2),
                           "algorithmic_gb": round(by / 1e9, 3), "gbps": round(by / ms / 1e6, 1) if by and ms > 0 else None}
    dist_kernels = [k for k in kernels if kernels[k]["algorithmic_gb"]]
    dom = max(dist_kernels, key=lambda k: kernels[k]["ms"])
    achieved = kernels[dom]["gbps"]
    batch = batch_regime(lib, codes, local_rank) if rank == 0 else None
    roofline = {"bound": "hbm", "kernel": dom + " -- the dominant kernel of the step: one launch per per-join candidate list "
                "(40-250 pairs of 0.2-4.4 KB), latency bound, the slab is L2 resident",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json (burst copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": int(1e9 * kernels[dom]["algorithmic_gb"] / kernels[dom]["launches"]),
                "us_per_launch": kernels[dom]["us_per_launch"],
                "traffic": 256, "traffic_note": "dram__bytes_read+write per launch from profiles/r1_k_eval_small_ncu_full.md: "
                "the 75 MB slab stays in the 126 MB L2, so DRAM traffic is ~0 and far BELOW the algorithmic bytes",
                "distance_sweep_all_kernels": {"algorithmic_bytes_per_step": prof["distBytes"], "ms_per_step": prof["msDist"],
                                               "gbps": (prof["distBytes"] / (prof["msDist"] * 1e-3)) / 1e9 if prof["msDist"] > 0 else None},
                "batch_regime": batch}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu = None
    if args.gpus == 1:
        threads = best_thread_count(host_cores)
        t, nu = run_reference(chars, threads)
        if t is not None:
            cpu = {"value": nu / t, "unit": "taxa/s", "cores": threads, "kind": "reference", "host_cores": host_cores,
                   "sample": "the full workload once (%d unique taxa x %d columns), unmodified reference binary, -threads %d "
                             "(fastest of a 1..%d sweep on a 3000-taxon sample), %.2f s" % (nu, WORKLOAD["pos"], threads, host_cores, t)}
        else:
            cpu = {"value": None, "unit": "taxa/s", "cores": host_cores, "kind": "reference", "sample": "oracle/_ref/VeryFastTree not available"}

    value = taxa_total * args.steps / (dev_ms_max * 1e-3)
    line = {"metric": "taxa/sec on initial NJ+TopHits build", "value": value, "unit": "taxa/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "taxa_per_gpu": int(n_unique), "columns": WORKLOAD["pos"],
                       "parallelism": "replicas x%d" % args.gpus, "host_threads_per_rank": host_threads,
                       "l2": "flushed between steps (256 MiB write)",
                       "arithmetic": "f32 storage, f64 accumulation of top/denom and criteria -- the reference's own mix (SURVEY 9.1)",
                       "parity": "join order, top-hit lists and branch lengths identical to the reference at -threads 1"},
            "e2e": {"value": taxa_total * args.steps / e2e_max, "unit": "taxa/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps},
            "gpu_launches": int(launches_total), "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "wall_s": t_wall}
    if world > 1:
        dist.destroy_process_group()
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
