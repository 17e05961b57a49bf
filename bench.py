#!/usr/bin/env python
"""bench.py -- taxa/sec on the initial NJ + TopHits build (BASELINE.json `metric`).

A "step" is one pass of the hot path over one synthetic alignment: NeighbourJoining ctor tail (out-profile + initial
out-distances) + fastNJ() with top-hits, i.e. what the reference reports as "Initial topology".

Workload at N=1 (`--workload c3`, the default): C3-SHAPED -- amino acid, 1287 columns, BLOSUM45 distances, fp32, the
regime BASELINE.json's >=50x target is quoted on -- at 20 000 taxa, the largest size for which the 25 reference runs of the
driver's reference arm fit its time limit (the reference needs ~21 s per run at 16 threads; 100 000 taxa would need hours).
`--workload c2` is BASELINE.json configs[1] (16 000 x 200 nt), the workload of round 1.

  value    taxa/s with the leaf codes already resident in HBM (device stopwatch, CUDA events on the context's stream,
           around ctor tail + fastNJ)
  e2e      taxa/s through the public C-ABI call vft_nj_build with HOST buffers: context creation, pinned staging + H2D of
           the alignment, every list copy and the tree read-back inside the timed region
  roofline the kernel with the largest share of the step's device time that moves profile data: algorithmic bytes (SURVEY
           8d, counted by the library per launch) / its device time from CUDA events on the launching stream, measured in one
           extra profiled pass after the timed steps.  `traffic` is null here: DRAM bytes need ncu, see profiles/
  parity_checked  the product's Newick against the `NJ` line of the reference's -log.  The reference's tree depends on its
           thread count (psort tie placement, SURVEY 9.4: measured, -threads 4 != -threads 1 on this workload), so the check
           runs the reference at `-threads 1 -ext AVX2` on a bounded prefix of the SAME alignment (first 3 000 taxa); the
           full sizes are checked by tests/test_gpu_at_size.py
  cpu_baseline / --impl reference: the UNMODIFIED reference binary (oracle/_ref, -mavx2 build) with -ext AVX2 -fastexp 3 on the
           same alignment: "Initial topology" stamp minus the "Identified unique sequences" stamp; its thread count is the
           fastest of {cores/2, cores} measured ON THE FULL WORKLOAD (the reference gets slower with too many threads).

Multi-GPU (torchrun, one rank per GPU): `--gpus N` builds ONE tree sharded over the N GPUs (`--multi sharded`, the default;
SURVEY 8e): every rank runs the same host loop on a replicated slab, the candidate axis of the all-candidate sweeps
(setBestHit, the all-node out-distances, the list merges of a refresh) is split over the ranks and their results are
all-gathered over NVLink peer memory (csrc/vft_dist.cuh; VFT_EXCHANGE=nccl for ncclAllGather).  STRONG scaling: the same
20 000-taxon tree at every N, value = its taxa / the max over ranks of the device time; every rank's tree is checked to be the
same (`trees_identical_across_ranks`).  The reference arm times the same one tree at every N.  `--multi replicas` keeps the
earlier mode (N independent trees, no data-path collective, weak scaling).
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


def _bind_openmp():
    """The step is bound by ONE host thread (the join loop) plus short OpenMP bursts (the refreshes): threads that stay on their
    cores are worth 3 % of the step on one GPU (measured: 2.23 s -> 2.16 s with OMP_PLACES=cores).  With one rank per GPU a per-rank
    slice of the CPUs measured slightly WORSE than no binding (2.26 s vs 2.23 s at N=2), so the ranks stay unbound.  Not applied
    when the caller has set OMP_PROC_BIND / OMP_PLACES, and never passed on to the reference's processes."""
    if "OMP_PROC_BIND" in os.environ or "OMP_PLACES" in os.environ or os.environ.get("VFT_BENCH_BIND", "1") == "0":
        return
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        return
    os.environ["OMP_PLACES"] = "cores"
    os.environ["OMP_PROC_BIND"] = "true"
    os.environ["VFT_BENCH_BOUND"] = "1"


_bind_openmp()
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

WORKLOADS = {
    "c3": {"name": "C3-shaped: 20k-taxa x 1287-col amino acid, BLOSUM45 distances, fp32 (BASELINE.json configs[2] at the largest N "
                   "whose 25 reference runs fit the driver's reference arm)",
           "n": 20000, "pos": 1287, "kind": "aa", "precision": 32, "ref_flags": ["-ext", "AVX2", "-fastexp", "3"]},
    "c2": {"name": "16k-taxa x 200-col nucleotide, JC/%different distances, fp32 (BASELINE.json configs[1])",
           "n": 16000, "pos": 200, "kind": "nt", "precision": 32, "ref_flags": ["-nt"]},
    # BASELINE.json configs[2] at its full size: by hand only (the reference needs ~4 min per run at this size)
    "c3full": {"name": "C3: 100k-taxa x 1287-col amino acid, BLOSUM45 distances, fp32 (BASELINE.json configs[2])",
               "n": 100000, "pos": 1287, "kind": "aa", "precision": 32, "ref_flags": ["-ext", "AVX2", "-fastexp", "3"]},
}
# DRAM bytes per launch of the kernels, from committed `ncu --set full` captures of this workload (profiles/r2/): the
# "traffic" side of the roofline, which cannot be measured inside this script
NCU_TRAFFIC = os.path.join(ROOT, "profiles", "r2", "ncu_traffic.json")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "VeryFastTree")
REF_BIN512 = os.path.join(ROOT, "oracle", "_ref", "VeryFastTree_avx512")      # the same sources with the reference's -ext AVX512 path compiled in
PARITY_PREFIX = 3000


def make_workload(wl, seed: int, n: int):
    chars = synth.make_alignment(n, wl["pos"], wl["kind"], seed)
    chars = chars[synth.unique_rows(chars)]          # the reference's Uniquify; report nUnique
    return chars


def tables_for(wl):
    if wl["kind"] != "aa":
        return None
    z = np.load(os.path.join(ROOT, "tests", "golden", "blosum45_f%d.npz" % wl["precision"]))
    return [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md), 5 samples per second.  Read in-process through
    NVML (pynvml): the step is bound by ONE host thread's list logic, and spawning nvidia-smi five times a second next to it
    costs it measurable time; nvidia-smi is the fallback when NVML cannot be loaded."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates every GPU of the box; honour CUDA_VISIBLE_DEVICES when it lists ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        return [str(mhz), str(self.max_mhz)] + ["Active" if mask & b else "Not Active" for b in bits]

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return [x.strip() for x in out.split(",")] if out else None

    def run(self):
        while not self.stop_flag:
            try:
                row = self._sample_nvml() if self.nvml is not None else self._sample_smi()
                if row:
                    self.rows.append(row)
            except Exception:
                self.nvml = None             # NVML misbehaving: nvidia-smi from now on
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = [nm for k, nm in enumerate(self.NAMES) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def run_reference(wl, chars, threads: int, timeout: float = 1500, want_tree: bool = False, extra=None, binary=None):
    """Times the unmodified reference on `chars`; returns (seconds of the NJ+TopHits phase, nUnique[, NJ tree])."""
    none = (None, None, None) if want_tree else (None, None)
    binary = binary or REF_BIN
    if not os.path.exists(binary):
        return none
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "a.fa")
        synth.write_fasta(fa, chars)
        args = [binary] + list(extra if extra is not None else wl["ref_flags"]) + ["-threads", str(threads), "-noml", "-nni", "0", "-spr", "0",
                                                                                   "-nosupport", "-log", os.path.join(td, "log"), fa]
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        env.pop("OMP_WAIT_POLICY", None)      # (set above for the repo arm's ranks only: the reference runs with its own defaults at every N)
        if env.pop("VFT_BENCH_BOUND", None):
            env.pop("OMP_PLACES", None); env.pop("OMP_PROC_BIND", None)
        try:
            p = subprocess.run(args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env, timeout=timeout)
        except subprocess.TimeoutExpired:
            return none
        if p.returncode != 0:
            return none
        log = open(os.path.join(td, "log")).read()
        m = re.search(r"Initial topology in ([0-9.]+) seconds", log)
        # stderr progress stamps: "   0.05 seconds: Identified unique sequences"
        u = re.search(r"([0-9.]+) seconds: Identified unique sequences", p.stderr.replace("\r", "\n"))
        if not m:
            return none
        t = float(m.group(1)) - (float(u.group(1)) if u else 0.0)
        if want_tree:
            nj = [line.split("\t", 1)[1].strip() for line in log.splitlines() if line.startswith("NJ\t")]
            return t, chars.shape[0], (nj[0] if nj else None)
        return t, chars.shape[0]


def calibrate_threads(wl, chars, host_cores: int):
    """The reference's OpenMP build gets SLOWER with too many threads on this path, so "all the host threads it can use" is
    calibrated ON THE FULL WORKLOAD: the faster of cores/2 and cores.  Returns (threads, {threads: seconds})."""
    cands = sorted({max(1, host_cores // 2), host_cores})
    seen = {}
    for th in cands:
        t, _ = run_reference(wl, chars, th)
        if t is not None:
            seen[th] = t
    if not seen:
        return host_cores, {}
    return min(seen, key=seen.get), seen


def avx512_variant(wl, chars, threads):
    """The reference's AVX-512 path ("AVX2/AVX512 OpenMP build"), same sources built with -mavx512*, one run at
    the calibrated thread count.  Returns (seconds or None, flags)."""
    if wl["kind"] != "aa" or not os.path.exists(REF_BIN512):
        return None, None
    flags = [f if f != "AVX2" else "AVX512" for f in wl["ref_flags"]]
    t, _ = run_reference(wl, chars, threads, extra=flags, binary=REF_BIN512)
    return t, flags


def parity_check(wl, chars, lib, device):
    """Product tree == reference tree (-threads 1 -ext AVX2) on the first PARITY_PREFIX taxa of the bench alignment."""
    sub = chars[:PARITY_PREFIX]
    extra = (["-nt"] if wl["kind"] == "nt" else []) + ["-ext", "AVX2"]
    t, nu, want = run_reference(wl, sub, 1, timeout=600, want_tree=True, extra=extra)
    if want is None:
        return {"checked": False, "why": "oracle/_ref/VeryFastTree not available"}
    A = 4 if wl["kind"] == "nt" else 20
    tree = api.nj_build(api.encode(sub, wl["kind"]), A, wl["precision"], lib=lib, device=device, tables=tables_for(wl), trace=False)
    got = tree.newick(["t%d" % i for i in range(sub.shape[0])])
    return {"checked": True, "identical": got == want,
            "sample": "first %d taxa of the bench alignment, reference %s -threads 1 (%.1f s); whole Newick string of the NJ phase compared "
                      "(topology, join order and branch lengths)" % (sub.shape[0], " ".join(extra), t),
            "full_size": "tests/test_gpu_at_size.py (16 000 x 200 nt and 4 000 x 1287 aa trees vs the reference at -threads 1)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--taxa", type=int, default=0, help="override the workload size (debug only)")
    ap.add_argument("--device-loop", type=int, default=-1, help="-1: library default; 0/1: host-driven / device-resident join loop")
    ap.add_argument("--skip-cpu", action="store_true", help="no cpu_baseline / parity legs (hand runs of the big workloads)")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "replicas"], help="--gpus N > 1: one tree sharded over the GPUs, or N independent trees")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    n_taxa = args.taxa or wl["n"]

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
    metric = "taxa/sec on initial NJ+TopHits build"

    if args.impl == "reference":
        if rank != 0:
            return 0
        if not os.path.exists(REF_BIN):
            emit({"impl": "reference", "unavailable": "oracle/_ref/VeryFastTree was not built (no /root/reference at build time)"})
            return 0
        # The host has the same cores whatever --gpus is, and N trees built one after the other at all cores give the same
        # aggregate taxa/s as one: the reference arm therefore times ONE tree per step at every N (N concurrent reference
        # processes would only split the same cores, and 25 rounds of them would not fit the driver's time limit).
        n_rep = 1
        workloads = [make_workload(wl, 1, n_taxa)]
        threads, seen = calibrate_threads(wl, workloads[0], host_cores)
        per_proc = threads
        t512, flags512 = avx512_variant(wl, workloads[0], per_proc)
        use512 = t512 is not None and seen and t512 < seen[threads]          # the faster of the two builds is the reference arm
        ref_binary, ref_flags = (REF_BIN512, flags512) if use512 else (REF_BIN, wl["ref_flags"])
        times, taxa = [], 0
        for it in range(args.warmup + args.steps):
            t, nu = run_reference(wl, workloads[0], per_proc, extra=ref_flags, binary=ref_binary)
            if t is None:
                emit({"impl": "reference", "unavailable": "reference binary failed to run"})
                return 0
            if it >= args.warmup:
                times.append(t)
                taxa = nu
        total = sum(times)
        value = taxa * len(times) / total
        flags = " ".join(ref_flags)
        line = {"impl": "reference", "metric": metric, "value": value, "unit": "taxa/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
                "higher_is_better": True, "scaling": "strong" if (args.multi == "sharded" and args.gpus > 1) else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "taxa": int(taxa // n_rep), "columns": wl["pos"],
                           "parallelism": "one tree per step at all host cores (the repo arm at --gpus N shards this same tree over N GPUs; with --multi replicas it builds N trees, "
                                          "and the CPU's aggregate taxa/s does not depend on how many trees are queued)",
                           "reference_flags": "%s -threads %d -noml -nni 0 -spr 0 -nosupport%s" % (
                               flags, per_proc, " (nt fp32: the reference silently runs its SSE3 path)" if wl["kind"] == "nt" else ""),
                           "reference_build": ("oracle/_ref/VeryFastTree_avx512: the unmodified sources, g++ -O3 -mavx2 -mavx512{f,bw,dq,vl}" if use512
                                               else "oracle/_ref/VeryFastTree: the unmodified sources, g++ -O3 -mavx2 (no FMA; the parity oracle)"),
                           "avx512_build_s": None if t512 is None else round(t512, 2),
                           "thread_calibration_s": {str(k): round(v, 2) for k, v in seen.items()}, "host_cores": host_cores},
                "cpu_baseline": {"value": value, "unit": "taxa/s", "cores": per_proc * n_rep, "kind": "reference",
                                 "sample": "the full workload, %d timed runs of the unmodified reference binary; threads = the faster of cores/2 and cores, "
                                           "measured on the full workload" % len(times)},
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
    from veryfasttree_b200 import dist as vdist
    sharded = world > 1 and args.multi == "sharded"
    dist_info = vdist.init_sharded(lib, local_rank) if sharded else {"mode": "none"}
    # host threads of the driver's list-processing regions: this rank's share of the cores
    host_threads = max(1, min(16, host_cores // max(1, world)))
    A = 4 if wl["kind"] == "nt" else 20
    tables = tables_for(wl)
    dev_loop = None if args.device_loop < 0 else args.device_loop

    chars = make_workload(wl, 1 if sharded else 1 + rank, n_taxa)      # sharded: every rank holds the same alignment
    codes = api.encode(chars, wl["kind"])
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
        return api.nj_build(codes, A, wl["precision"], lib=lib, device=local_rank, trace=False, profile=profile,
                            host_threads=host_threads, tables=tables, device_loop=dev_loop)

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
    dev_ms_max, e2e_max, taxa_total, launches_total = vdist.aggregate_step_times(dev_ms, e2e_s, float(n_unique), float(launches), device="cuda")
    trees_identical = None
    if sharded:
        taxa_total = float(n_unique)                  # ONE tree: the ranks share its taxa
        import zlib
        sig = float(zlib.crc32(tr.parent.tobytes() + tr.branchlength.tobytes()))
        t = torch.tensor([sig, -sig], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        trees_identical = bool(t[0].item() == -t[1].item())
        dist_info = lib.dist_info()

    # one extra profiled pass: per-kernel device time from CUDA events on the launching stream
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
                print("[bench] profiled pass  %-24s %7d events %9.2f ms  %8.2f us/event" % (nm, cnt, ms, 1e3 * ms / cnt), file=sys.stderr)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # per-kernel table of the profiled pass (CUDA events on the launching stream around every launch)
    kernels = {}
    prof_ms_total = sum(prof["msKernel"])
    for nm, ms, cnt, by in zip(api.KERNEL_NAMES, prof["msKernel"], prof["nKernel"], prof["bytesKernel"]):
        if cnt:
            kernels[nm] = {"launches": int(cnt), "ms": round(ms, 2), "us_per_launch": round(1e3 * ms / cnt, 2), "share_of_device_time": round(ms / prof_ms_total, 3),
                           "algorithmic_gb": round(by / 1e9, 3), "gbps": round(by / ms / 1e6, 1) if by and ms > 0 else None}
    traffic = {}
    try:
        traffic = json.load(open(NCU_TRAFFIC)).get(args.workload, {})
    except Exception:
        pass
    dist_kernels = [k for k in kernels if kernels[k]["algorithmic_gb"]]
    dom = max(dist_kernels, key=lambda k: kernels[k]["ms"])
    achieved = kernels[dom]["gbps"]
    roofline = {"bound": "hbm", "kernel": dom + " -- the profile-moving kernel with the largest share of the step's device time (%.0f %%)" % (100 * kernels[dom]["share_of_device_time"]),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json (burst copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": int(1e9 * kernels[dom]["algorithmic_gb"] / kernels[dom]["launches"]),
                "us_per_launch": kernels[dom]["us_per_launch"],
                "bytes_rule": "SURVEY 8d dense rule: L B per leaf, L*(A*4+4+1) B per internal node, the list's query once.  An internal profile holds a "
                              "vector only where its code is NOCODE, so DRAM traffic is BELOW this figure (ncu, profiles/): the fraction is an upper bound "
                              "of HBM utilisation; the kernels are bound by the ordered double-precision sums (one dependent DADD per position), not by HBM",
                "traffic": traffic.get(dom, {}).get("dram_bytes_per_launch"),
                "traffic_note": traffic.get(dom, {}).get("source", "no ncu capture of this kernel is committed under profiles/"),
                "distance_sweep_all_kernels": {"algorithmic_bytes_per_step": prof["distBytes"], "ms_per_step": prof["msDist"],
                                               "gbps": (prof["distBytes"] / (prof["msDist"] * 1e-3)) / 1e9 if prof["msDist"] > 0 else None}}
    if rank != 0:
        if sharded:
            lib.dist_finalize()
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu, parity = None, None
    if args.gpus == 1 and not args.skip_cpu:
        threads, seen = calibrate_threads(wl, chars, host_cores)
        t512, flags512 = avx512_variant(wl, chars, threads) if seen else (None, None)
        if seen:
            t = seen[threads] if t512 is None else min(seen[threads], t512)
            cpu = {"value": n_unique / t, "unit": "taxa/s", "cores": threads, "kind": "reference", "host_cores": host_cores,
                   "sample": "the full workload once per candidate thread count (%d unique taxa x %d columns), unmodified reference binary %s: %s; fastest kept"
                             % (n_unique, wl["pos"], " ".join(wl["ref_flags"]), ", ".join("-threads %d %.2f s" % kv for kv in sorted(seen.items())))
                             + ("" if t512 is None else "; the AVX-512 build (%s) at -threads %d: %.2f s" % (" ".join(flags512), threads, t512))}
        else:
            cpu = {"value": None, "unit": "taxa/s", "cores": host_cores, "kind": "reference", "sample": "oracle/_ref/VeryFastTree not available"}
        parity = parity_check(wl, chars, lib, local_rank)

    value = taxa_total * args.steps / (dev_ms_max * 1e-3)
    line = {"metric": metric, "value": value, "unit": "taxa/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "taxa": int(n_unique), "columns": wl["pos"],
                       "parallelism": ("one tree sharded over %d GPUs: replicated slab and join loop, candidate-sharded sweeps, all-gather of top-2m records / "
                                       "out-distances / merged lists via %s (%d exchanges, %.1f MB per step and rank)"
                                       % (args.gpus, {"peer": "NVLink peer memory (one push-and-wait kernel per exchange)", "nccl": "ncclAllGather",
                                                      "host": "a host all-gather"}.get(dist_info["mode"], dist_info["mode"]),
                                          dist_info["exchanges"] // max(1, args.warmup + args.steps + 1),
                                          dist_info["bytes"] / 1e6 / max(1, args.warmup + args.steps + 1))) if sharded else ("single GPU" if args.gpus == 1 else "replicas x%d: one independent tree per GPU" % args.gpus),
                       "trees_identical_across_ranks": trees_identical, "host_threads_per_rank": host_threads,
                       "openmp_binding": os.environ.get("OMP_PLACES") if os.environ.get("OMP_PROC_BIND") else None,
                       "join_loop": "device-resident" if ptree.stats["counters"]["nKernel"][11] else "host-driven",
                       "l2": "flushed between steps (256 MiB write)",
                       "arithmetic": "f32 storage, f64 accumulation of top/denom and criteria -- the reference's own mix (SURVEY 9.1)"},
            "parity_checked": parity,
            "e2e": {"value": taxa_total * args.steps / e2e_max, "unit": "taxa/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps},
            "gpu_launches": int(launches_total), "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "wall_s": t_wall}
    if sharded:
        lib.dist_finalize()
    if world > 1:
        dist.destroy_process_group()
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
