"""Turns the raw ncu output under gpurun_out/ (profiles/capture.sh) into the committed summaries under profiles/."""
import csv, collections, io, os, subprocess, sys

OUT = os.path.dirname(os.path.abspath(__file__))
RAW = os.path.join(os.path.dirname(OUT), "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def opcode_mix(rep, kernel_index=0):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []; blocks.append(cur); continue
        if cur is not None:
            cur.append(r)
    b = blocks[kernel_index]
    hdr = b[0]
    isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ops, stalls = collections.Counter(), collections.Counter()
    for r in b[1:]:
        try:
            ex, st = int(r[iex]), int(r[ist])
        except Exception:
            continue
        src = r[isrc].split()
        op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        ops[op] += ex; stalls[op] += st
    return ops, stalls


def summarize(rep, title, command, reading, out_name):
    hdr, units, data = raw_page(rep)
    lines = ["# %s" % title, "", "Command (under gpurun, 1 GPU): `%s`" % command, "",
             "| metric | " + " | ".join("launch %d" % (k + 1) for k in range(len(data))) + " | unit |", "|---|" + "---:|" * len(data) + "---|"]
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            lines.append("| %s | %s | %s |" % (m, " | ".join(r[i] for r in data), units[i]))
    ops, stalls = opcode_mix(rep)
    tot, tots = sum(ops.values()), max(1, sum(stalls.values()))
    lines += ["", "SASS mix of launch 1 (instructions executed / warp-stall samples):", "", "| opcode | inst % | stall-sample % |", "|---|---:|---:|"]
    for op, n in ops.most_common(12):
        lines.append("| %s | %.1f | %.1f |" % (op, 100.0 * n / tot, 100.0 * stalls[op] / tots))
    lines += ["", reading, ""]
    open(os.path.join(OUT, out_name), "w").write("\n".join(lines))
    print("wrote", out_name)


def launch_list():
    rows = [r for r in csv.reader(open(os.path.join(RAW, "launches.csv"))) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        nm = r[ki].split("<")[0].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"): v /= 1e3
        elif r[ui] in ("ms", "msecond"): v *= 1e3
        a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ["# Round 1 -- ncu launch list of one full NJ+TopHits step (4 000 taxa x 200 nt, fp32)", "",
             "Command (under gpurun, 1 GPU): `ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/launches.csv python profiles/one_step.py 4000` (profiles/capture.sh).  %d launches.  Per-launch times are cold-cache and serialised: compare SHARES, not absolutes." % sum(a[0] for a in agg.values()), "",
             "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for nm, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.3f | %.1f%% | %.2f |" % (nm, n, us / 1e3, 100 * us / tot, us / n))
    open(os.path.join(OUT, "%s_launches_4000taxa.md" % TAG), "w").write("\n".join(lines) + "\n")
    print("wrote launch list")
    return agg, tot


def ml():
    """k_pair_loglk in its two regimes (profiles/capture_ml.sh)."""
    summarize(os.path.join(RAW, "prof_loglk_round.ncu-rep"), "Round 1 -- `ncu --set full` of `k_pair_loglk<float,20>` inside a lock-step Brent round (aa 4 000 x 1287, JTT, ~70-130 items x 8 warps, warm L2)",
              "ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_pair_loglk -s 600 -c 2 -o gpurun_out/prof_loglk_round python profiles/ml_opt.py 4000 1287 0",
              "Reading: one round of csrc/ml_opt.cpp = the pending pairLogLk of every live branch optimisation: ~70-130 (pair, length) items, one CTA of 8 warps each (W = 8), so at most one CTA per SM on half of the SMs.  The 256 threads compute the 1287 per-site likelihoods in 5-6 strides (20-state dot products with the exp(eigenvalue x rate x length) table in shared memory), then the CTA's first thread runs the reference's sequential product over the sites -- double multiplies with the 1e-4 / 1e4 rescale, whose rounding depends on the order, so it cannot be split (DSETP + BRA + DMUL = the top stall samples).  The launch is LATENCY bound: `sm__cycles_elapsed.max` is one CTA's critical path, issue slots ~14 % busy on the SMs in use, DRAM traffic ~0 (the 4 000 profiles = 436 MB are mostly L2 hits between rounds).  Before the W-warps-per-item change the same round took 156 us (one warp per item).  What bounds a sweep is the number of rounds on the tree's critical path (height x 6 Brent minimisations), ~1 500 per sweep here.", "%s_k_pair_loglk_round_ncu_full.md" % TAG)
    summarize(os.path.join(RAW, "prof_loglk_tree.ncu-rep"), "Round 1 -- `ncu --set full` of `k_pair_loglk<float,20>` on a whole-tree batch (treeLogLk, aa 20 000 x 1287: 19 999 items, one warp each)",
              "ncu --set full --import-source on --clock-control none -k regex:k_pair_loglk -c 2 -o gpurun_out/prof_loglk_tree python profiles/ml_sweep.py 20000 1287",
              "Reading: all pairLogLk terms of treeLogLk in one launch: 19 999 items x 2 profiles x 109 KB = 4.38 GB algorithmic, W = 1 (8 items per CTA, 86 KB of shared memory: two CTAs = 16 warps per SM).  Live CUDA-event time without the profiler: 2.66 ms = 1.65 TB/s = 26 % of the measured 6.45 TB/s HBM peak (was 5.7 ms with 4 items per CTA and a 10 KB table per warp).  DRAM traffic 2.2 GB: half the algorithmic bytes (leaf profiles are 1 byte per site; vectors are fetched only where a site has no known code).  Limiter: each warp's lane 0 runs the 1287-step sequential product while the other 31 lanes idle (DSETP/BRA/DMUL chain), and 16 warps per SM cannot hide it; next step is to give the product chains of several items to the lanes of one warp (transpose through shared memory, as the distance kernels do for their ordered sums).", "%s_k_pair_loglk_tree_ncu_full.md" % TAG)


def late():
    """k_posterior on one tree level and the rebuilt k_outprofile_rebuild (captured late in the round)."""
    summarize(os.path.join(RAW, "prof_posterior.ncu-rep"), "Round 1 -- `ncu --set full` of `k_posterior<float,20>` on one tree level of recomputeMLProfiles (aa 20 000 x 1287, JTT: 1 721 parents, grid 11 x 1 721)",
              "ncu --set full --import-source on --clock-control none -k regex:k_posterior -s 3 -c 1 -o gpurun_out/prof_posterior python profiles/ml_sweep.py 20000 1287",
              "Reading: one thread per site, grid.y = the parents of the level.  Per site the kernel does the reference's three 20 x 20 matrix-vector products (two rotations by exp(eigenvalue x rate x length) into the eigenbasis of the parent, one back through eigeninv: ~2.4 kFLOP) in separately rounded FMUL + FADD -- 63 % of the instruction stream, FMA pipe 32 % busy, issue slots 49 % busy at 4 CTAs of 128 threads per SM (106 registers).  340 MB read + 167 MB written for 1 721 x 3 profiles of 109 KB = 563 MB algorithmic: every byte moves once.  The kernel is ISSUE bound (non-fused multiply-adds are the price of the reference's rounding), with LDG latency (30 % of the stall samples: the 20-float vectors are read per thread, 80-byte stride) as the second limiter; the whole 29-level sweep takes 6.9 ms for 6.6 GB = 0.95 TB/s.", "%s_k_posterior_ncu_full.md" % TAG)
    summarize(os.path.join(RAW, "prof_rebuild.ncu-rep"), "Round 1 -- `ncu --set full` of `k_outprofile_rebuild<float,4,false>` (16 000 x 200 nt, ~15 800 active nodes, cold)",
              "ncu --set full --import-source on --clock-control none -k regex:k_outprofile_rebuild -s 2 -c 1 -o gpurun_out/prof_rebuild python profiles/one_step.py 16000",
              "Reading: 7 CTAs (32 positions each) of 256 threads: the machine is empty by construction -- the reference's out-profile is a P-typed running sum over the nodes in id order (NJ.tcc:741, :771), one sequential chain per (position, state), 16 k links long.  After this round's change the producers (all 256 threads) turn every (node, position) into ready-made addends and the 128 chain threads do one dependent addition per node: live time 0.56 ms per rebuild (1.07 ms before), i.e. ~67 cycles per link, of which the weight chain's float -> double -> float round trip (F2F + DADD + F2F, 7.6 % of the instructions are conversions) is the longest dependency.  The 1.28 M shared-memory bank conflicts of this capture came from the producers' scalar stores into the A-strided addend rows; they were replaced by 128-bit stores afterwards, which did not move the time (573 us): the kernel is chain bound, not store bound.  74 rebuilds per C2 step = 41 ms (3.4 %).", "%s_k_outprofile_rebuild_ncu_full.md" % TAG)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "late":
        late()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "ml":
        ml()
        sys.exit(0)
    launch_list()
    summarize(os.path.join(RAW, "prof_eval_small.ncu-rep"), "Round 1 -- `ncu --set full` of the dominant kernel: `k_eval<float,4,false>`, per-join request lists (16 000 taxa, warm L2)",
              "ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_eval -s 20000 -c 3 -o gpurun_out/prof_eval_small python profiles/one_step.py 16000",
              "Reading: 11-35 CTAs of 4 warps, one warp per request item (G=1), ~10.7 us per launch.  `smsp__cycles_active.avg` is averaged over all 592 sub-partitions; scaled to the 44-140 that hold a warp it is ~10 k cycles (5 us) of dependent latency per warp: two 128-position units, each followed by the ordered 128-term denom/top accumulation (DADD = 27 % of the instructions and 31 % of the stall samples; 8.5 cycles per dependent DADD measured by profiles/ubench/latency.cu), the out-distance algebra with its three divisions, then fence + atomic + the last CTA's copy of the results.  DRAM traffic is 256 B per launch against ~1 MB algorithmic: the 75 MB slab lives in the 126 MB L2.  The launch is LATENCY bound (issue slots 19 % busy on the few SMs in use); nothing here is bandwidth.  What bounds the step is the number of such launches on the join loop's dependency chain (2.2 per join), not this kernel's throughput.", "%s_k_eval_small_ncu_full.md" % TAG)
    summarize(os.path.join(RAW, "prof_batch_nt.ncu-rep"), "Round 1 -- `ncu --set full` of `k_eval<float,4,false>` on one 262 144-pair batch (bandwidth regime, nt 16 000 x 200)",
              "ncu --set full --import-source on --clock-control none -k regex:k_eval -c 1 -o gpurun_out/prof_batch_nt python profiles/big_batch.py 16000 nt 200 262144 1",
              "Reading: 262 144 pairs (1.10 GB algorithmic) in one launch, 8 pairs per warp (tile [8][128], 4 positions per lane).  Live CUDA-event time of the same request without the profiler: 0.41 ms = 2.66 TB/s algorithmic = 41 % of the measured 6.45 TB/s HBM peak (`smsp__cycles_active.avg` 0.80 M cycles = 0.41 ms agrees; the 1.29 ms `gpu__time_duration` is the profiler's replay).  DRAM traffic 30 MB << 1.10 GB algorithmic: the 6 000 internal profiles of this test (26 MB) are read once and then served by L2, so the kernel is not HBM bound here either -- it is ISSUE bound: 476 instructions per (pair x 128 positions) unit, ALU pipe 48 %, issue slots 53 % busy at 12 warps/SM (157 registers).  FSEL/ISETP (the branch-free profileDistPiece selects) and IMAD (addressing) are 50 % of the instruction stream; the ordered DADD chain is 10 %.", "%s_k_eval_batch_nt_ncu_full.md" % TAG)
    summarize(os.path.join(RAW, "prof_batch_aa.ncu-rep"), "Round 1 -- `ncu --set full` of `k_eval<float,20,true>` on one 131 072-pair batch (bandwidth regime, aa 20 000 x 1287, BLOSUM45)",
              "ncu --set full --import-source on --clock-control none -k regex:k_eval -c 1 -o gpurun_out/prof_batch_aa python profiles/big_batch.py 20000 aa 1287 131072 1",
              "Reading: 131 072 pairs x 1287 positions x 20 states (14.35 GB algorithmic), 32 pairs per warp (tile [32][32]).  Live CUDA-event time without the profiler: 4.43 ms = 3.24 TB/s algorithmic = 50 % of the measured HBM peak.  DRAM traffic 1.85 GB: 8x below the algorithmic bytes, because the vectors are fetched only where a position has no known code on that side (dense rows, conditional 128-bit loads) and the 660 MB of internal profiles are partly L2 hits.  Limiter: latency at 12 warps/SM (166 registers -> 3 CTAs): issue slots 37 % busy, ISETP (predicates waiting for the code/weight loads) 25 % of the stall samples.  Next: trim registers to reach 16 warps/SM and stage the vectors of a unit with cp.async.bulk.", "%s_k_eval_batch_aa_ncu_full.md" % TAG)
    summarize(os.path.join(RAW, "prof_wide_aa.ncu-rep"), "Round 1 -- `ncu --set full` of `k_eval_wide<float,20,true>` (one CTA per pair) on per-join lists (aa 4 000 x 1287, warm L2)",
              "ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_eval_wide -s 3000 -c 2 -o gpurun_out/prof_wide_aa python profiles/scale_aa.py 4000 1287",
              "Reading: one CTA of 8 warps per pair; 22-59 pairs per launch at 4 000 taxa, ~20.7 us.  The warps split the 41 chunks of the pair (5-6 units each, loads two units ahead), then lane 0 adds the 1312-term row in order: 1312 x 8.5 cycles = 11 k cycles = 5.6 us of the launch is that single dependent chain (DADD 15 % of instructions).  Before this kernel the same lists ran one warp per pair at ~61 us per launch (profiles/r1_scale_aa.md).  DRAM traffic 0: L2 resident at this size.", "%s_k_eval_wide_aa_ncu_full.md" % TAG)
