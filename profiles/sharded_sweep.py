"""Candidate-sharded one-vs-all over NCCL (run with torchrun, one rank per GPU): every rank holds the
replicated leaf slab, evaluates its block of candidates, all-gathers K records per rank and merges.
Checks the merged top-K against the unsharded call and reports the device-side time per sweep."""
import os, sys, time, json
sys.path.insert(0, '.')
import numpy as np
import torch
import torch.distributed as dist
from veryfasttree_b200 import api, synth, dist as vdist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, L = int(sys.argv[1]), int(sys.argv[2])
chars = synth.make_alignment(N, L, "nt", 1)
codes = api.encode(chars, "nt")
lib = api.load()
cfg = api.make_config(N, L, 4, 32, device=local)
with api.Context(lib, cfg) as ctx:
    ctx.upload_leaves(codes)
    ctx.outprofile_rebuild()
    ctx.out_distance_all(N, 0.0)
    K = 2 * int(0.5 + np.sqrt(N))
    ok = True
    for q in (0, N // 2, N - 1):
        full = ctx.dist_one_vs_all(q, N, K)
        sh = vdist.sharded_one_vs_all(ctx, q, N, K, device="cuda")
        ok &= bool(np.array_equal(full[0], sh[0]) and full[3].tobytes() == sh[3].tobytes())
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    reps = 20
    for r in range(reps):
        vdist.sharded_one_vs_all(ctx, (r * 7919) % N, N, K, device="cuda")
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    if rank == 0:
        print(json.dumps({"sharded_one_vs_all": {"ranks": world, "taxa": N, "columns": L, "K": K, "identical_to_unsharded": ok,
                                                 "ms_per_sweep_wall": dt * 1e3}}))
dist.destroy_process_group()
