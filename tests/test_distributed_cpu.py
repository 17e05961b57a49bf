"""world_size-2 gloo tests (CPU) of the N>1 path: the measurement aggregation bench.py uses for its
replicas, and the candidate-sharded one-vs-all (per-rank block + all_gather + merge) against the
unsharded call.  The ABI implementation behind the ranks is the CPU oracle (tests only)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import replay
    from veryfasttree_b200 import api, synth, dist as vdist
    lib = api.load(replay.ORACLE_LIB)
    # --- replicas: each rank builds its own tree; aggregation = max time, sum taxa
    chars = synth.make_alignment(200 + 40 * rank, 120, "nt", seed=1 + rank)
    chars = chars[synth.unique_rows(chars)]
    tree = api.nj_build(api.encode(chars, "nt"), 4, 32, lib=lib, trace=False)
    dev_ms, e2e, taxa, launches = vdist.aggregate_step_times(100.0 + 10 * rank, 1.0 + rank, float(chars.shape[0]), 7.0)
    # --- sharded one-vs-all on a replicated slab
    chars2 = synth.make_alignment(300, 150, "nt", seed=9)
    codes = api.encode(chars2, "nt")
    cfg = api.make_config(codes.shape[0], codes.shape[1], 4, 64)
    with api.Context(lib, cfg) as ctx:
        ctx.upload_leaves(codes)
        ctx.outprofile_rebuild()
        ctx.out_distance_all(codes.shape[0], 0.0)
        n_active = codes.shape[0]
        for k in range(40):           # some internal nodes so that both kernels' paths are exercised
            ctx.profile_average_update(300 + k, 2 * k, 2 * k + 1, n_active)
            n_active -= 1
        ctx.out_distance_all(n_active, 0.0)
        full = ctx.dist_one_vs_all(320, n_active, 34)
        shard = vdist.sharded_one_vs_all(ctx, 320, n_active, 34)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), agg=np.array([dev_ms, e2e, taxa, launches]), n=chars.shape[0],
             root=tree.root, full_j=full[0], shard_j=shard[0], full_c=full[3], shard_c=shard[3], full_d=full[1], shard_d=shard[1])
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    import replay
    replay.ensure_oracle_built()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    # aggregation: max of the times, sum of the units -- identical on both ranks
    assert np.array_equal(r0["agg"], r1["agg"])
    assert r0["agg"][0] == 110.0 and r0["agg"][1] == 2.0
    assert r0["agg"][2] == float(r0["n"] + r1["n"]) and r0["agg"][3] == 14.0
    assert int(r0["root"]) == 2 * int(r0["n"]) - 3 and int(r1["root"]) == 2 * int(r1["n"]) - 3
    # sharded sweep == unsharded sweep, bit for bit, on every rank
    for r in (r0, r1):
        assert np.array_equal(r["full_j"], r["shard_j"])
        assert r["full_c"].tobytes() == r["shard_c"].tobytes() and r["full_d"].tobytes() == r["shard_d"].tobytes()


def test_candidate_blocks_cover_everything():
    from veryfasttree_b200 import dist as vdist
    for n in (1, 7, 100, 32001):
        for world in (1, 2, 3, 8):
            blocks = [vdist.candidate_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


# ---- ONE tree sharded over the ranks: the CPU double of csrc/vft_dist.cuh behind the same C-ABI, the exchange
#      routed through torch.distributed.all_gather (gloo).  Every rank must return the unsharded tree, bit for bit.
def _sharded_case(kind):
    from veryfasttree_b200 import api, synth
    if kind == "tiny":                      # fewer lists than ranks, shares that are empty on some ranks
        chars = synth.make_alignment(14, 40, "nt", seed=2)
        chars = chars[synth.unique_rows(chars)]
        return api.encode(chars, "nt"), 4, 64, None
    if kind == "nt":
        chars = synth.make_alignment(420, 160, "nt", seed=5)
        chars = chars[synth.unique_rows(chars)]
        return api.encode(chars, "nt"), 4, 32, None
    chars = synth.make_alignment(330, 90, "aa", seed=6)
    chars = chars[synth.unique_rows(chars)]
    z = np.load(os.path.join(ROOT, "tests", "golden", "blosum45_f64.npz"))
    return api.encode(chars, "aa"), 20, 64, [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]


def _sharded_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ["VFT_SHARD_MIN"] = "1"        # shard every sweep, however small (the default shards only shares of >= 4096 units)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import replay
    from veryfasttree_b200 import api, dist as vdist
    lib = api.load(replay.ORACLE_LIB)
    out = {}
    for kind in ("nt", "aa", "tiny"):
        codes, A, prec, tables = _sharded_case(kind)
        info = vdist.init_sharded_host(lib)
        assert info["world"] == world and info["rank"] == rank and info["mode"] == "host"
        tree = api.nj_build(codes, A, prec, lib=lib, tables=tables, host_threads=1)
        info = lib.dist_info()
        lib.dist_finalize()
        out[kind + "_joins"] = tree.joins
        out[kind + "_bl"] = tree.branchlength
        out[kind + "_lth"] = tree.leaf_top_hits
        out[kind + "_exchanges"] = info["exchanges"]
        out[kind + "_refresh"] = tree.stats["nRefreshTopHits"]
    np.savez(os.path.join(out_dir, "s%d.npz" % rank), **out)
    dist.destroy_process_group()


def _run_sharded(tmp_path, world):
    import replay
    from veryfasttree_b200 import api
    replay.ensure_oracle_built()
    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    lib = api.load(replay.ORACLE_LIB)
    for kind in ("nt", "aa", "tiny"):
        codes, A, prec, tables = _sharded_case(kind)
        ref = api.nj_build(codes, A, prec, lib=lib, tables=tables, host_threads=1)
        for r in range(world):
            z = np.load(tmp_path / ("s%d.npz" % r))
            assert np.array_equal(z[kind + "_joins"], ref.joins), "join order differs on rank %d (%s)" % (r, kind)
            assert z[kind + "_bl"].tobytes() == ref.branchlength.tobytes()
            assert np.array_equal(z[kind + "_lth"], ref.leaf_top_hits)
            if kind == "tiny":
                continue
            # the data path really crossed ranks: one exchange per seed, at least two per refresh (out-distances, one-vs-all;
            # the list merge when there are lists), the initial out-distances
            assert int(z[kind + "_exchanges"]) > ref.stats["nSeeds"] + 2 * int(z[kind + "_refresh"])
            assert int(z[kind + "_refresh"]) == ref.stats["nRefreshTopHits"] > 0


def test_one_tree_sharded_over_two_ranks(tmp_path):
    _run_sharded(tmp_path, 2)


def test_one_tree_sharded_over_three_ranks(tmp_path):
    _run_sharded(tmp_path, 3)        # uneven shares: list chunks and strided slots that do not divide


def _threshold_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ["VFT_SHARD_MIN"] = "150"      # some sweeps of this tree reach 150 units per rank, most do not: both paths in one run
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import replay
    from veryfasttree_b200 import api, dist as vdist
    lib = api.load(replay.ORACLE_LIB)
    codes, A, prec, tables = _sharded_case("nt")
    vdist.init_sharded_host(lib)
    tree = api.nj_build(codes, A, prec, lib=lib, tables=tables, host_threads=1)
    info = lib.dist_info()
    lib.dist_finalize()
    np.savez(os.path.join(out_dir, "t%d.npz" % rank), joins=tree.joins, bl=tree.branchlength, exchanges=info["exchanges"])
    dist.destroy_process_group()


def test_sharding_threshold_mixes_sharded_and_replicated_sweeps(tmp_path):
    """The default policy shards a sweep only when every rank's share is large enough; the decision is taken from replicated
    state, so the ranks stay in step whatever mix results."""
    import replay
    from veryfasttree_b200 import api
    replay.ensure_oracle_built()
    mp.spawn(_threshold_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    lib = api.load(replay.ORACLE_LIB)
    codes, A, prec, tables = _sharded_case("nt")
    ref = api.nj_build(codes, A, prec, lib=lib, tables=tables, host_threads=1)
    for r in range(2):
        z = np.load(tmp_path / ("t%d.npz" % r))
        assert np.array_equal(z["joins"], ref.joins) and z["bl"].tobytes() == ref.branchlength.tobytes()
        assert 0 < int(z["exchanges"]) < ref.stats["nSeeds"] + 3 * ref.stats["nRefreshTopHits"]      # some, not all
