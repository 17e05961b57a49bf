"""ONE tree sharded over GPUs (csrc/vft_dist.cuh) on real devices: every rank must return the tree of the
unsharded build, bit for bit.

  * one GPU is enough for the HOST-exchange mode: two ranks share cuda:0, the exchange goes through gloo -- this runs the
    compact strided sweeps, the per-rank select, k_rank_merge, k_scatter_outdist and the chunked list merge on the device;
  * with >= 2 GPUs the NCCL mode and the NVLink peer-memory mode (CUDA IPC + k_peer_allgather) run as well.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(kind):
    from veryfasttree_b200 import api, synth
    if kind == "tiny":                      # fewer lists than ranks, shares that are empty on some ranks
        chars = synth.make_alignment(14, 40, "nt", seed=2)
        chars = chars[synth.unique_rows(chars)]
        return api.encode(chars, "nt"), 4, 64, None
    if kind == "nt":
        chars = synth.make_alignment(3000, 200, "nt", seed=5)
        chars = chars[synth.unique_rows(chars)]
        return api.encode(chars, "nt"), 4, 32, None
    chars = synth.make_alignment(2500, 1287, "aa", seed=6)
    chars = chars[synth.unique_rows(chars)]
    z = np.load(os.path.join(ROOT, "tests", "golden", "blosum45_f32.npz"))
    return api.encode(chars, "aa"), 20, 32, [z["distances"], z["eigenval"], z["eigentot"], z["codeFreq"]]


def _worker(rank, world, port, out_dir, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ["VFT_SHARD_MIN"] = "1"        # shard every sweep, however small (the default shards only shares of >= 4096 units)
    os.environ["VFT_XBUF_INIT_KB"] = "4"     # start with a 4 KB exchange buffer: the growth path (re-mapping the peers) runs too
    if mode == "nccl":
        os.environ["VFT_EXCHANGE"] = "nccl"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from veryfasttree_b200 import api, dist as vdist
    lib = api.load()
    device = 0 if mode == "host" else rank
    torch.cuda.set_device(device)
    out = {}
    for kind in ("nt", "aa", "tiny"):
        codes, A, prec, tables = _case(kind)
        info = vdist.init_sharded_host(lib, device) if mode == "host" else vdist.init_sharded(lib, device)
        assert info["world"] == world and info["rank"] == rank
        tree = api.nj_build(codes, A, prec, lib=lib, tables=tables, device=device, host_threads=2)
        info = lib.dist_info()
        lib.dist_finalize()
        out[kind + "_joins"] = tree.joins
        out[kind + "_bl"] = tree.branchlength
        out[kind + "_lth"] = tree.leaf_top_hits
        out[kind + "_exchanges"] = info["exchanges"]
        out[kind + "_mode"] = info["mode"]
        out[kind + "_ms"] = tree.stats["deviceMsResident"]
    np.savez(os.path.join(out_dir, "g%d.npz" % rank), **out)
    dist.destroy_process_group()


def _run(tmp_path, world, mode):
    from veryfasttree_b200 import api
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    lib = api.load()
    for kind in ("nt", "aa", "tiny"):
        codes, A, prec, tables = _case(kind)
        ref = api.nj_build(codes, A, prec, lib=lib, tables=tables, host_threads=2)
        for r in range(world):
            z = np.load(tmp_path / ("g%d.npz" % r))
            assert np.array_equal(z[kind + "_joins"], ref.joins), "join order differs on rank %d (%s, %s)" % (r, kind, mode)
            assert z[kind + "_bl"].tobytes() == ref.branchlength.tobytes()
            assert np.array_equal(z[kind + "_lth"], ref.leaf_top_hits)
            if kind == "tiny":
                continue
            assert int(z[kind + "_exchanges"]) > ref.stats["nSeeds"] + 2 * ref.stats["nRefreshTopHits"]
            if mode != "peer":
                assert str(z[kind + "_mode"]) == mode
            print("[sharded %s x%d %s] rank %d: %.0f ms (unsharded %.0f ms), exchange mode %s, %d exchanges"
                  % (kind, world, mode, r, float(z[kind + "_ms"]), ref.stats["deviceMsResident"], str(z[kind + "_mode"]), int(z[kind + "_exchanges"])))


def test_sharded_tree_host_exchange_two_ranks_one_gpu(tmp_path):
    _run(tmp_path, 2, "host")


def test_sharded_tree_host_exchange_three_ranks_one_gpu(tmp_path):
    _run(tmp_path, 3, "host")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_tree_nccl(tmp_path):
    _run(tmp_path, 2, "nccl")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_tree_peer_memory(tmp_path):
    _run(tmp_path, 2, "peer")
    z = np.load(tmp_path / "g0.npz")
    assert str(z["aa_mode"]) == "peer", "NVLink peer-memory exchange was not enabled (fell back to %s)" % str(z["aa_mode"])
