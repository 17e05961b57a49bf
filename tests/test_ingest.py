"""vft_ingest: seqsToProfiles' character decoding (NeighbourJoining.tcc:415-457) + Uniquify
(Alignment.cpp:494-526).  CPU: the oracle's restatement against an independent NumPy statement of the same rules;
GPU: the device path (decode + row hashes + gather on the device, grouping confirmed byte for byte) against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import replay  # noqa: E402
from veryfasttree_b200 import api, synth  # noqa: E402


def _text(kind, n, L, seed):
    rs = np.random.RandomState(seed)
    chars = synth.make_alignment(n, L, kind, seed)
    # duplicates (adjacent, far apart, repeated), lower case, gaps, unknown characters, '.' left as is
    dup = rs.randint(0, n, size=n // 5)
    chars = np.concatenate([chars, chars[dup], chars[:7], chars[dup[:11]]])
    rows = rs.randint(0, chars.shape[0], size=60)
    cols = rs.randint(0, L, size=60)
    for k, (r, c) in enumerate(zip(rows, cols)):
        chars[r, c] = [ord("-"), ord("x"), ord("N"), ord("?"), ord(chr(chars[r, c]).lower()), ord(".")][k % 6]
    return np.ascontiguousarray(chars[rs.permutation(chars.shape[0])])


def _check(lib, kind, n, L, seed):
    chars = _text(kind, n, L, seed)
    codes, first, to_uniq = api.ingest(chars, kind, lib=lib)
    keep = synth.unique_rows(chars)
    assert np.array_equal(first, keep)
    assert np.array_equal(codes, api.encode(chars[keep], kind))
    assert np.array_equal(chars[first[to_uniq]], chars)          # every row maps to an identical first occurrence
    return codes, first, to_uniq


@pytest.mark.parametrize("kind,n,L", [("nt", 400, 90), ("aa", 300, 257), ("aa", 5, 1)])
def test_oracle_ingest_matches_numpy(kind, n, L):
    replay.ensure_oracle_built()
    _check(api.load(replay.ORACLE_LIB), kind, n, L, 11)


def test_ingest_rejects_bad_arguments():
    replay.ensure_oracle_built()
    for lib in (api.load(replay.ORACLE_LIB), api.Lib(api.PRODUCT_LIB)):
        nu = api.C.c_int64()
        assert lib.dll.vft_ingest(None, 4, 4, b"ACGT", 0, None, None, None, api.C.byref(nu)) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,L", [("nt", 16000, 200), ("aa", 6000, 1287), ("aa", 33, 5)])
def test_device_ingest_matches_oracle(kind, n, L):
    replay.ensure_oracle_built()
    got = _check(api.load(), kind, n, L, 5)
    want = _check(api.load(replay.ORACLE_LIB), kind, n, L, 5)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
