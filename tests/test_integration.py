"""The binding made real: the REFERENCE program -- its own main(), Options, Alignment,
pipeline driver, NNI / SPR / ML phases -- built with integration/reference_b200.patch (`make -C oracle ref_b200`), whose only
functional change is that `-ext B200` routes the NJ + TopHits phase (VeryFastTreeImpl.tcc:140) through vft_nj_build.

The whole-program Newick (every phase downstream of the NJ tree included: any difference in topology or branch lengths of the
NJ phase would propagate) must equal the unmodified reference's at `-ext NONE`, whose per-element lane order is the one
B200Operations' own primitives use for the later phases.
  not gpu : the patched reference linked against the CPU double of the library (same glue, same ABI)
  gpu     : the patched reference linked against the CUDA product library
"""
import os
import subprocess

import pytest

import replay
from veryfasttree_b200 import synth

REF = replay.REF_BIN
B200_CPU = os.path.join(replay.ROOT, "oracle", "_ref", "VeryFastTree_b200cpu")
B200_GPU = os.path.join(replay.ROOT, "oracle", "_ref", "VeryFastTree_b200")


def whole_program(binary, ext, fasta, kind, prec, env=None):
    args = [binary] + (["-nt"] if kind == "nt" else []) + (["-double-precision"] if prec == 64 else []) + ["-ext", ext, "-threads", "1", fasta]
    p = subprocess.run(args, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr[-800:]
    return p.stdout.strip()


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", ["c1", "aa60"])
def test_reference_with_ext_b200_cpu_double(tmp_path, name, prec):
    if not (os.path.exists(REF) and os.path.exists(B200_CPU)):
        pytest.skip("oracle/_ref/VeryFastTree{,_b200cpu} not built (needs /root/reference: make -C oracle ref ref_b200)")
    chars, kind = replay.golden_case(name)
    fa = str(tmp_path / "a.fa")
    synth.write_fasta(fa, chars)
    want = whole_program(REF, "NONE", fa, kind, prec)
    for loop in ("0", "1"):                     # host-driven and device-resident join loop
        got = whole_program(B200_CPU, "B200", fa, kind, prec, env={"VFT_DEVICE_LOOP": loop})
        assert got == want and len(got) > 100


def test_patched_reference_rejects_what_the_backend_does_not_do(tmp_path):
    if not os.path.exists(B200_CPU):
        pytest.skip("oracle/_ref/VeryFastTree_b200cpu not built")
    chars, kind = replay.golden_case("nt60")
    fa = str(tmp_path / "a.fa")
    synth.write_fasta(fa, chars)
    p = subprocess.run([B200_CPU, "-nt", "-slow", "-ext", "B200", fa], capture_output=True, text=True)
    assert p.returncode != 0 and "B200 backend" in (p.stderr + p.stdout)      # main.cpp:673-678 reports std::invalid_argument and exits 1


@pytest.mark.gpu
@pytest.mark.parametrize("name,prec", [("c1", 32), ("c1", 64), ("aa300", 32)])
def test_reference_with_ext_b200_on_the_device(tmp_path, name, prec):
    if not (os.path.exists(REF) and os.path.exists(B200_GPU)):
        pytest.fail("oracle/_ref/VeryFastTree{,_b200} did not travel with the snapshot (make -C oracle ref ref_b200 where /root/reference exists)")
    chars, kind = replay.golden_case(name)
    fa = str(tmp_path / "a.fa")
    synth.write_fasta(fa, chars)
    want = whole_program(REF, "NONE", fa, kind, prec)
    got = whole_program(B200_GPU, "B200", fa, kind, prec)
    assert got == want and len(got) > 100
