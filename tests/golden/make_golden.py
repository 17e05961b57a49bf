"""Regenerates the golden fixtures from the UNMODIFIED reference (run in the build container,
where /root/reference exists; `make -C oracle ref` first).

  *.refdump.bin   kernel-level vectors from the reference's own templates (oracle/refdump.cpp)
  *.tophits.bin   leaf top-hit lists from the reference's own setAllLeafTopHits
  *.mldump.bin    pairLogLk / posteriorProfile of the reference (refdump "ml" mode: JC, GTR, JTT; CAT rates)
  *.bionj.tree    the same with -bionj (BIONJ-weighted joins, NJ.tcc:2921-2966)
  *.nj.tree       the `NJ` line of the reference binary's -log (tree after the metric phase),
                  run with -threads 1 -ext AVX2 [-nt] [-double-precision] -noml -nni 0 -spr 0 -nosupport
  blosum45_f{32,64}.npz   the BLOSUM45 tables as the reference hands them to its kernels

Alignments are not stored: they are regenerated from (n, nPos, kind, seed) in tests/replay.py::CASES.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import replay  # noqa: E402
from veryfasttree_b200 import synth  # noqa: E402

REFDUMP_CASES = ["nt60", "aa60"]
TOPHITS_CASES = ["c1", "aa300"]
TREE_CASES = ["nt60", "aa60", "c1", "nt1000", "aa300"]


BIONJ_CASES = ["nt60", "aa60", "c1", "aa300"]


def ref_tree(fasta, kind, prec, threads=1, bionj=False):
    with tempfile.TemporaryDirectory() as td:
        log = os.path.join(td, "log")
        args = [replay.REF_BIN] + (["-nt"] if kind == "nt" else []) + (["-double-precision"] if prec == 64 else [])
        args += (["-bionj"] if bionj else []) + ["-ext", "AVX2", "-threads", str(threads), "-noml", "-nni", "0", "-spr", "0", "-nosupport", "-log", log, fasta]
        subprocess.run(args, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        for line in open(log):
            if line.startswith("NJ\t"):
                return line.split("\t", 1)[1].strip()
    raise RuntimeError("no NJ line")


def main():
    with tempfile.TemporaryDirectory() as td:
        for name in sorted(set(REFDUMP_CASES + TOPHITS_CASES + TREE_CASES)):
            chars, kind = replay.golden_case(name)
            fasta = os.path.join(td, name + ".fa")
            synth.write_fasta(fasta, chars)
            for prec in (32, 64):
                if name in REFDUMP_CASES:
                    subprocess.run([replay.REFDUMP_BIN, fasta, kind, str(prec),
                                    os.path.join(HERE, "%s_f%d.refdump.bin" % (name, prec))], check=True)
                if name in TOPHITS_CASES:
                    subprocess.run([replay.REFDUMP_BIN, fasta, kind, str(prec),
                                    os.path.join(HERE, "%s_f%d.tophits.bin" % (name, prec)), "tophits"], check=True)
                if name in TREE_CASES:
                    with open(os.path.join(HERE, "%s_f%d.nj.tree" % (name, prec)), "w") as f:
                        f.write(ref_tree(fasta, kind, prec) + "\n")
                if name in BIONJ_CASES:
                    with open(os.path.join(HERE, "%s_f%d.bionj.tree" % (name, prec)), "w") as f:
                        f.write(ref_tree(fasta, kind, prec, bionj=True) + "\n")
                print("golden", name, prec)
        for name, (n, L, kind, seed, model, combos) in replay.ML_CASES.items():
            chars, kind, model = replay.ml_case_chars(name)
            fasta = os.path.join(td, name + ".fa")
            synth.write_fasta(fasta, chars)
            for prec, lvl in combos:
                subprocess.run([replay.REFDUMP_BIN, fasta, kind, str(prec),
                                os.path.join(HERE, "%s_f%d_e%d.mldump.bin" % (name, prec, lvl)), "ml", model, str(lvl)], check=True)
                print("golden", name, prec, lvl)
        for prec in (32, 64):
            d = replay.read_refdump(os.path.join(HERE, "aa60_f%d.refdump.bin" % prec))
            np.savez(os.path.join(HERE, "blosum45_f%d.npz" % prec), distances=d["tables.distances"],
                     eigenval=d["tables.eigenval"], eigentot=d["tables.eigentot"], codeFreq=d["tables.codeFreq"])


if __name__ == "__main__":
    main()
