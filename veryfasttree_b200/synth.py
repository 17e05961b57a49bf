"""Seeded synthetic alignments for the NJ+TopHits hot path.

Shape of the data follows the survey's recipe: a random root sequence over the
alphabet, grown to N taxa by copying an earlier taxon (70 % of the time one of
the most recent `recent` taxa, which creates clades) and substituting every
site with probability `p_sub`; the amino-acid variant additionally punches a
1-20 column gap run into 30 % of the sequences.  Everything is driven by one
`numpy.random.RandomState(seed)` so the same (n, L, kind, seed) gives the same
bytes on every box.
"""
from __future__ import annotations

import numpy as np

NT = b"ACGT"
AA = b"ARNDCQEGHILKMFPSTWYV"  # reference order, src/Constants.h:50


def make_alignment(n: int, n_pos: int, kind: str = "nt", seed: int = 1,
                   p_sub: float = 0.05, recent: int = 300,
                   gap_frac: float | None = None) -> np.ndarray:
    """Returns an (n, n_pos) uint8 array of ASCII characters ('-' for gaps)."""
    alphabet = np.frombuffer(NT if kind == "nt" else AA, dtype=np.uint8)
    a = len(alphabet)
    if gap_frac is None:
        gap_frac = 0.0 if kind == "nt" else 0.3
    rs = np.random.RandomState(seed)
    idx = np.empty((n, n_pos), dtype=np.uint8)      # alphabet indices
    idx[0] = rs.randint(0, a, size=n_pos)
    # draw everything that does not depend on earlier rows up front
    pick_recent = rs.random_sample(n) < 0.7
    u_parent = rs.random_sample(n)
    for t in range(1, n):
        lo = max(0, t - recent) if pick_recent[t] else 0
        parent = lo + int(u_parent[t] * (t - lo))
        row = idx[parent].copy()
        mut = rs.random_sample(n_pos) < p_sub
        nm = int(mut.sum())
        if nm:
            # substitute with a *different* state
            row[mut] = (row[mut] + rs.randint(1, a, size=nm)) % a
        idx[t] = row
    chars = alphabet[idx]
    if gap_frac > 0:
        gapped = np.nonzero(rs.random_sample(n) < gap_frac)[0]
        starts = rs.randint(0, n_pos, size=len(gapped))
        lens = rs.randint(1, 21, size=len(gapped))
        for r, s, l in zip(gapped, starts, lens):
            chars[r, s:min(n_pos, s + l)] = ord("-")
    return chars


def unique_rows(chars: np.ndarray) -> np.ndarray:
    """First-occurrence unique sequences, in input order (what Uniquify keeps,
    /root/reference/src/Alignment.cpp:494-526)."""
    seen = {}
    keep = []
    for i in range(chars.shape[0]):
        k = chars[i].tobytes()
        if k not in seen:
            seen[k] = i
            keep.append(i)
    return np.asarray(keep, dtype=np.int64)


def write_fasta(path: str, chars: np.ndarray, names=None) -> None:
    n = chars.shape[0]
    with open(path, "wb") as f:
        for i in range(n):
            name = names[i] if names is not None else "t%d" % i
            f.write(b">" + name.encode() + b"\n" + chars[i].tobytes() + b"\n")


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, required=True)
    ap.add_argument("--pos", type=int, required=True)
    ap.add_argument("--kind", default="nt", choices=["nt", "aa"])
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    write_fasta(args.out, make_alignment(args.n, args.pos, args.kind, args.seed))
