"""Seeded synthetic workloads for the ntEdit hot path (SURVEY.md §8d recipe).

truth  = iid uniform ACGT (+ optional duplicated segments for repeat content)
draft  = truth with substitutions (rate 1e-3) and indels (rate 1e-4, lengths 1-5) -- the recipe of the
         reference's own demo file name, "ecoliWithMismatches001Indels0001" -- plus optional lower-case
         bases and N runs.
Pure numpy; used by tests and (for small cases) by bench.py.  The GPU-resident generator for the
100 Mbp - 3 Gbp bench workloads lives in bench.py (torch), same recipe.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
SEED = 20261017


def random_genome(n, rng, dup_frac=0.0, dup_len=(1000, 10000)):
    g = ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]
    if dup_frac > 0 and n > 4 * dup_len[0]:
        target = int(n * dup_frac)
        done = 0
        while done < target:
            ln = int(rng.integers(dup_len[0], min(dup_len[1], n // 4)))
            src = int(rng.integers(0, n - ln))
            dst = int(rng.integers(0, n - ln))
            g[dst:dst + ln] = g[src:src + ln].copy()
            done += ln
    return g


def mutate(truth, rng, sub_rate=1e-3, indel_rate=1e-4, max_indel=5, lower_frac=0.0, n_frac=0.0, n_run=(10, 200),
           iupac_frac=0.0):
    """Return the draft as a uint8 array."""
    n = len(truth)
    out = truth.copy()
    # substitutions: rotate within ACGT
    subs = np.flatnonzero(rng.random(n) < sub_rate)
    if len(subs):
        code = np.searchsorted(ACGT, out[subs])  # A,C,G,T are sorted
        out[subs] = ACGT[(code + rng.integers(1, 4, size=len(subs))) % 4]
    # indels
    keep = np.ones(n, dtype=bool)
    ins_after = {}
    sites = np.flatnonzero(rng.random(n) < indel_rate)
    for s in sites:
        ln = int(rng.integers(1, max_indel + 1))
        if rng.random() < 0.5:
            keep[s:s + ln] = False
        else:
            ins_after[int(s)] = ACGT[rng.integers(0, 4, size=ln)]
    if len(sites):
        pieces = []
        last = 0
        for s in sorted(ins_after):
            seg = out[last:s + 1][keep[last:s + 1]]
            pieces.append(seg)
            pieces.append(ins_after[s])
            last = s + 1
        pieces.append(out[last:][keep[last:]])
        out = np.concatenate(pieces)
    n = len(out)
    if iupac_frac > 0:
        idx = np.flatnonzero(rng.random(n) < iupac_frac)
        codes = np.frombuffer(b"RYSWKMBDHV", dtype=np.uint8)
        out[idx] = codes[rng.integers(0, len(codes), size=len(idx))]
    if n_frac > 0:
        covered = 0
        while covered < n * n_frac:
            ln = int(rng.integers(n_run[0], n_run[1]))
            st = int(rng.integers(0, max(1, n - ln)))
            out[st:st + ln] = ord("N")
            covered += ln
    if lower_frac > 0:
        idx = np.flatnonzero(rng.random(n) < lower_frac)
        out[idx] = out[idx] | 0x20
    return out


def write_fasta(path, contigs, width=70):
    """contigs: [(header bytes, seq bytes)]"""
    with open(path, "wb") as fh:
        for hdr, seq in contigs:
            fh.write(b">" + hdr + b"\n")
            if width:
                for i in range(0, len(seq), width):
                    fh.write(seq[i:i + width] + b"\n")
            else:
                fh.write(seq + b"\n")
