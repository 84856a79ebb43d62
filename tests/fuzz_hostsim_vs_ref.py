"""Fuzz the CPU simulation of the device engine + stitcher + replay + writer against the unmodified reference binary."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
from ntedit_b200 import synth
from tests.hostsim import pyhostsim as hs

# (reference CLI flags, ntb_params overrides, generator overrides)
CASES = [
    (("-i", 4, "-d", 5, "-m", 0), dict(max_insertions=4, max_deletions=5, mode=0), {}),
    (("-m", 1), dict(mode=1), {}),
    (("-m", 2, "-i", 2, "-d", 3), dict(mode=2, max_insertions=2, max_deletions=3), {}),
    (("-s", 1), dict(snv=1), {}),
    (("-s", 1, "-m", 2), dict(snv=1, mode=2), dict(n=5000)),
    (("-a", 1), dict(mask=1), {}),
    (("-X", 0.4, "-Y", 0.6), dict(use_ratio=1, missing_ratio=0.4, edit_ratio=0.6), {}),
    (("-j", 2, "-x", 4, "-y", 7), dict(jump=2, missing_threshold=4, edit_threshold=7), {}),
    (("-m", 1), dict(mode=1), dict(counting=True, cov=3, fbytes=1 << 17)),
    (("-m", 1, "-p", 2, "-q", 200), dict(mode=1, min_threshold=2, max_threshold=200), dict(counting=True, cov=3, fbytes=1 << 17)),
    (("-s", 1), dict(snv=1), dict(counting=True, cov=3, fbytes=1 << 17, n=8000)),
    (("-m", 0), dict(mode=0), dict(rep=True)),
    (("-m", 1), dict(mode=1), dict(iupac=0.002)),
    (("-m", 0), dict(mode=0), dict(fbytes=1 << 14)),
    (("-m", 2), dict(mode=2), dict(fbytes=1 << 14, n=6000)),
    (("-m", 1), dict(mode=1), dict(k=32, fbytes=100003)),
    (("-m", 0, "-z", 1000), dict(mode=0, min_contig_len=1000), dict(short=True)),
    (("-m", 1, "-i", 1, "-d", 4), dict(mode=1, max_insertions=1, max_deletions=4), {}),
    (("-m", 0, "-i", 0, "-d", 3), dict(mode=0, max_insertions=0, max_deletions=3), {}),
    # novel stretches (not in the filter, too long for any indel): long chains of sites that end without an edit -- the
    # records' jump information (SITE_FL_SKIP) and the stale site locals mode 2 reports
    (("-m", 0), dict(mode=0), dict(novel=True)),
    (("-m", 1), dict(mode=1), dict(novel=True)),
    (("-m", 2), dict(mode=2), dict(novel=True)),
    (("-m", 2, "-i", 0, "-d", 0), dict(mode=2, max_insertions=0, max_deletions=0), dict(novel=True)),
    (("-m", 1), dict(mode=1), dict(novel=True, fbytes=1 << 14)),
]


def make_case(seed, n=20000, k=25, h=3, fbytes=1 << 16, counting=False, sub_rate=2e-3, indel_rate=5e-4, ncontigs=2,
              lower=0.01, nfrac=0.005, iupac=0.0, rep=False, cov=1, short=False, novel=False):
    rng = np.random.default_rng(seed)
    contigs = []
    filt = po.OracleFilter.new(fbytes, k, h, counting)
    repf = po.OracleFilter.new(fbytes // 4, k, h, False) if rep else None
    for c in range(ncontigs):
        truth = synth.random_genome(n, rng, dup_frac=0.05)
        for _ in range(cov):
            filt.insert_seq(truth.tobytes())
        if rep and c == 0:
            repf.insert_seq(truth[: n // 10].tobytes())
        draft = synth.mutate(truth, rng, sub_rate, indel_rate, lower_frac=lower, n_frac=nfrac, iupac_frac=iupac)
        if novel:
            d = bytearray(draft.tobytes())
            for start in range(700, len(d) - 300, 1900):
                m = int(rng.integers(6, 150))
                d[start:start + m] = bytes(rng.choice(list(b"ACGT"), m).astype(np.uint8))
            draft = np.frombuffer(bytes(d), dtype=np.uint8)
        contigs.append((b"ctg%d some comment" % c, draft.tobytes()))
    if short:
        contigs.append((b"tiny", b"ACGTACGTAC" * 30))
        contigs.append((b"k", contigs[0][1][:k]))
        contigs.append((b"kplus", contigs[0][1][:1500]))
    return contigs, filt, repf


def run_case(ci, seed, segment_len=0, tmp=None):
    flags, pkw, gkw = CASES[ci]
    contigs, filt, repf = make_case(1000 * ci + seed, **gkw)
    tmp = tmp or tempfile.mkdtemp(prefix="hz_")
    fpath = os.path.join(tmp, "f.bf")
    filt.save(fpath)
    rpath = None
    if repf:
        rpath = os.path.join(tmp, "rep.bf")
        repf.save(rpath)
    dpath = os.path.join(tmp, "draft.fa")
    synth.write_fasta(dpath, contigs)
    ref = po.run_ref(dpath, fpath, workdir=tmp, extra=flags, rep_path=rpath)
    params = hs.default_params(segment_len=segment_len, **pkw)
    rep = (repf.data().tobytes(), repf.h, repf.counting) if repf else None
    fa, tsv, vcf, st = hs.polish(contigs, filt.data().tobytes(), filt.k, filt.h, filt.counting, params, rep=rep)
    vcf_ref = b"".join(l for l in ref[2].splitlines(True) if not l.startswith(b"#"))
    ok = (ref[0] == fa, ref[1] == tsv, vcf_ref == vcf)
    filt.free()
    if repf:
        repf.free()
    return ok, ref, (fa, tsv, vcf), st, tmp


if __name__ == "__main__":
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    seglens = [0, 128, 300]
    bad = 0
    first_case = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for ci in range(first_case, len(CASES)):
        for seed in range(nseeds):
            for sl in seglens:
                ok, ref, mine, st, tmp = run_case(ci, seed, sl)
                good = all(ok)
                print("OK " if good else "BAD", ci, seed, "seg", sl, CASES[ci][0], ok, "rows", ref[1].count(b"\n") - 1,
                      "rounds", st.rounds, "segs", st.segments, "reruns", st.reruns, "" if good else tmp, flush=True)
                if not good:
                    bad += 1
                    for nm, a, b in (("fa", ref[0], mine[0]), ("tsv", ref[1], mine[1]), ("vcf", ref[2], mine[2])):
                        open(os.path.join(tmp, "ref." + nm), "wb").write(a)
                        open(os.path.join(tmp, "mine." + nm), "wb").write(b)
    print("bad:", bad)
