"""Fuzz the C restatement (oracle/ntedit_oracle.c) against the unmodified reference binary
(oracle/_ref/ntedit_ref).  Developer tool; the pytest wrapper is tests/test_oracle_vs_ref.py."""
import os, sys, tempfile, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
from ntedit_b200 import synth


def one_case(seed, n=20000, k=25, h=3, fbytes=1 << 16, counting=False, flags=(), pkw=None, sub_rate=2e-3,
             indel_rate=5e-4, ncontigs=2, lower=0.01, nfrac=0.005, iupac=0.0, rep=False, cov=1, tmp=None):
    rng = np.random.default_rng(seed)
    tmp = tmp or tempfile.mkdtemp(prefix="fz_")
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
        contigs.append((b"ctg%d some comment" % c, draft.tobytes()))
    fpath = os.path.join(tmp, "f.bf")
    filt.save(fpath)
    rpath = None
    if rep:
        rpath = os.path.join(tmp, "rep.bf")
        repf.save(rpath)
    dpath = os.path.join(tmp, "draft.fa")
    synth.write_fasta(dpath, contigs)
    ref = po.run_ref(dpath, fpath, workdir=tmp, extra=flags, rep_path=rpath)
    params = po.default_params(k, h, **(pkw or {}))
    if rep:
        params.secbf = 1
    mine = po.polish(contigs, filt, params, bloomrep=repf)
    ok_fa = ref[0] == mine[0]
    ok_tsv = ref[1] == mine[1]
    vcf_body = b"".join(l for l in ref[2].splitlines(True) if not l.startswith(b"#"))
    ok_vcf = vcf_body == mine[2]
    filt.free()
    if repf:
        repf.free()
    return ok_fa, ok_tsv, ok_vcf, ref, mine, tmp


if __name__ == "__main__":
    cases = [
        dict(flags=("-i", 4, "-d", 5, "-m", 0), pkw=dict(max_insertions=4, max_deletions=5, mode=0)),
        dict(flags=("-m", 1), pkw=dict(mode=1)),
        dict(flags=("-m", 2, "-i", 2, "-d", 3), pkw=dict(mode=2, max_insertions=2, max_deletions=3)),
        dict(flags=("-s", 1), pkw=dict(snv=1, max_insertions=0, max_deletions=0)),
        dict(flags=("-s", 1, "-m", 2), pkw=dict(snv=1, mode=2, max_insertions=0, max_deletions=0), n=5000),
        dict(flags=("-a", 1), pkw=dict(mask=1)),
        dict(flags=("-X", 0.4, "-Y", 0.6), pkw=dict(use_ratio=1, missing_ratio=0.4, edit_ratio=0.6)),
        dict(flags=("-j", 2, "-x", 4, "-y", 7), pkw=dict(jump=2, missing_threshold=4, edit_threshold=7)),
        dict(flags=("-m", 1), pkw=dict(mode=1), counting=True, cov=3, fbytes=1 << 17),
        dict(flags=("-m", 1, "-p", 2, "-q", 200), pkw=dict(mode=1, min_threshold=2, max_threshold=200), counting=True, cov=3, fbytes=1 << 17),
        dict(flags=("-s", 1), pkw=dict(snv=1, max_insertions=0, max_deletions=0), counting=True, cov=3, fbytes=1 << 17, n=8000),
        dict(flags=("-m", 0), pkw=dict(mode=0), rep=True),
        dict(flags=("-m", 1), pkw=dict(mode=1), iupac=0.002),
        dict(flags=("-m", 0), pkw=dict(mode=0), fbytes=1 << 14),   # high FPR
        dict(flags=("-m", 2), pkw=dict(mode=2), fbytes=1 << 14, n=6000),
        dict(flags=("-m", 1), pkw=dict(mode=1), k=32, fbytes=100003),
    ]
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    bad = 0
    for ci, c in enumerate(cases):
        for seed in range(nseeds):
            ok_fa, ok_tsv, ok_vcf, ref, mine, tmp = one_case(1000 * ci + seed, **c)
            status = "OK " if (ok_fa and ok_tsv and ok_vcf) else "BAD"
            print(status, ci, seed, c.get("flags"), "fa", ok_fa, "tsv", ok_tsv, "vcf", ok_vcf,
                  "rows", ref[1].count(b"\n") - 1, tmp if status == "BAD" else "")
            if status == "BAD":
                bad += 1
                for nm, a, b in (("fa", ref[0], mine[0]), ("tsv", ref[1], mine[1])):
                    open(os.path.join(tmp, "ref." + nm), "wb").write(a)
                    open(os.path.join(tmp, "mine." + nm), "wb").write(b)
                open(os.path.join(tmp, "ref.vcf"), "wb").write(ref[2])
                open(os.path.join(tmp, "mine.vcf"), "wb").write(mine[2])
    print("bad:", bad)
