import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import ntedit_b200 as nb
from ntedit_b200 import synth
rng = np.random.default_rng(1)
truth = synth.random_genome(2_000_000, rng)
bloom = nb.BloomFilter.create(1 << 24, 25, 3, device=0)
bloom.insert([(b"t", truth.tobytes())])
draft = synth.mutate(truth, rng, 1e-3, 1e-4)
# 400k contigs of ~250 bp cut from 50 copies of the draft
pieces = []
L = len(draft)
t0 = time.time()
for rep in range(50):
    cuts = np.sort(rng.choice(L, size=8000, replace=False))
    prev = 0
    for c in cuts:
        if c - prev > 0:
            pieces.append(draft[prev:c].tobytes())
        prev = c
contigs = [(b"c%d" % i, s) for i, s in enumerate(pieces)]
print("contigs", len(contigs), "bases", sum(len(s) for s in pieces), "gen %.1fs" % (time.time() - t0))
buf, offs = nb.pack_contigs(contigs)
p = nb.default_params(mode=1)
for it in range(2):
    b2 = buf.copy()
    t0 = time.time()
    res = nb.kmerize_and_correct(b2, offs, bloom, p)
    dt = time.time() - t0
    st = res.stats().as_dict()
    print("call %.3fs" % dt, {k: st[k] for k in ("bases", "contigs", "segments", "reruns", "rounds", "edits", "ms_scan", "ms_walk", "ms_host")})
    res.free()
# spot-check parity on the first 300 contigs
from oracle import pyoracle as po
of = po.OracleFilter.new(1 << 24, 25, 3, False); of.insert_seq(truth.tobytes())
sub = contigs[:300]
fa, tsv, vcf, st = nb.polish(sub, bloom, p)
ofa, otsv, ovcf = po.polish(sub, of, po.default_params(25, 3, mode=1))
print("parity on 300 contigs:", fa == ofa and tsv == otsv and vcf == ovcf)
