"""Shared synthetic parity cases: (reference CLI flags, ntb_params overrides, oracle-param overrides, generator overrides)."""
import numpy as np

from ntedit_b200 import synth

CASES = [
    dict(name="m0_i4_d5", flags=("-i", 4, "-d", 5, "-m", 0), p=dict(max_insertions=4, max_deletions=5, mode=0)),
    dict(name="m1", flags=("-m", 1), p=dict(mode=1)),
    dict(name="m2_i2_d3", flags=("-m", 2, "-i", 2, "-d", 3), p=dict(mode=2, max_insertions=2, max_deletions=3)),
    dict(name="snv", flags=("-s", 1), p=dict(snv=1)),
    dict(name="snv_m2", flags=("-s", 1, "-m", 2), p=dict(snv=1, mode=2), g=dict(n=5000)),
    dict(name="mask", flags=("-a", 1), p=dict(mask=1)),
    dict(name="ratio", flags=("-X", 0.4, "-Y", 0.6), p=dict(use_ratio=1, missing_ratio=0.4, edit_ratio=0.6)),
    dict(name="j2_x4_y7", flags=("-j", 2, "-x", 4, "-y", 7), p=dict(jump=2, missing_threshold=4, edit_threshold=7)),
    dict(name="cbf_m1", flags=("-m", 1), p=dict(mode=1), g=dict(counting=True, cov=3, fbytes=1 << 17)),
    dict(name="cbf_p2_q200", flags=("-m", 1, "-p", 2, "-q", 200), p=dict(mode=1, min_threshold=2, max_threshold=200),
         g=dict(counting=True, cov=3, fbytes=1 << 17)),
    dict(name="cbf_snv", flags=("-s", 1), p=dict(snv=1), g=dict(counting=True, cov=3, fbytes=1 << 17, n=8000)),
    dict(name="secondary_filter", flags=("-m", 0), p=dict(mode=0), g=dict(rep=True)),
    dict(name="iupac", flags=("-m", 1), p=dict(mode=1), g=dict(iupac=0.002)),
    dict(name="high_fpr_m0", flags=("-m", 0), p=dict(mode=0), g=dict(fbytes=1 << 14)),
    dict(name="high_fpr_m2", flags=("-m", 2), p=dict(mode=2), g=dict(fbytes=1 << 14, n=6000)),
    dict(name="k32_odd_size", flags=("-m", 1), p=dict(mode=1), g=dict(k=32, fbytes=100003)),
    dict(name="short_contigs_z1000", flags=("-m", 0, "-z", 1000), p=dict(mode=0, min_contig_len=1000), g=dict(short=True)),
    dict(name="i1_d4_clamp", flags=("-m", 1, "-i", 1, "-d", 4), p=dict(mode=1, max_insertions=1, max_deletions=4)),
    dict(name="i0_d3_clamp", flags=("-m", 0, "-i", 0, "-d", 3), p=dict(mode=0, max_insertions=0, max_deletions=3)),
    dict(name="k64_h4", flags=("-m", 1), p=dict(mode=1), g=dict(k=64, h=4, fbytes=1 << 17)),
]


def oracle_param_overrides(p):
    """ntb_params overrides -> oracle param overrides (the oracle takes the post-CLI values of ntedit.cpp:2411-2493)."""
    o = dict(p)
    if o.get("snv"):
        o["max_insertions"] = 0
        o["max_deletions"] = 0
    i = o.get("max_insertions", 5)
    d = o.get("max_deletions", 5)
    if (i == 0 and d > 0) or (i == 1 and d > 1):
        o["max_deletions"] = i
    o.pop("min_contig_len", None)
    o.pop("segment_len", None)
    return o


def make_inputs(seed, n=20000, k=25, h=3, fbytes=1 << 16, counting=False, sub_rate=2e-3, indel_rate=5e-4, ncontigs=2,
                lower=0.01, nfrac=0.005, iupac=0.0, rep=False, cov=1, short=False):
    """Returns dict(contigs, truths, rep_truth, k, h, fbytes, counting, cov)."""
    rng = np.random.default_rng(seed)
    contigs, truths = [], []
    for c in range(ncontigs):
        truth = synth.random_genome(n, rng, dup_frac=0.05)
        truths.append(truth.tobytes())
        draft = synth.mutate(truth, rng, sub_rate, indel_rate, lower_frac=lower, n_frac=nfrac, iupac_frac=iupac)
        contigs.append((b"ctg%d some comment" % c, draft.tobytes()))
    if short:
        contigs.append((b"tiny", b"ACGTACGTAC" * 30))
        contigs.append((b"k", contigs[0][1][:k]))
        contigs.append((b"kplus", contigs[0][1][:1500]))
    return dict(contigs=contigs, truths=truths, rep_truth=truths[0][: n // 10] if rep else None, k=k, h=h, fbytes=fbytes,
                counting=counting, cov=cov)


def oracle_filters(po, inp):
    filt = po.OracleFilter.new(inp["fbytes"], inp["k"], inp["h"], inp["counting"])
    for t in inp["truths"]:
        for _ in range(inp["cov"]):
            filt.insert_seq(t)
    repf = None
    if inp["rep_truth"] is not None:
        repf = po.OracleFilter.new(inp["fbytes"] // 4, inp["k"], inp["h"], False)
        repf.insert_seq(inp["rep_truth"])
    return filt, repf
