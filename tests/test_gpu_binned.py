"""K1b (binned scan: bin_kernel + probe_bin_kernel) must give the same visit bitmap as the direct scan kernel -- checked
end to end: polishing through the binned path is bit-exact against the oracle.  The binned path is normally taken for
filters larger than L2; the environment switches below force it onto the small test filters, with small regions
(many buckets), one-tile chunks (several chunks per batch) and, in one variant, bucket rows so short that most records
take the direct-probe overflow path."""
import numpy as np
import pytest

from tests import cases as tc

pytestmark = pytest.mark.gpu

BINNED_CASES = [c for c in tc.CASES if c["name"] in ("m0_i4_d5", "m1", "m2_i2_d3", "mask", "cbf_m1", "cbf_p2_q200",
                                                     "secondary_filter", "iupac", "high_fpr_m0", "k32_odd_size",
                                                     "short_contigs_z1000", "k64_h4")]
VARIANTS = {
    "regions": {"NTB_BIN_MIN_BYTES": "0", "NTB_BIN_REGION_LOG2": "13"},
    "chunks": {"NTB_BIN_MIN_BYTES": "0", "NTB_BIN_REGION_LOG2": "12", "NTB_BIN_SCRATCH_MB": "1"},
    "overflow": {"NTB_BIN_MIN_BYTES": "0", "NTB_BIN_REGION_LOG2": "14", "NTB_BIN_BUCKET_CAP": "64"},
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("case", BINNED_CASES, ids=[c["name"] for c in BINNED_CASES])
def test_binned_scan_polish_matches_oracle(nb, oracle, monkeypatch, case, variant):
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    inp = tc.make_inputs(5000 + 13 * tc.CASES.index(case), n=45000, **{k: v for k, v in case.get("g", {}).items() if k != "n"})
    ofilt, orep = tc.oracle_filters(oracle, inp)
    bloom = nb.BloomFilter.create(inp["fbytes"], inp["k"], inp["h"], counting=inp["counting"], device=0)
    for t in inp["truths"]:
        for _ in range(inp["cov"]):
            bloom.insert([(b"t", t)])
    rep = None
    if inp["rep_truth"] is not None:
        rep = nb.BloomFilter.create(inp["fbytes"] // 4, inp["k"], inp["h"], counting=False, device=0)
        rep.insert([(b"r", inp["rep_truth"])])
    fa, tsv, vcf, st = nb.polish(inp["contigs"], bloom, nb.default_params(**case["p"]), bloomrep=rep)
    # the binned path launches two kernels per text chunk instead of one scan kernel
    n_tiles = -(-sum(len(s) + 1 for _, s in inp["contigs"]) // 33792)
    want_scan_launches = 2 * (n_tiles if variant == "chunks" else 1)
    assert st["kernel_launches"] >= want_scan_launches + 2
    op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
    if orep:
        op.secbf = 1
    ofa, otsv, ovcf = oracle.polish(inp["contigs"], ofilt, op, bloomrep=orep,
                                    min_contig_len=case["p"].get("min_contig_len", 100))
    assert fa == ofa
    assert tsv == otsv
    assert vcf == ovcf
    assert st["edits"] > 0
    ofilt.free()
    if orep:
        orep.free()


def test_binned_equals_direct_on_a_larger_draft(nb, monkeypatch):
    """Same inputs through the direct and the binned scan: identical outputs and identical site counts."""
    from ntedit_b200 import synth
    rng = np.random.default_rng(2026)
    truth = synth.random_genome(2_000_000, rng, dup_frac=0.05)
    draft = synth.mutate(truth, rng, 1e-3, 1e-4, lower_frac=0.002, n_frac=0.001)
    bloom = nb.BloomFilter.create(3_000_017, 25, 3, device=0)   # odd size: multiply-high remainder path
    bloom.insert([(b"t", truth.tobytes())])
    contigs = [(b"c%d" % i, draft[i * 500_000:(i + 1) * 500_000].tobytes()) for i in range(4)]
    p = nb.default_params(mode=1)
    direct = nb.polish(contigs, bloom, p)
    monkeypatch.setenv("NTB_BIN_MIN_BYTES", "0")
    monkeypatch.setenv("NTB_BIN_REGION_LOG2", "17")
    monkeypatch.setenv("NTB_BIN_SCRATCH_MB", "16")
    binned = nb.polish(contigs, bloom, p)
    assert direct[0] == binned[0] and direct[1] == binned[1] and direct[2] == binned[2]
    assert direct[3]["sites"] == binned[3]["sites"] and direct[3]["edits"] == binned[3]["edits"]
    assert binned[3]["kernel_launches"] > direct[3]["kernel_launches"]
