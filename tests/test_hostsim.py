"""CPU tests of the host logic around the kernels: the engine header the CUDA walker instantiates (engine.h, compiled
for the host as a one-lane warp by tests/hostsim), the segment stitcher, the rope replay and the writer -- against the
golden fixtures of the unmodified reference and against the oracle on seeded inputs."""
import numpy as np
import pytest

from ntedit_b200 import synth
from tests import cases as tc
from tests import golden_util as gu
from tests.hostsim import pyhostsim as hs


@pytest.fixture(autouse=True)
def _cross_check_the_dense_first_pass(monkeypatch):
    """Every polishing call of this file also runs the walker form of the first pre-evaluation pass and fails when a record
    differs from the dense (thread-per-site) form's, site_dense.h vs engine.h: evaluate_site_core."""
    monkeypatch.setenv("HOSTSIM_CHECK_DENSE", "1")


def run_hostsim(contigs, filt, params_kw, rep=None, segment_len=0):
    params = hs.default_params(segment_len=segment_len, **params_kw)
    repa = (rep.data().tobytes(), rep.h, rep.counting) if rep else None
    return hs.polish(contigs, filt.data().tobytes(), filt.k, filt.h, filt.counting, params, rep=repa)


@pytest.mark.parametrize("segment_len", [0, 150])
@pytest.mark.parametrize("name", gu.names())
def test_host_engine_reproduces_reference_golden(oracle, name, segment_len):
    g = gu.load(name)
    filt = oracle.OracleFilter.load(g["filter_path"])
    rep = oracle.OracleFilter.load(g["rep_path"]) if g["rep_path"] else None
    fa, tsv, vcf, st = run_hostsim(g["contigs"], filt, g["case"]["params"], rep=rep, segment_len=segment_len)
    assert fa == g["fa"]
    assert tsv == g["tsv"]
    assert vcf == g["vcf"]
    filt.free()
    if rep:
        rep.free()


@pytest.mark.parametrize("case", tc.CASES, ids=[c["name"] for c in tc.CASES])
def test_host_engine_matches_oracle(oracle, case):
    inp = tc.make_inputs(500 + tc.CASES.index(case), **dict(case.get("g", {}), n=min(case.get("g", {}).get("n", 9000), 9000)))
    filt, rep = tc.oracle_filters(oracle, inp)
    op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
    if rep:
        op.secbf = 1
    ofa, otsv, ovcf = oracle.polish(inp["contigs"], filt, op, bloomrep=rep,
                                    min_contig_len=case["p"].get("min_contig_len", 100))
    for seg in (0, 130):
        fa, tsv, vcf, st = run_hostsim(inp["contigs"], filt, case["p"], rep=rep, segment_len=seg)
        assert fa == ofa and tsv == otsv and vcf == ovcf
    filt.free()
    if rep:
        rep.free()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_neighbouring_errors_at_every_distance(oracle, mode):
    """Two draft errors d bases apart, d = 1..70 (k = 25): the dirty-window paths of the walker (look-ahead, the jump over
    a dirty run, indel ropes) against the oracle.  Kinds: substitution, 1-3 base deletion from / insertion into the draft."""
    rng = np.random.default_rng(17 + mode)
    k, h = 25, 3
    truth = synth.random_genome(1600, rng)
    filt = oracle.OracleFilter.new(1 << 16, k, h, False)
    filt.insert_seq(truth.tobytes())
    op = oracle.default_params(k, h, mode=mode)

    def apply(seq, pos, kind):
        if kind == 0:
            alt = b"ACGT"[(b"ACGT".index(seq[pos]) + 1) % 4]
            return seq[:pos] + bytes([alt]) + seq[pos + 1:]
        if kind in (1, 2):
            return seq[:pos] + seq[pos + kind:]
        return seq[:pos] + b"GATTACA"[: kind - 2] + seq[pos:]

    contigs = []
    for d in range(1, 71):
        for kinds in ((0, 0), (0, 1), (3, 0), (2, 4), (5, 0)):
            s = truth.tobytes()
            s = apply(s, 800 + d, kinds[1])
            s = apply(s, 800, kinds[0])
            contigs.append((b"d%d_%d%d" % (d, kinds[0], kinds[1]), s))
    ofa, otsv, ovcf = oracle.polish(contigs, filt, op)
    fa, tsv, vcf, st = run_hostsim(contigs, filt, dict(mode=mode))
    assert fa == ofa and tsv == otsv and vcf == ovcf
    assert st.edits > 300
    filt.free()


@pytest.mark.parametrize("piece_events", [1, 7])
@pytest.mark.parametrize("name", ["m0_i4_d5", "m2_i2_d3", "snv", "mask", "cbf_p2_q200", "high_fpr_m2", "short_contigs_z1000"])
def test_piecewise_replay_joins_ropes_exactly(oracle, monkeypatch, name, piece_events):
    """The host replays long contigs as independent pieces cut between walker results and joins the ropes
    (polish_driver.hpp).  Tiny pieces + short segments put a cut behind (almost) every result."""
    monkeypatch.setenv("NTB_REPLAY_PIECE_EVENTS", str(piece_events))
    g = gu.load(name)
    filt = oracle.OracleFilter.load(g["filter_path"])
    rep = oracle.OracleFilter.load(g["rep_path"]) if g["rep_path"] else None
    for seg in (0, 110):
        fa, tsv, vcf, st = run_hostsim(g["contigs"], filt, g["case"]["params"], rep=rep, segment_len=seg)
        assert fa == g["fa"]
        assert tsv == g["tsv"]
        assert vcf == g["vcf"]
    filt.free()
    if rep:
        rep.free()


@pytest.mark.parametrize("piece_events", [None, 3])
def test_fragmented_draft_of_tiny_contigs(oracle, monkeypatch, piece_events):
    """Thousands of contigs of 20-600 bp (shorter than k, shorter than -z, event-less, multi-event): the host's flat
    per-contig bookkeeping (accepted results, pieces, rope joins) against the oracle."""
    if piece_events:
        monkeypatch.setenv("NTB_REPLAY_PIECE_EVENTS", str(piece_events))
    monkeypatch.setenv("NTB_HOST_THREADS", "5")
    rng = np.random.default_rng(77)
    k, h = 25, 3
    truth = synth.random_genome(300_000, rng)
    filt = oracle.OracleFilter.new(1 << 20, k, h, False)
    filt.insert_seq(truth.tobytes())
    draft = synth.mutate(truth, rng, 4e-3, 8e-4, lower_frac=0.01, n_frac=0.002)
    contigs = []
    p = 0
    while p < len(draft):
        ln = int(rng.integers(20, 600))
        contigs.append((b"frag%d" % len(contigs), draft[p:p + ln].tobytes()))
        p += ln
    assert len(contigs) > 900
    op = oracle.default_params(k, h, mode=1)
    ofa, otsv, ovcf = oracle.polish(contigs, filt, op)
    fa, tsv, vcf, st = run_hostsim(contigs, filt, dict(mode=1))
    assert fa == ofa and tsv == otsv and vcf == ovcf
    assert st.edits > 500
    filt.free()


@pytest.mark.parametrize("env", [{"HOSTSIM_NO_PRESITE": "1"}, {"HOSTSIM_DENSE": "0"}, {"HOSTSIM_TABLE_SLOTS": "64"}])
def test_pre_evaluation_is_optional(oracle, monkeypatch, env):
    """Records only run ahead of the walk: without the pre-evaluation pass, with its walker form, and with a table so small
    that most records are dropped, the walkers evaluate what is missing themselves and the output does not change."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    for name in ("m1", "m2_i2_d3", "cbf_p2_q200", "secondary_filter", "high_fpr_m0"):
        case = [c for c in tc.CASES if c["name"] == name][0]
        inp = tc.make_inputs(7000 + tc.CASES.index(case), **case.get("g", {}))
        filt, rep = tc.oracle_filters(oracle, inp)
        fa, tsv, vcf, st = run_hostsim(inp["contigs"], filt, case["p"], rep=rep, segment_len=300)
        op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
        if rep:
            op.secbf = 1
        ofa, otsv, ovcf = oracle.polish(inp["contigs"], filt, op, bloomrep=rep, min_contig_len=case["p"].get("min_contig_len", 100))
        assert fa == ofa and tsv == otsv and vcf == ovcf
        filt.free()
        if rep:
            rep.free()


def test_contig_groups_pipeline(oracle, monkeypatch):
    """The contigs of a call are polished as groups -- device phase of one beside the replay of the one before -- that share
    the site table; any cut into groups gives the same bytes (NTB_CONTIG_GROUP_MIN lets tiny inputs form groups)."""
    case = [c for c in tc.CASES if c["name"] == "m1"][0]
    inp = tc.make_inputs(31337, ncontigs=7, n=9000)
    filt, rep = tc.oracle_filters(oracle, inp)
    op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
    ofa, otsv, ovcf = oracle.polish(inp["contigs"], filt, op)
    monkeypatch.setenv("NTB_CONTIG_GROUP_MIN", "1000")
    for groups in ("1", "2", "3", "7", "50"):
        monkeypatch.setenv("NTB_CONTIG_GROUPS", groups)
        fa, tsv, vcf, st = run_hostsim(inp["contigs"], filt, case["p"], segment_len=400)
        assert fa == ofa and tsv == otsv and vcf == ovcf
    filt.free()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_walkers_jump_over_no_edit_chains(oracle, mode):
    """Stretches of the draft that are not in the filter and cannot be fixed (novel sequence, longer than any indel) leave
    long runs of flagged positions whose sites all end without an edit: the first pass's chain rounds tell the records how
    far the walker may jump (SITE_FL_SKIP), and the outputs must not notice -- including the reference's stale site locals,
    which such sites do write (mode 2 reports them)."""
    rng = np.random.default_rng(77 + mode)
    truth = synth.random_genome(60000, rng)
    draft = bytearray(synth.mutate(truth, rng, 1.5e-3, 3e-4).tobytes())
    for start in range(1500, len(draft) - 400, 2500):
        n = int(rng.integers(8, 120))
        draft[start:start + n] = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))   # novel sequence
    contigs = [(b"c0 novel stretches", bytes(draft[:35000])), (b"c1", bytes(draft[35000:]))]
    filt = oracle.OracleFilter.new(1 << 17, 25, 3, False)
    filt.insert_seq(truth.tobytes())
    for seg in (0, 2048):
        fa, tsv, vcf, st = run_hostsim(contigs, filt, dict(mode=mode), segment_len=seg)
        ofa, otsv, ovcf = oracle.polish(contigs, filt, oracle.default_params(25, 3, mode=mode))
        assert fa == ofa and tsv == otsv and vcf == ovcf
        assert st.pad_ > 200, "no chain site was jumped over"
    filt.free()


@pytest.mark.parametrize("streaming", [False, True])
@pytest.mark.parametrize("threads", ["16", "3"])
def test_group_rules_for_many_and_few_host_threads(oracle, monkeypatch, threads, streaming):
    """The contig groups of a call are sized by the host threads it may use and by how fragmented the draft is
    (polish_driver.hpp); every rule must give the same bytes."""
    monkeypatch.setenv("NTB_HOST_THREADS", threads)
    monkeypatch.setenv("NTB_CONTIG_GROUP_MIN", "2000")
    if streaming:
        monkeypatch.setenv("HOSTSIM_STREAMING", "1")   # with few threads: six groups that start small and grow
    rng = np.random.default_rng(5 + int(threads))
    truth = synth.random_genome(90000, rng)
    draft = synth.mutate(truth, rng, 2e-3, 4e-4).tobytes()
    filt = oracle.OracleFilter.new(1 << 17, 25, 3, False)
    filt.insert_seq(truth.tobytes())
    whole = [(b"c%d" % i, draft[i * 30000:(i + 1) * 30000]) for i in range(3)]
    cuts = sorted(int(x) for x in rng.choice(len(draft), 400, replace=False))
    frag = [(b"f%d" % i, draft[a:b]) for i, (a, b) in enumerate(zip([0] + cuts, cuts + [len(draft)])) if b - a > 0]
    for contigs in (whole, frag):
        fa, tsv, vcf, st = run_hostsim(contigs, filt, dict(mode=1))
        ofa, otsv, ovcf = oracle.polish(contigs, filt, oracle.default_params(25, 3, mode=1))
        assert fa == ofa and tsv == otsv and vcf == ovcf
    filt.free()


@pytest.mark.parametrize("mode", [0, 2])
def test_ragged_and_degenerate_contigs(oracle, mode):
    """Lengths around k, contigs without a single valid window (all N, an N every k-1 bases), one repeated base, lower case
    only, IUPAC codes, and an error in the very first / very last window -- next to ordinary contigs in one batch."""
    rng = np.random.default_rng(4242 + mode)
    k = 25
    truth = synth.random_genome(30000, rng)
    filt = oracle.OracleFilter.new(1 << 17, k, 3, False)
    filt.insert_seq(truth.tobytes())
    t = truth.tobytes()
    body = bytearray(synth.mutate(truth, rng, 2e-3, 4e-4).tobytes())
    first_last = bytearray(t[5000:9000])
    first_last[3] = ord("A") if first_last[3] != ord("A") else ord("C")       # inside the first window
    first_last[-2] = ord("G") if first_last[-2] != ord("G") else ord("T")     # inside the last window
    no_window = bytearray(t[1000:1600])
    for i in range(k - 2, len(no_window), k - 1):
        no_window[i] = ord("N")
    contigs = [
        (b"one", t[:1]), (b"kminus1", t[100:100 + k - 1]), (b"k", t[200:200 + k]), (b"kplus1", t[300:300 + k + 1]),
        (b"twok", t[400:400 + 2 * k]), (b"allN", b"N" * 500), (b"no valid window", bytes(no_window)),
        (b"polyA", b"A" * 400), (b"lower", t[2000:2600].lower()), (b"iupac", t[3000:3300] + b"RYSWKM" + t[3306:3700]),
        (b"first and last window", bytes(first_last)), (b"ordinary", bytes(body)),
    ]
    for min_len in (1, 100):
        for seg in (0, 150):
            params = dict(mode=mode, min_contig_len=min_len)
            fa, tsv, vcf, st = run_hostsim(contigs, filt, params, segment_len=seg)
            ofa, otsv, ovcf = oracle.polish(contigs, filt, oracle.default_params(k, 3, mode=mode), min_contig_len=min_len)
            assert fa == ofa and tsv == otsv and vcf == ovcf
    filt.free()
