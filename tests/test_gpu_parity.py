"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle (C restatement) and, when the
prebuilt oracle/_ref/ntedit_ref travelled to the box, against the unmodified reference binary.  Bit-exact."""
import os
import tempfile

import numpy as np
import pytest

from ntedit_b200 import synth
from tests import cases as tc

pytestmark = pytest.mark.gpu


def device_filters(nb, inp):
    bloom = nb.BloomFilter.create(inp["fbytes"], inp["k"], inp["h"], counting=inp["counting"], device=0)
    for t in inp["truths"]:
        for _ in range(inp["cov"]):
            bloom.insert([(b"t", t)])
    rep = None
    if inp["rep_truth"] is not None:
        rep = nb.BloomFilter.create(inp["fbytes"] // 4, inp["k"], inp["h"], counting=False, device=0)
        rep.insert([(b"r", inp["rep_truth"])])
    return bloom, rep


@pytest.mark.parametrize("case", tc.CASES, ids=[c["name"] for c in tc.CASES])
@pytest.mark.parametrize("segment_len", [0, 160])
def test_polish_matches_oracle_and_reference(nb, oracle, case, segment_len):
    inp = tc.make_inputs(1000 * tc.CASES.index(case) + 1, **case.get("g", {}))
    ofilt, orep = tc.oracle_filters(oracle, inp)
    bloom, rep = device_filters(nb, inp)
    # filter construction parity (K5) -- byte-identical arrays
    assert np.array_equal(ofilt.data(), bloom.download())
    if rep:
        assert np.array_equal(orep.data(), rep.download())

    params = nb.default_params(segment_len=segment_len, **case["p"])
    fa, tsv, vcf, st = nb.polish(inp["contigs"], bloom, params, bloomrep=rep)

    op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
    if orep:
        op.secbf = 1
    ofa, otsv, ovcf = oracle.polish(inp["contigs"], ofilt, op, bloomrep=orep,
                                    min_contig_len=case["p"].get("min_contig_len", 100))
    assert fa == ofa
    assert tsv == otsv
    assert vcf == ovcf

    if oracle.have_ref():
        tmp = tempfile.mkdtemp(prefix="gpupar_")
        fpath = os.path.join(tmp, "f.bf")
        bloom.save(fpath)          # the product's own writer of the btllib format feeds the reference
        rpath = None
        if rep:
            rpath = os.path.join(tmp, "rep.bf")
            rep.save(rpath)
        dpath = os.path.join(tmp, "draft.fa")
        synth.write_fasta(dpath, inp["contigs"])
        rfa, rtsv, rvcf = oracle.run_ref(dpath, fpath, workdir=tmp, extra=case["flags"], rep_path=rpath)
        assert fa == rfa
        assert tsv == rtsv
        assert vcf == b"".join(l for l in rvcf.splitlines(True) if not l.startswith(b"#"))
    ofilt.free()
    if orep:
        orep.free()


@pytest.mark.parametrize("counting", [False, True])
@pytest.mark.parametrize("k,h,fbytes", [(25, 3, 1 << 16), (32, 4, 100003), (19, 1, 77777), (64, 2, 1 << 15)])
def test_scan_matches_oracle(nb, oracle, k, h, fbytes, counting):
    """K1 through the ABI: per-position count and window validity, including contig borders, N runs, lower case."""
    rng = np.random.default_rng(k * 100 + h)
    truth = synth.random_genome(90000, rng)
    draft = synth.mutate(truth, rng, 5e-3, 1e-3, lower_frac=0.02, n_frac=0.01, iupac_frac=0.001)
    contigs = [(b"a", draft[:40000].tobytes()), (b"b", draft[40000:40000 + k - 1].tobytes()), (b"c", b"A"),
               (b"d", draft[45000:].tobytes())]
    ofilt = oracle.OracleFilter.new(fbytes, k, h, counting)
    ofilt.insert_seq(truth.tobytes())
    if counting:
        ofilt.insert_seq(truth[:30000].tobytes())
    bloom = nb.BloomFilter.create(fbytes, k, h, counting=counting, device=0)
    bloom.insert([(b"t", truth.tobytes())])
    if counting:
        bloom.insert([(b"t2", truth[:30000].tobytes())])
    assert np.array_equal(ofilt.data(), bloom.download())
    counts, valid, offs = nb.scan(bloom, contigs)
    for c, (_, s) in enumerate(contigs):
        want = ofilt.scan_counts(s)
        o = int(offs[c])
        got_valid = valid[o:o + len(s)]
        got = counts[o:o + len(s)]
        assert np.array_equal(got_valid, want != 0xFF)  # counts stay far below 255 here, so 0xFF only marks invalid
        inval = ~got_valid
        assert np.array_equal(got[got_valid], want[got_valid])
        assert not got[inval].any()
    # NUL separators are never valid windows
    assert not valid[[int(x) - 1 for x in offs[1:]]].any()
    ofilt.free()


def test_filter_file_roundtrip_and_fpr(nb, oracle, tmp_path):
    rng = np.random.default_rng(5)
    truth = synth.random_genome(50000, rng)
    for counting in (False, True):
        bloom = nb.BloomFilter.create(123457, 25, 3, counting=counting, device=0)
        bloom.insert([(b"t", truth.tobytes())])
        p = str(tmp_path / ("f%d.bf" % counting))
        bloom.save(p)
        of = oracle.OracleFilter.load(p)          # the oracle's reader of the btllib format
        assert of.k == 25 and of.h == 3 and of.counting == counting and of.nbytes == 123457
        assert np.array_equal(of.data(), bloom.download())
        assert abs(of.fpr() - bloom.get_fpr()) < 1e-12
        again = nb.BloomFilter.load(p, device=0)  # and the product's reader
        assert np.array_equal(again.download(), bloom.download())
        assert again.get_k() == 25 and again.get_hash_num() == 3 and again.is_counting() == counting
        of.free()


def test_large_batch_properties(nb, oracle):
    """Size-independent properties at a size the oracle cannot check exhaustively in the test budget:
    polishing an error-free draft is the identity; polishing twice is idempotent on the second pass."""
    rng = np.random.default_rng(11)
    truth = synth.random_genome(3_000_000, rng)
    bloom = nb.BloomFilter.create(1 << 24, 25, 3, device=0)
    bloom.insert([(b"t", truth.tobytes())])
    p = nb.default_params(mode=0)
    fa, tsv, vcf, st = nb.polish([(b"clean", truth.tobytes())], bloom, p)
    assert fa == b">clean\n" + truth.tobytes() + b"\n"
    assert tsv.count(b"\n") == 1 and vcf == b""
    draft = synth.mutate(truth, rng, 1e-3, 1e-4)
    fa1, tsv1, _, st1 = nb.polish([(b"d", draft.tobytes())], bloom, p)
    seq1 = fa1.split(b"\n")[1]
    fa2, tsv2, _, st2 = nb.polish([(b"d", seq1)], bloom, p)
    assert st1["edits"] > 2000
    # a second pass finds (almost) nothing left to edit and never un-does the first pass
    assert st2["edits"] <= st1["edits"] // 50
    # spot-check the big case against the oracle on a prefix
    ofilt = oracle.OracleFilter.new(1 << 24, 25, 3, False)
    ofilt.insert_seq(truth.tobytes())
    ofa, otsv, _ = oracle.polish([(b"d", draft.tobytes())], ofilt, oracle.default_params(25, 3, mode=0))
    assert fa1 == ofa and tsv1 == otsv
    ofilt.free()


def test_piecewise_replay_on_device_results(nb, oracle, monkeypatch):
    """Same as tests/test_hostsim.py::test_piecewise_replay_joins_ropes_exactly, with the walkers' events coming from the GPU."""
    monkeypatch.setenv("NTB_REPLAY_PIECE_EVENTS", "2")
    for name in ("m1", "m2_i2_d3", "high_fpr_m0"):
        case = [c for c in tc.CASES if c["name"] == name][0]
        inp = tc.make_inputs(9000 + tc.CASES.index(case), **case.get("g", {}))
        ofilt, orep = tc.oracle_filters(oracle, inp)
        bloom, rep = device_filters(nb, inp)
        fa, tsv, vcf, st = nb.polish(inp["contigs"], bloom, nb.default_params(segment_len=200, **case["p"]), bloomrep=rep)
        op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
        ofa, otsv, ovcf = oracle.polish(inp["contigs"], ofilt, op, bloomrep=orep)
        assert fa == ofa and tsv == otsv and vcf == ovcf
        ofilt.free()


def test_rerun_rounds_survive_a_moving_event_arena(nb, oracle, monkeypatch):
    """Several walker rounds, and the pinned event arena is re-allocated (moved) by every round after the first: the
    events of segments accepted in earlier rounds must still be the ones replayed (round-1 ADVICE: stored pointers)."""
    monkeypatch.setenv("NTB_NO_BORDER_ADJUST", "1")       # nominal borders: successors start inside dirty stretches
    monkeypatch.setenv("NTB_TEST_TIGHT_EVENT_ARENA", "1")  # no slack in the arena: every later round moves it
    case = [c for c in tc.CASES if c["name"] == "m1"][0]
    inp = tc.make_inputs(4242, n=60000, sub_rate=1.5e-2, indel_rate=3e-3)
    ofilt, orep = tc.oracle_filters(oracle, inp)
    bloom, rep = device_filters(nb, inp)
    fa, tsv, vcf, st = nb.polish(inp["contigs"], bloom, nb.default_params(segment_len=100, **case["p"]), bloomrep=rep)
    assert st["rounds"] >= 2 and st["reruns"] > 0
    op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
    ofa, otsv, ovcf = oracle.polish(inp["contigs"], ofilt, op, bloomrep=orep)
    assert fa == ofa and tsv == otsv and vcf == ovcf
    ofilt.free()


@pytest.mark.parametrize("env", [{"NTB_NO_PRESITE": "1"}, {"NTB_PRESITE_DENSE": "0"}, {"NTB_SITE_TABLE_SLOTS": "64"}],
                         ids=["no_presite", "warp_first_pass", "tiny_table"])
def test_pre_evaluation_is_optional_on_the_device(nb, oracle, monkeypatch, env):
    """The site records only run ahead of the walk: without them, with the warp-per-site form of the first pass, and with a
    table that drops most records, the walkers evaluate the missing sites themselves -- same bytes out."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    for name in ("m1", "m2_i2_d3", "cbf_p2_q200", "secondary_filter", "k64_h4"):
        case = [c for c in tc.CASES if c["name"] == name][0]
        inp = tc.make_inputs(7000 + tc.CASES.index(case), **case.get("g", {}))
        ofilt, orep = tc.oracle_filters(oracle, inp)
        bloom, rep = device_filters(nb, inp)
        fa, tsv, vcf, st = nb.polish(inp["contigs"], bloom, nb.default_params(segment_len=300, **case["p"]), bloomrep=rep)
        op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
        if orep:
            op.secbf = 1
        ofa, otsv, ovcf = oracle.polish(inp["contigs"], ofilt, op, bloomrep=orep)
        assert fa == ofa and tsv == otsv and vcf == ovcf
        ofilt.free()
        if orep:
            orep.free()


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("env", [{}, {"NTB_PRESITE_DENSE": "0"}, {"NTB_NO_BORDER_ADJUST": "1"}], ids=["dense", "warp_first_pass", "nominal_borders"])
def test_walkers_jump_over_no_edit_chains_on_the_device(nb, oracle, monkeypatch, mode, env):
    """Novel stretches of the draft (not in the filter, too long for any indel) leave long runs of flagged positions whose
    sites all end without an edit; the pre-evaluation passes tell the records how far the walker may jump (SITE_FL_SKIP) --
    the dense chain rounds through shuffles, the second pass and the warp form through the chain's first record -- and the
    outputs must not notice, stale site locals of the reference included (mode 2 reports them).  With nominal segment
    borders the jumps also meet task ends and re-run rounds."""
    from ntedit_b200 import synth
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    rng = np.random.default_rng(770 + mode)
    truth = synth.random_genome(120000, rng)
    draft = bytearray(synth.mutate(truth, rng, 1.5e-3, 3e-4).tobytes())
    for start in range(1500, len(draft) - 400, 2500):
        n = int(rng.integers(8, 160))
        draft[start:start + n] = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
    contigs = [(b"c0 novel stretches", bytes(draft[:70000])), (b"c1", bytes(draft[70000:]))]
    ofilt = oracle.OracleFilter.new(1 << 18, 25, 3, False)
    ofilt.insert_seq(truth.tobytes())
    bloom = nb.BloomFilter.create(1 << 18, 25, 3, counting=False, device=0)
    bloom.insert([(b"t", truth.tobytes())])
    ofa, otsv, ovcf = oracle.polish(contigs, ofilt, oracle.default_params(25, 3, mode=mode))
    for seg in (0, 300, 4096):
        fa, tsv, vcf, st = nb.polish(contigs, bloom, nb.default_params(mode=mode, segment_len=seg))
        assert fa == ofa and tsv == otsv and vcf == ovcf
    ofilt.free()


def test_contig_groups_pipeline_on_the_device(nb, oracle, monkeypatch):
    """Contig groups (device phase of one beside the host replay of the one before, one site table for all) on the GPU."""
    case = [c for c in tc.CASES if c["name"] == "m1"][0]
    inp = tc.make_inputs(31337, ncontigs=7, n=9000)
    ofilt, orep = tc.oracle_filters(oracle, inp)
    bloom, rep = device_filters(nb, inp)
    op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
    ofa, otsv, ovcf = oracle.polish(inp["contigs"], ofilt, op)
    monkeypatch.setenv("NTB_CONTIG_GROUP_MIN", "1000")
    for groups in ("1", "2", "7"):
        monkeypatch.setenv("NTB_CONTIG_GROUPS", groups)
        fa, tsv, vcf, st = nb.polish(inp["contigs"], bloom, nb.default_params(segment_len=400, **case["p"]))
        assert fa == ofa and tsv == otsv and vcf == ovcf
    ofilt.free()


def test_concurrent_calls_from_two_host_threads(nb, oracle):
    """The boundary's threading contract (the reference calls kmerizeAndCorrect from every OpenMP thread,
    ntedit.cpp:2242-2245): two host threads polish different batches against the same filter handles at the same time, on
    one device; each checks its own workspace out, and both get the bytes the oracle gives."""
    import threading
    cases = [[c for c in tc.CASES if c["name"] == n][0] for n in ("m1", "k64_h4")]
    jobs = []
    for case in cases:
        for seed in (1, 2):
            inp = tc.make_inputs(500 + seed + 10 * tc.CASES.index(case), ncontigs=3, **case.get("g", {}))
            ofilt, _ = tc.oracle_filters(oracle, inp)
            bloom, _ = device_filters(nb, inp)
            op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
            want = oracle.polish(inp["contigs"], ofilt, op)
            ofilt.free()
            jobs.append((inp, bloom, nb.default_params(segment_len=500, **case["p"]), want))
    results = [None] * len(jobs)
    errors = []

    def worker(i):
        try:
            inp, bloom, params, _ = jobs[i]
            for _ in range(6):  # keep the calls overlapping
                results[i] = nb.polish(inp["contigs"], bloom, params)[:3]
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))
    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (inp, bloom, params, want), got in zip(jobs, results):
        assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]


@pytest.mark.parametrize("counting", [False, True])
def test_torch_builder_matches_the_insert_kernel(nb, counting):
    """bench.py's reference arm fabricates its filter with torch tensor operations only (TorchFilterBuilder: ntHash from
    rotation tables + btllib addressing, written independently of the CUDA code); it must build the very bytes the product's
    filter construction kernel (K5) builds."""
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    rng = np.random.default_rng(17)
    truth = synth.random_genome(400_000, rng, dup_frac=0.05)
    k, h, nbytes = (32, 3, 1 << 19) if counting else (25, 3, 1 << 18)
    bloom = nb.BloomFilter.create(nbytes, k, h, counting=counting, device=0)
    bloom.insert([(b"t", truth.tobytes())])
    dev = torch.device("cuda", 0)
    filt = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    tb = bench.TorchFilterBuilder(filt, nbytes, k, h, counting, dev)
    tb.insert(torch.from_numpy(truth.copy()).to(dev), chunk=100_003)
    assert np.array_equal(bloom.download(), filt.cpu().numpy())
