"""CPU tests of the oracle (oracle/ntedit_oracle.c): known-answer vectors of the un-vendored btllib ntHash, algebraic
properties of the NTMC64 forms (ntedit.cpp:403-452), and the golden fixtures produced by the unmodified reference."""
import ctypes as C

import numpy as np
import pytest

from tests import cases as tc
from tests import golden_util as gu


# btllib tests/nthash.cpp known-answer vector: seq, k, h and the hashes of the first 3 k-mers
KAT_SEQ, KAT_K, KAT_H = b"ACATGCATGCA", 5, 3
KAT = [
    [0xf59ecb45f0e22b9c, 0x4969c33ac240c129, 0x688d616f0d7e08c3],
    [0x38cc00f940aebdae, 0xab7e1b110e086fc6, 0x011a1818bcfdd553],
    [0x603a48c5a11c794a, 0xe66016e61816b9c4, 0xc5b13cb146996ffe],
]


def test_nthash_known_answers(oracle):
    for i, want in enumerate(KAT):
        _, _, hv = oracle.nthash_kmer(KAT_SEQ[i:i + KAT_K], KAT_H)
        assert hv == want


def test_nthash_canonical_is_strand_symmetric(oracle):
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rng = np.random.default_rng(1)
    for k in (12, 25, 32, 64, 96):
        s = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=k)])
        rc = s.translate(comp)[::-1]
        assert oracle.nthash_kmer(s, 3)[2] == oracle.nthash_kmer(rc, 3)[2]


def test_roll_and_changelast_equal_reseeding(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    u64 = C.c_uint64
    for k, h in ((25, 3), (32, 4), (19, 1), (64, 2)):
        seq = bytes(np.frombuffer(b"ACGTacgtNRY", dtype=np.uint8)[rng.integers(0, 11, size=400)])
        fh, rh = u64(), u64()
        hv = (u64 * h)()
        L.orc_ntmc64_seed(seq[:k], k, h, C.byref(fh), C.byref(rh), hv)
        for i in range(1, len(seq) - k):
            L.orc_ntmc64_roll(seq[i - 1], seq[i + k - 1], k, h, C.byref(fh), C.byref(rh), hv)
            f2, r2, hv2 = oracle.nthash_kmer(seq[i:i + k], h)
            assert (fh.value, rh.value, list(hv)) == (f2, r2, hv2)
            # replace the last base and compare with hashing the edited k-mer from scratch
            f3, r3 = u64(fh.value), u64(rh.value)
            hv3 = (u64 * h)()
            new = b"ACGT"[i % 4]
            L.orc_ntmc64_changelast(seq[i + k - 1], new, k, h, C.byref(f3), C.byref(r3), hv3)
            edited = seq[i:i + k - 1] + bytes([new])
            assert list(hv3) == oracle.nthash_kmer(edited, h)[2]


def test_filter_addressing_and_file_roundtrip(oracle, tmp_path):
    """bit n of a bit filter lives in byte n/8 under mask 1<<(n%8); counters are bytes; header round-trips."""
    for counting in (False, True):
        f = oracle.OracleFilter.new(1000, 25, 3, counting)
        kmer = b"ACGTTGCATGCATGCATTTGACCAG"
        f.insert_seq(kmer)
        hv = oracle.nthash_kmer(kmer, 3)[2]
        data = f.data()
        mod = 1000 if counting else 8000
        for x in hv:
            n = x % mod
            assert data[n] >= 1 if counting else (data[n // 8] >> (n % 8)) & 1
        assert int(np.count_nonzero(data)) <= 3
        p = str(tmp_path / "f.bf")
        f.save(p)
        head = open(p, "rb").read(200)
        assert head.startswith(b"[BTLKmerCountingBloomFilter_v" if counting else b"[BTLKmerBloomFilter_v")
        assert b"[HeaderEnd]\n" in head
        g = oracle.OracleFilter.load(p)
        assert (g.k, g.h, g.nbytes, g.counting) == (25, 3, 1000, counting)
        assert np.array_equal(g.data(), data)
        f.free()
        g.free()


@pytest.mark.parametrize("name", gu.names())
def test_oracle_reproduces_reference_golden(oracle, name):
    """The C restatement gives byte-identical _edited.fa / _changes.tsv / VCF rows to the unmodified reference."""
    g = gu.load(name)
    case = g["case"]
    filt = oracle.OracleFilter.load(g["filter_path"])
    rep = oracle.OracleFilter.load(g["rep_path"]) if g["rep_path"] else None
    assert filt.k == case["k"] and filt.h == case["hash_num"] and filt.counting == case["counting"]
    op = oracle.default_params(filt.k, filt.h, **tc.oracle_param_overrides(case["params"]))
    if rep:
        op.secbf = 1
    fa, tsv, vcf = oracle.polish(g["contigs"], filt, op, bloomrep=rep,
                                 min_contig_len=case["params"].get("min_contig_len", 100))
    assert fa == g["fa"]
    assert tsv == g["tsv"]
    assert vcf == g["vcf"]
    filt.free()
    if rep:
        rep.free()


@pytest.mark.parametrize("ci", [0, 2, 8, 11])
def test_oracle_matches_reference_binary_when_present(oracle, ci, tmp_path):
    """Fresh seeded inputs against oracle/_ref/ntedit_ref (built from /root/reference in the build container)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/ntedit_ref not built")
    import os
    from ntedit_b200 import synth
    case = tc.CASES[ci]
    inp = tc.make_inputs(31 + ci, **dict(case.get("g", {}), n=8000))
    filt, rep = tc.oracle_filters(oracle, inp)
    fpath = str(tmp_path / "f.bf")
    filt.save(fpath)
    rpath = None
    if rep:
        rpath = str(tmp_path / "rep.bf")
        rep.save(rpath)
    dpath = str(tmp_path / "draft.fa")
    synth.write_fasta(dpath, inp["contigs"])
    rfa, rtsv, rvcf = oracle.run_ref(dpath, fpath, workdir=str(tmp_path), extra=case["flags"], rep_path=rpath)
    op = oracle.default_params(inp["k"], inp["h"], **tc.oracle_param_overrides(case["p"]))
    if rep:
        op.secbf = 1
    fa, tsv, vcf = oracle.polish(inp["contigs"], filt, op, bloomrep=rep)
    assert fa == rfa and tsv == rtsv
    assert vcf == b"".join(l for l in rvcf.splitlines(True) if not l.startswith(b"#"))
    assert os.path.getsize(fpath) > inp["fbytes"]
    filt.free()
    if rep:
        rep.free()
