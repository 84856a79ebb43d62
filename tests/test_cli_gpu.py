"""The `ntedit-b200` command line (ntedit_b200/csrc/cli.cpp) against the unmodified reference binary on the same files:
gz multi-line FASTA with comments in, `_edited.fa` / `_changes.tsv` / `_variants.vcf` out, bit-exact (the VCF's
`##fileDate` line aside)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from ntedit_b200 import lib, synth
from tests import cases as tc

pytestmark = pytest.mark.gpu


def run_cli(draft, filt, prefix, extra=(), rep=None):
    cmd = [lib.CLI, "-f", draft, "-r", filt, "-b", prefix] + [str(x) for x in extra]
    if rep:
        cmd += ["-e", rep]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-2000:]
    out = []
    for suffix in ("_edited.fa", "_changes.tsv", "_variants.vcf"):
        with open(prefix + suffix, "rb") as fh:
            out.append(fh.read())
    return out


def strip_date(vcf):
    return b"".join(l for l in vcf.splitlines(True) if not l.startswith(b"##fileDate"))


def write_inputs(nb, tmp_path, inp, gz=True):
    bloom = nb.BloomFilter.create(inp["fbytes"], inp["k"], inp["h"], counting=inp["counting"], device=0)
    for t in inp["truths"]:
        for _ in range(inp["cov"]):
            bloom.insert([(b"t", t)])
    fpath = str(tmp_path / "reads.bf")
    bloom.save(fpath)
    rpath = None
    if inp["rep_truth"] is not None:
        rep = nb.BloomFilter.create(inp["fbytes"] // 4, inp["k"], inp["h"], counting=False, device=0)
        rep.insert([(b"r", inp["rep_truth"])])
        rpath = str(tmp_path / "rep.bf")
        rep.save(rpath)
    plain = str(tmp_path / "draft.fa")
    synth.write_fasta(plain, inp["contigs"])
    dpath = plain
    if gz:
        dpath = plain + ".gz"
        with open(plain, "rb") as src, gzip.open(dpath, "wb") as dst:
            dst.write(src.read())
    return dpath, fpath, rpath


CLI_CASES = [c for c in tc.CASES if c["name"] in ("m0_i4_d5", "m1", "m2_i2_d3", "snv", "mask", "ratio", "cbf_p2_q200",
                                                  "secondary_filter", "short_contigs_z1000", "i1_d4_clamp")]


@pytest.mark.parametrize("case", CLI_CASES, ids=[c["name"] for c in CLI_CASES])
def test_cli_matches_reference_files(nb, oracle, tmp_path, case):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/ntedit_ref not present")
    assert os.path.exists(lib.CLI), "ntedit-b200 is not built"
    inp = tc.make_inputs(77 + tc.CASES.index(case), ncontigs=3, **case.get("g", {}))
    dpath, fpath, rpath = write_inputs(nb, tmp_path, inp)
    # small batches: the contigs go through several ntb_polish_batch calls
    got = run_cli(dpath, fpath, str(tmp_path / "ours"), extra=tuple(case["flags"]) + ("--batch_bases", 15000), rep=rpath)
    rfa, rtsv, rvcf = oracle.run_ref(dpath, fpath, workdir=str(tmp_path), extra=case["flags"], rep_path=rpath)
    assert got[0] == rfa
    assert got[1] == rtsv
    assert strip_date(got[2]) == strip_date(rvcf)


def test_cli_clinvar_annotation_and_default_prefix(nb, oracle, tmp_path):
    """-l cross-references substitutions with a VCF (ntedit.cpp:2261-2274); default output prefix (ntedit.cpp:2496-2502)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/ntedit_ref not present")
    inp = tc.make_inputs(4242, ncontigs=2)
    dpath, fpath, _ = write_inputs(nb, tmp_path, inp, gz=False)
    rfa, rtsv, rvcf = oracle.run_ref(dpath, fpath, workdir=str(tmp_path), extra=("-s", 1))
    rows = [l.split(b"\t") for l in rvcf.splitlines() if l and not l.startswith(b"#")]
    assert len(rows) > 10
    clin = str(tmp_path / "clinvar.vcf")
    with open(clin, "wb") as fh:
        fh.write(b"##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for i, r in enumerate(rows[::3]):
            alt = r[4].split(b",")[0]
            fh.write(b"\t".join([r[0], r[1], b"%d" % (1000 + i), r[3], alt, b".", b".", b"CLNSIG=Pathogenic;N=%d" % i]) + b"\n")
    rfa, rtsv, rvcf = oracle.run_ref(dpath, fpath, workdir=str(tmp_path), extra=("-s", 1, "-l", clin))
    got = run_cli(dpath, fpath, str(tmp_path / "ours"), extra=("-s", 1, "-l", clin))
    assert got[0] == rfa and got[1] == rtsv
    assert strip_date(got[2]) == strip_date(rvcf)
    assert b"CLNSIG=Pathogenic" in got[2]
    # default prefix, in the working directory
    r = subprocess.run([lib.CLI, "-f", dpath, "-r", fpath, "-m", "1", "-k", "25"], cwd=str(tmp_path), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    assert os.path.exists(str(tmp_path / "draft.fa_k25_z100_rreads.bf_i5_d5_m1_edited.fa"))
    assert b"BLOOM::\tcounting: NO" in r.stdout
    # a -k that contradicts the filter header is an error
    r = subprocess.run([lib.CLI, "-f", dpath, "-r", fpath, "-k", "31"], cwd=str(tmp_path), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600)
    assert r.returncode != 0 and b"does not match" in r.stderr


def test_cli_fastq_and_crlf_input(nb, oracle, tmp_path):
    """kseq semantics (lib/kseq.h:175-215): FASTQ records, CRLF line ends, tabs in the header."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/ntedit_ref not present")
    inp = tc.make_inputs(99, ncontigs=2, n=6000)
    dpath, fpath, _ = write_inputs(nb, tmp_path, inp, gz=False)
    (h0, s0), (h1, s1) = inp["contigs"]
    fq = str(tmp_path / "draft.fq")
    with open(fq, "wb") as fh:
        fh.write(b"@" + h0.replace(b" ", b"\t") + b"\r\n" + s0 + b"\r\n+\r\n" + b"I" * len(s0) + b"\r\n")
        fh.write(b"@" + h1 + b"\n" + s1[:3000] + b"\n" + s1[3000:] + b"\n+" + h1 + b"\n" + b"@" * 3000 + b"\n" + b"+" * (len(s1) - 3000) + b"\n")
    got = run_cli(fq, fpath, str(tmp_path / "ours"), extra=("-m", 1))
    rfa, rtsv, rvcf = oracle.run_ref(fq, fpath, workdir=str(tmp_path), extra=("-m", 1))
    assert got[0] == rfa and got[1] == rtsv
    assert strip_date(got[2]) == strip_date(rvcf)


def test_make_bf_cli_builds_the_filter_the_oracle_builds(nb, oracle, tmp_path):
    """ntedit-b200-make-bf (ntedit_make_genome_bf's role, src/ntedit_make_genome_bf.cpp:49-165): FASTA in, btllib filter
    file out, byte-identical to the oracle's builder; sized by --bf, --num_elements or the genome length; then used to
    polish with both our CLI and the reference."""
    import math
    rng = np.random.default_rng(808)
    truths = [synth.random_genome(n, rng) for n in (30000, 24, 25, 12000)]   # one record shorter than k: skipped
    g1 = str(tmp_path / "g1.fa")
    g2 = str(tmp_path / "g2.fa.gz")
    synth.write_fasta(g1, [(b"chr1 x", truths[0].tobytes()), (b"tiny", truths[1].tobytes())])
    plain = str(tmp_path / "g2.fa")
    synth.write_fasta(plain, [(b"k", truths[2].tobytes()), (b"chr2", truths[3].tobytes())], width=0)
    with open(plain, "rb") as src, gzip.open(g2, "wb") as dst:
        dst.write(src.read())

    def run(args):
        r = subprocess.run([lib.MAKE_BF] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, r.stderr.decode(errors="replace")
        return r.stdout.decode()

    out = str(tmp_path / "genome.bf")
    stdout = run(["--genome", g1, g2, "-k", 25, "--bf", 100003, "-o", out])
    assert "BF size (bytes): 100000" in stdout          # whole 64-bit words
    of = oracle.OracleFilter.load(out)
    assert (of.k, of.h, of.counting, of.nbytes) == (25, 3, False, 100000)
    want = oracle.OracleFilter.new(100000, 25, 3, False)
    for t in (truths[0], truths[2], truths[3]):
        want.insert_seq(t.tobytes())
    assert np.array_equal(of.data(), want.data())
    assert ("Bloom filter FPR: %g" % want.fpr()) in stdout

    # sizing from the genome length (Broder & Mitzenmacher, src/ntedit_make_genome_bf.cpp:41-47)
    stdout = run(["--genome", g1, g2, "-k", 25, "--fpr", 0.001, "--hashes", 4, "-o", out])
    n = sum(len(t) for t in truths)
    r = -4 / math.log(1.0 - math.exp(math.log(0.001) / 4))
    expect = int(math.ceil(n * r) / 8) // 8 * 8
    assert ("Genome size (bp): %d" % n) in stdout and ("BF size (bytes): %d" % expect) in stdout
    of2 = oracle.OracleFilter.load(out)
    assert (of2.h, of2.nbytes) == (4, expect)

    # counting variant + polishing through the file with both command lines
    cbf = str(tmp_path / "genome.cbf")
    run(["--genome", g1, g2, "-k", 25, "--num_elements", 50000, "--counting", "-o", cbf])
    draft = synth.mutate(truths[0], rng, 2e-3, 5e-4)
    dpath = str(tmp_path / "draft.fa")
    synth.write_fasta(dpath, [(b"chr1 draft", draft.tobytes())])
    for fpath, flags in ((str(tmp_path / "genome1.bf"), ("-m", 1)), (cbf, ("-m", 2))):
        if not os.path.exists(fpath):
            run(["--genome", g1, "-k", 25, "--bf", 1 << 17, "-o", fpath])
        got = run_cli(dpath, fpath, str(tmp_path / "o"), extra=flags)
        if oracle.have_ref():
            rfa, rtsv, rvcf = oracle.run_ref(dpath, fpath, workdir=str(tmp_path), extra=flags)
            assert got[0] == rfa and got[1] == rtsv and strip_date(got[2]) == strip_date(rvcf)
        assert got[1].count(b"\n") > 20
    of.free()
    of2.free()
    want.free()


@pytest.mark.parametrize("gen", [dict(), dict(k=64, h=4, fbytes=1 << 17)], ids=["k25", "k64_large_rope_kernels"])
def test_cli_shards_batches_over_two_gpus(nb, oracle, tmp_path, gen):
    """--gpus 2: batches of contigs go round-robin to two devices (host threads of one process), each with its own replica
    of the filter -- one file read, one device-to-device copy; the output is still in input order and bit-exact.  The k = 64
    case runs the large-rope kernel instantiations (more dynamic shared memory than the default limit: the attribute
    must be set on BOTH devices).  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/ntedit_ref not present")
    inp = tc.make_inputs(606, ncontigs=7, n=12000, **gen)
    dpath, fpath, _ = write_inputs(nb, tmp_path, inp)
    got = run_cli(dpath, fpath, str(tmp_path / "ours"), extra=("-m", 1, "--gpus", 2, "--batch_bases", 20000))
    rfa, rtsv, rvcf = oracle.run_ref(dpath, fpath, workdir=str(tmp_path), extra=("-m", 1))
    assert got[0] == rfa and got[1] == rtsv and strip_date(got[2]) == strip_date(rvcf)
