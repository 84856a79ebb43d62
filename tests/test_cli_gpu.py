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
