"""Loader for the golden fixtures under tests/golden/ (made by tests/golden/make_golden.py with the unmodified reference)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")


def names():
    return sorted(d for d in os.listdir(GOLDEN_DIR) if os.path.isfile(os.path.join(GOLDEN_DIR, d, "case.json")))


def read_fasta(path):
    """[(header bytes without '>', sequence bytes)] -- headers as kseq re-emits them (name + ' ' + comment)."""
    out = []
    hdr, parts = None, []
    with open(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if hdr is not None:
                    out.append((hdr, b"".join(parts)))
                hdr, parts = line[1:], []
            else:
                parts.append(line)
    if hdr is not None:
        out.append((hdr, b"".join(parts)))
    return out


def load(name):
    d = os.path.join(GOLDEN_DIR, name)
    case = json.load(open(os.path.join(d, "case.json")))
    rd = lambda f: open(os.path.join(d, f), "rb").read()  # noqa: E731
    rep = os.path.join(d, "rep.bf")
    return dict(case=case, dir=d, contigs=read_fasta(os.path.join(d, "draft.fa")), filter_path=os.path.join(d, "filter.bf"),
                rep_path=rep if os.path.exists(rep) else None, fa=rd("edited.fa"), tsv=rd("changes.tsv"),
                vcf=rd("variants.vcf"))
