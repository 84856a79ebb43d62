"""Generates the golden fixtures under tests/golden/ with the UNMODIFIED reference (oracle/_ref/ntedit_ref, i.e.
/root/reference/ntedit.cpp compiled where it lies against oracle/shim -- see oracle/Makefile).

Run in the build container (the GPU box has no /root/reference):  python tests/golden/make_golden.py
Each case directory holds the exact inputs (draft.fa, filter.bf[, rep.bf], case.json with the reference's command-line
flags) and the reference's own outputs (edited.fa, changes.tsv, variants.vcf without its date-stamped header lines).
The drafts are small (2 contigs x ~6 kbp) so the whole set stays well under 1 MB.
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ntedit_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests import cases as tc  # noqa: E402

# a subset of tests/cases.py that covers every mode, both filter types, the -e filter and the IUPAC / clamp corners
GOLDEN = ["m0_i4_d5", "m1", "m2_i2_d3", "snv", "mask", "ratio", "cbf_p2_q200", "cbf_snv", "secondary_filter", "iupac",
          "high_fpr_m2", "k32_odd_size", "short_contigs_z1000", "i0_d3_clamp"]


def main():
    assert po.have_ref(), "oracle/_ref/ntedit_ref missing: run `make -C oracle` in the build container"
    by_name = {c["name"]: c for c in tc.CASES}
    for name in GOLDEN:
        case = by_name[name]
        g = dict(case.get("g", {}))
        g["n"] = min(g.get("n", 6000), 6000)
        g["fbytes"] = min(g.get("fbytes", 1 << 14), 1 << 15)
        inp = tc.make_inputs(7000 + tc.CASES.index(case), **g)
        filt, repf = tc.oracle_filters(po, inp)
        out = os.path.join(HERE, name)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        filt.save(os.path.join(out, "filter.bf"))
        rpath = None
        if repf:
            rpath = os.path.join(out, "rep.bf")
            repf.save(rpath)
        synth.write_fasta(os.path.join(out, "draft.fa"), inp["contigs"])
        tmp = tempfile.mkdtemp(prefix="golden_")
        fa, tsv, vcf = po.run_ref(os.path.join(out, "draft.fa"), os.path.join(out, "filter.bf"), workdir=tmp,
                                  extra=case["flags"], rep_path=rpath)
        shutil.rmtree(tmp, ignore_errors=True)
        open(os.path.join(out, "edited.fa"), "wb").write(fa)
        open(os.path.join(out, "changes.tsv"), "wb").write(tsv)
        open(os.path.join(out, "variants.vcf"), "wb").write(
            b"".join(l for l in vcf.splitlines(True) if not l.startswith(b"#")))
        json.dump({"name": name, "flags": [str(x) for x in case["flags"]], "params": case["p"],
                   "k": inp["k"], "hash_num": inp["h"], "counting": inp["counting"],
                   "reference": "bcgsc/ntEdit v2.1.1 ntedit.cpp, compiled unmodified against oracle/shim, -t 1"},
                  open(os.path.join(out, "case.json"), "w"), indent=1, sort_keys=True)
        print(name, "rows", tsv.count(b"\n") - 1, "fa bytes", len(fa))
        filt.free()
        if repf:
            repf.free()


if __name__ == "__main__":
    main()
