#!/usr/bin/env python3
"""run-ntedit front end for the B200 path: `python -m ntedit_b200.run_ntedit polish|snv ...`.

Takes the reference wrapper's sub-commands, flags and defaults (/root/reference/run-ntedit:17-252) and plans the same
steps its Snakefile runs (ntedit_run_pipeline.smk:141-270), with `ntedit-b200` / `ntedit-b200-make-bf` standing where
`ntedit` / `ntedit_make_genome_bf` do:

  polish --draft D --reads P -k K      [ntcard, ntstat filter -> P_kK.bf]        ntedit-b200 ... -s 0   -> ntedit_kK_*
  snv --reference R --reads P -k K     [ntcard, ntstat filter -> P_kK.bf]        ntedit-b200 ... -s 1   -> P_ntedit_kK_*
  snv --reference R --genome G.. -k K  ntedit-b200-make-bf -> <prefix>_kK.bf     ntedit-b200 ... -s 1   -> <prefix>_ntedit_kK_*

ntcard and ntstat (read k-mer counting and filter construction) are external tools, as they are for the reference; their
steps are planned with the reference's command lines and run only when the filter file is missing.  Without ntcard, the
genome filter is sized from the total sequence length (an upper bound of the F0 ntcard would report).
`-n/--dry-run` prints the plan; `-f/--force` runs every step regardless of existing outputs.
"""
import argparse
import gzip
import os
import shlex
import shutil
import subprocess
import sys

from . import lib as _lib

NTEDIT_VERSION = "ntEdit v2.1.1 (ntedit_b200 hot path)"


def build_parser():
    parser = argparse.ArgumentParser(description="ntEdit: Fast, lightweight, scalable genome sequence "
                                                 "polishing and SNV detection & annotation (B200 path)",
                                     formatter_class=argparse.RawTextHelpFormatter)
    sub = parser.add_subparsers(help="ntEdit can be run in polishing or SNV modes.", dest="mode")
    pol = sub.add_parser("polish", help="Run ntEdit polishing")
    snv = sub.add_parser("snv", help="Run ntEdit SNV mode")
    pol.add_argument("--draft", required=True, help="Draft genome assembly (FASTA, Multi-FASTA, and/or gzipped), REQUIRED")
    snv.add_argument("--reference", help="Reference genome assembly for SNV calling, REQUIRED")
    snv.add_argument("--draft", help=argparse.SUPPRESS)
    snv.add_argument("--reads", type=str, help="Prefix of input reads file(s) for variant calling")
    snv.add_argument("--genome", nargs="+", help="Genome assembly file(s) for detecting SNV on --reference")
    pol.add_argument("--reads", required=True, help="Prefix of reads file(s), REQUIRED")
    pol.add_argument("-i", default=5, type=int, choices=range(0, 6), help="Maximum number of insertion bases to try [default=5]")
    pol.add_argument("-d", default=5, type=int, choices=range(0, 11), help="Maximum number of deletions bases to try [default=5]")
    pol.add_argument("-x", default=5.000, type=float, help="k/x ratio for the number of k-mers that should be missing [default=5.000]")
    pol.add_argument("--cap", type=float, help="Cap for the number of base insertions that can be made at one position [default=k*1.5]")
    pol.add_argument("-m", default=0, type=int, choices=range(0, 3), help="Mode of editing, range 0-2 [default=0]")
    pol.add_argument("-a", default=0, type=int, choices=range(0, 2), help="Soft masks missing k-mer positions having no fix [default=0]")
    for sp in (pol, snv):
        sp.add_argument("-k", required=True, type=int, help="k-mer size, REQUIRED")
        sp.add_argument("-l", type=str, help="input VCF file with annotated variants (e.g., clinvar.vcf)")
        sp.add_argument("--cutoff", default=2, type=int, help="The minimum coverage of k-mers in output Bloom filter [default=2]")
        sp.add_argument("--solid", action="store_true", default=False, help="Output the solid k-mers [default=False]")
        sp.add_argument("-t", default=4, type=int, help="Number of threads [default=4]")
        sp.add_argument("-z", default=100, type=int, help="Minimum contig length [default=100]")
        sp.add_argument("-y", default=9.000, type=float, help="k/y ratio for the number of edited k-mers that should be present [default=9.000]")
        sp.add_argument("-j", default=3, type=int, help="check every jth k-mer of the subset [default=3]")
        sp.add_argument("-X", default=-1, type=float, help="Ratio of missing k-mers of the subset needed to attempt a fix")
        sp.add_argument("-Y", default=-1, type=float, help="Ratio of present k-mers of the subset needed to accept an edit")
        sp.add_argument("-e", default=0.01, type=float, help="False positive rate for ntStat Bloom filter")
        sp.add_argument("-v", action="store_true", default=False, help="Verbose mode")
        sp.add_argument("-V", "--version", action="version", version=NTEDIT_VERSION)
        sp.add_argument("-n", "--dry-run", action="store_true", help="Print out the commands that will be executed")
        sp.add_argument("-f", "--force", action="store_true", help="Run all steps, regardless of existing output files")
        sp.add_argument("--gpus", default=1, type=int, help="GPUs the polishing step shards its contigs over [default=1]")
    return parser


def reads_files(prefix):
    """Files of the working directory with the reads prefix (ntedit_run_pipeline.smk: reads_files)."""
    ext = (".fq", ".fastq", ".fa", ".fasta", ".fq.gz", ".fastq.gz", ".fa.gz", ".fasta.gz")
    return sorted(f for f in os.listdir(".") if f.startswith(prefix) and f.endswith(ext))


def total_bases(paths):
    n = 0
    for p in paths:
        op = gzip.open if p.endswith(".gz") else open
        with op(p, "rb") as fh:
            for ln in fh:
                if not ln.startswith(b">"):
                    n += len(ln.rstrip())
    return n


def plan(args):
    """[(description, argv, output file the step makes)] in execution order, and the intro lines the reference prints."""
    intro = ["Running ntEdit...", "Parameter settings:"]
    if args.mode == "snv":
        genome = args.reference or args.draft
        if not genome:
            raise SystemExit("Please specify --reference")
        if bool(args.reads) == bool(args.genome):
            raise SystemExit("Please specify --reads OR --genome")
        intro.append("\t--reference %s" % genome)
    else:
        genome = args.draft
        intro.append("\t--draft %s" % genome)
    k, t = args.k, args.t
    steps = []
    if args.mode == "snv" and args.genome:
        intro.append("\t--genome %s" % args.genome)
        gp = os.path.splitext(os.path.basename(args.genome[0]))[0] if len(args.genome) == 1 else "genomes"
        bf = "%s_k%d.bf" % (gp, k)
        prefix = "%s_ntedit_k%d" % (gp, k)
        cmd = [_lib.MAKE_BF, "--genome"] + list(args.genome) + ["-k", str(k), "-t", str(t), "-o", bf]
        if shutil.which("ntcard"):
            steps.append(("ntcard (F0 of the genomes)", ["ntcard", "-t", str(t), "-k", str(k), "-o", "%s.k%d.hist" % (gp, k)] + list(args.genome),
                          "%s.k%d.hist" % (gp, k)))
            cmd += ["--num_elements", "@F0:%s.k%d.hist" % (gp, k)]
        else:
            cmd += ["--num_elements", "@bases"]
        steps.append(("genome Bloom filter (GPU builder)", cmd, bf))
    else:
        intro.append("\t--reads %s" % args.reads)
        bf = "%s_k%d.bf" % (args.reads, k)
        prefix = "ntedit_k%d" % k if args.mode == "polish" else "%s_ntedit_k%d" % (args.reads, k)
        files = reads_files(args.reads)
        hist = "%s_k%d.hist" % (args.reads, k)
        steps.append(("ntcard", ["ntcard", "-k", str(k), "-t", str(t), "-p", args.reads] + files, hist))
        steps.append(("ntstat", ["ntstat", "filter", "-k", str(k), "-cmin", "0" if args.solid else str(args.cutoff), "-t", str(t), "-e",
                                 str(args.e), "-f", hist, "-o", bf] + files, bf))
    intro += ["\t-k %d" % k, "\t-t %d" % t, "\t-z %d" % args.z, "\t-y %s" % args.y, "\t-j %d" % args.j, "\t-e %s" % args.e]
    ratio = []
    if args.X != -1 or args.Y != -1:
        X = 0.5 if args.X == -1 else args.X
        Y = 0.5 if args.Y == -1 else args.Y
        intro += ["\t-X %s" % X, "\t-Y %s" % Y]
        ratio = ["-X", str(X), "-Y", str(Y)]
    if not (args.mode == "snv" and args.genome):
        intro.append("\t--solid" if args.solid else "\t--cutoff %d" % args.cutoff)
    if args.v:
        intro.append("\t-v")
    if args.l:
        if not os.path.isfile(args.l):
            raise FileNotFoundError("VCF file %s not found" % args.l)
        intro.append("\t-l %s" % args.l)
    cmd = [_lib.CLI, "-r", bf, "-f", genome, "-b", prefix, "-t", str(t), "-z", str(args.z)]
    if args.mode == "polish":
        intro += ["\t-i %d" % args.i, "\t-d %d" % args.d, "\t-x %s" % args.x, "\t-m %d" % args.m, "\t-a %d" % args.a]
        cmd += ["-i", str(args.i), "-d", str(args.d), "-x", str(args.x), "-y", str(args.y)]
        if args.cap:
            intro.append("\t--cap %s" % args.cap)
            cmd += ["-c", str(int(args.cap))]
        cmd += ["-m", str(args.m), "-v", "1" if args.v else "0", "-a", str(args.a), "-j", str(args.j)] + ratio + ["-s", "0"]
    else:
        cmd += ["-y", str(args.y), "-v", "1" if args.v else "0", "-a", "0", "-j", str(args.j)] + ratio + ["-s", "1"]
    if args.l:
        cmd += ["-l", args.l]
    if args.gpus > 1:
        cmd += ["--gpus", str(args.gpus)]
    out = prefix + ("_edited.fa" if args.mode == "polish" else "_variants.vcf")
    steps.append(("ntedit-b200", cmd, out))
    return intro, steps


def resolve(argv, genomes):
    """late-bound arguments of the filter builder: @bases = total sequence length, @F0:<hist> = ntcard's distinct k-mer count"""
    out = []
    for a in argv:
        if a == "@bases":
            a = str(total_bases(genomes))
        elif a.startswith("@F0:"):
            with open(a[4:], encoding="utf8") as fh:
                a = next(ln.split("\t")[-1].strip() for ln in fh if "F0" in ln.split("\t"))
        out.append(a)
    return out


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.mode is None:
        parser.print_help()
        return 0
    intro, steps = plan(args)
    print("\n".join(intro), flush=True)
    for what, cmd, out in steps:
        if os.path.exists(out) and not args.force:
            print("Skipping %s: %s exists" % (what, out), flush=True)
            continue
        print("Running %s" % " ".join(shlex.quote(c) for c in cmd), flush=True)
        if args.dry_run:
            continue
        if shutil.which(cmd[0]) is None and not os.path.exists(cmd[0]):
            raise SystemExit("%s is not in PATH (needed to make %s)" % (cmd[0], out))
        ret = subprocess.call(resolve(cmd, args.genome or []))
        if ret != 0:
            raise subprocess.SubprocessError("ntEdit failed - check the logs for the error.")
    print("Done ntEdit!")
    return 0


if __name__ == "__main__":
    sys.exit(main())
