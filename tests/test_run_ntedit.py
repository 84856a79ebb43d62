"""The run-ntedit compatible front end: flags, defaults and the planned command lines, dry-run only."""
import os

from ntedit_b200 import lib, run_ntedit


def run(capsys, argv):
    assert run_ntedit.main(argv) == 0
    return capsys.readouterr().out


def test_polish_plan_mirrors_the_reference_pipeline(capsys, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    for f in ("myReads_1.fq.gz", "myReads_2.fq.gz", "other.fq"):
        (tmp_path / f).write_bytes(b"")
    out = run(capsys, ["polish", "--draft", "myDraft.fa", "--reads", "myReads", "-k", "25", "-n"])
    lines = out.splitlines()
    # intro block of /root/reference/run-ntedit:147-232
    assert lines[:3] == ["Running ntEdit...", "Parameter settings:", "\t--draft myDraft.fa"]
    for want in ("\t--reads myReads", "\t-k 25", "\t-t 4", "\t-z 100", "\t-y 9.0", "\t-j 3", "\t--cutoff 2", "\t-i 5", "\t-d 5", "\t-x 5.0",
                 "\t-m 0", "\t-a 0"):
        assert want in lines
    cmds = [l[len("Running "):] for l in lines if l.startswith("Running ") and l != "Running ntEdit..."]
    assert cmds[0] == "ntcard -k 25 -t 4 -p myReads myReads_1.fq.gz myReads_2.fq.gz"
    assert cmds[1] == "ntstat filter -k 25 -cmin 2 -t 4 -e 0.01 -f myReads_k25.hist -o myReads_k25.bf myReads_1.fq.gz myReads_2.fq.gz"
    # ntedit_run_pipeline.smk:141-152 with ntedit-b200 in place of ntedit
    assert cmds[2] == lib.CLI + " -r myReads_k25.bf -f myDraft.fa -b ntedit_k25 -t 4 -z 100 -i 5 -d 5 -x 5.0 -y 9.0 -m 0 -v 0 -a 0 -j 3 -s 0"
    assert lines[-1] == "Done ntEdit!"


def test_existing_filter_is_not_rebuilt_and_ratio_flags(capsys, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    (tmp_path / "r_k32.hist").write_bytes(b"")
    (tmp_path / "r_k32.bf").write_bytes(b"")
    out = run(capsys, ["polish", "--draft", "d.fa", "--reads", "r", "-k", "32", "-X", "0.4", "-m", "1", "-i", "4", "--gpus", "8", "--solid", "-n"])
    assert "Skipping ntcard: r_k32.hist exists" in out and "Skipping ntstat: r_k32.bf exists" in out
    assert "\t-X 0.4" in out and "\t-Y 0.5" in out and "\t--solid" in out
    assert " -i 4 -d 5 -x 5.0 -y 9.0 -m 1 -v 0 -a 0 -j 3 -X 0.4 -Y 0.5 -s 0 --gpus 8" in out


def test_snv_plans(capsys, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("PATH", str(tmp_path))  # no ntcard: the genome filter is sized from the sequence length
    out = run(capsys, ["snv", "--reference", "ref.fa", "--genome", "hapA.fa", "-k", "35", "-t", "16", "-n"])
    assert "\t--reference ref.fa" in out
    assert "Running " + lib.MAKE_BF + " --genome hapA.fa -k 35 -t 16 -o hapA_k35.bf --num_elements @bases" in out
    assert "Running " + lib.CLI + " -r hapA_k35.bf -f ref.fa -b hapA_ntedit_k35 -t 16 -z 100 -y 9.0 -v 0 -a 0 -j 3 -s 1" in out
    out = run(capsys, ["snv", "--reference", "ref.fa", "--reads", "na12878", "-k", "35", "-n"])
    assert " -r na12878_k35.bf -f ref.fa -b na12878_ntedit_k35 " in out and out.rstrip().splitlines()[-2].endswith("-s 1")


def test_late_bound_filter_size(tmp_path):
    fa = tmp_path / "g.fa"
    fa.write_bytes(b">a x\nACGT\nAC\n>b\nGGG\n")
    assert run_ntedit.resolve(["--num_elements", "@bases"], [str(fa)]) == ["--num_elements", "9"]
    hist = tmp_path / "g.hist"
    hist.write_text("k\tF1\t100\nk\tF0\t77\n")
    assert run_ntedit.resolve(["@F0:%s" % hist], []) == ["77"]
