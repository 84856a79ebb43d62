"""The command line tools' input side on the CPU: kseq record semantics (lib/kseq.h:175-215) of fastx.hpp over every byte
source of bytesource.hpp -- plain, gzip (one and several members), BGZF inflated block-parallel."""
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "iocheck", "fastx_dump.cpp")
EXE = os.path.join(HERE, "iocheck", "_build", "fastx_dump")


@pytest.fixture(scope="module")
def dump():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    deps = [SRC] + [os.path.join(HERE, "..", "ntedit_b200", "csrc", f) for f in ("fastx.hpp", "bytesource.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", SRC, "-o", EXE, "-lz", "-pthread"], check=True)

    def run(path, threads=4):
        r = subprocess.run([EXE, path, str(threads)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
        recs = [tuple(l.split(b"\t")) for l in r.stdout.split(b"\n")[:-1]]
        return r.stderr.strip().decode(), recs
    return run


def bgzf(data, block=0xFF00):
    """BGZF as bgzip writes it (SAM spec 4.1): independent gzip members with the BC extra field, then the empty EOF block."""
    out = []
    for o in list(range(0, len(data), block)) + [None]:
        raw = b"" if o is None else data[o:o + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(raw) + c.flush()
        bsize = 12 + 6 + len(comp) + 8 - 1
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + comp +
                   struct.pack("<II", zlib.crc32(raw), len(raw)))
    return b"".join(out)


def make_records(rng, n):
    recs = []
    for i in range(n):
        ln = int(rng.integers(1, 200_000)) if i % 7 else int(rng.integers(0, 5))
        seq = bytes(np.frombuffer(b"ACGTNacgtRY", dtype=np.uint8)[rng.integers(0, 11, size=ln)])
        recs.append((b"ctg%d" % i, b"" if i % 3 == 0 else b"len=%d some  comment" % ln, seq))
    return recs


def fasta_bytes(recs, width, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    parts = []
    for name, comment, seq in recs:
        parts.append(b">" + name + (b" " + comment if comment else b"") + nl)
        for o in range(0, len(seq), width):
            parts.append(seq[o:o + width] + nl)
    return b"".join(parts)


def test_every_byte_source_gives_the_same_records(dump, tmp_path):
    rng = np.random.default_rng(3)
    recs = make_records(rng, 300)          # ~30 MB: several 16 MB chunks, hundreds of BGZF blocks
    data = fasta_bytes(recs, 80)
    want = [r for r in recs]
    files = {}
    files["plain"] = tmp_path / "d.fa"
    files["plain"].write_bytes(data)
    files["gzip"] = tmp_path / "d.fa.gz"
    with gzip.open(files["gzip"], "wb", compresslevel=1) as fh:
        fh.write(data)
    files["gzip_members"] = tmp_path / "d_members.fa.gz"
    with open(files["gzip_members"], "wb") as fh:  # concatenated members cut at arbitrary bytes: still one stream to zlib
        for o in range(0, len(data), 5_000_003):
            fh.write(gzip.compress(data[o:o + 5_000_003], 1))
    files["bgzf"] = tmp_path / "d.fa.bgz"
    files["bgzf"].write_bytes(bgzf(data))
    kinds = {"plain": "plain", "gzip": "gzip", "gzip_members": "gzip", "bgzf": "bgzf"}
    for tag, path in files.items():
        for threads in (1, 4):
            kind, got = dump(str(path), threads)
            assert kind == kinds[tag]
            assert got == want, tag


def test_crlf_fastq_and_odd_headers(dump, tmp_path):
    rng = np.random.default_rng(4)
    recs = make_records(rng, 20)
    p = tmp_path / "crlf.fa"
    p.write_bytes(fasta_bytes(recs, 61, crlf=True))
    assert dump(str(p))[1] == recs
    # FASTQ: quality lines are skipped, '@' / '>' inside them do not start records
    fq = tmp_path / "r.fq"
    fq.write_bytes(b"@r1 first\nACGT\nAC\n+\n@>II\nI>\n@r2\nGG\n+r2\n@@\n")
    assert dump(str(fq))[1] == [(b"r1", b"first", b"ACGTAC"), (b"r2", b"", b"GG")]
    # garbage in front of the first header, tab-separated comment, no trailing newline
    odd = tmp_path / "odd.fa"
    odd.write_bytes(b"junk\n>a\tx y\nAC\n>b\nG")
    assert dump(str(odd))[1] == [(b"a", b"x y", b"AC"), (b"b", b"", b"G")]
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b"")
    assert dump(str(empty))[1] == []
