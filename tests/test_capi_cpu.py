"""CPU tests of the C-ABI boundary: the in-tree CUDA library loads without a GPU, exports every symbol that
include/ntedit_b200.h declares, and every compute entry point fails loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from ntedit_b200 import lib
    if not os.path.exists(lib.SO):
        lib.build()
    return lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ntedit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ntb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(L):
    from ntedit_b200 import lib
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libntedit_b200.so does not export " + n
    # the ctypes binding covers the same set
    assert sorted(lib.SYMBOLS) == names


def test_struct_layouts_match_header(L):
    from ntedit_b200 import lib
    assert C.sizeof(lib.Node) == 20 and C.sizeof(lib.SRec) == 28
    assert C.sizeof(lib.Params) == 15 * 4
    assert C.sizeof(lib.FilterInfo) == 32
    p = lib.Params()
    L.ntb_params_init(C.byref(p))
    # defaults of namespace opt, ntedit.cpp:99-133
    assert (p.jump, p.mode, p.snv, p.mask, p.max_insertions, p.max_deletions) == (3, 0, 0, 0, 5, 5)
    assert (p.edit_threshold, p.missing_threshold, p.edit_ratio, p.missing_ratio) == (9.0, 5.0, 0.5, 0.5)
    assert (p.min_threshold, p.max_threshold, p.min_contig_len, p.use_ratio) == (1, 255, 100, 0)


def test_version_and_tsv_header(L):
    from ntedit_b200 import lib
    assert b"ntedit_b200" in L.ntb_version()
    sb = lib.StrBuf()
    assert L.ntb_format_tsv_header(25, 3, 0, C.byref(sb)) == 0
    line = C.string_at(sb.data, sb.len)
    L.ntb_strbuf_free(C.byref(sb))
    assert line.startswith(b"ID\tbpPosition+1\tOriginalBase\tNewBase\tSupport ")
    assert line.endswith(b"\n") and b"9" in line  # ceil(25/3) = 9 k-mers checked


def test_compute_fails_loudly_without_a_device(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from ntedit_b200 import lib
    assert L.ntb_device_count() == 0
    h = C.c_void_p()
    rc = L.ntb_filter_create(1 << 16, 25, 3, 0, 0, C.byref(h))
    assert rc == lib.C.c_int(-3).value  # NTB_ENODEV
    assert b"no CUDA device" in L.ntb_last_error()
    buf = np.frombuffer(b"ACGT" * 50 + b"\0", dtype=np.uint8).copy()
    offs = np.array([0, len(buf)], dtype=np.uint64)
    b = C.c_void_p()
    rc = L.ntb_batch_upload(buf.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.POINTER(C.c_uint64)), 1, 0, C.byref(b))
    assert rc == -3
    import ntedit_b200 as nb
    with pytest.raises(lib.NtbError):
        nb.BloomFilter.create(1 << 16, 25, 3)


def test_writer_through_the_abi_matches_golden(L, oracle):
    """ntb_format_contig is host code: feed it the oracle's rope / records and compare with the reference's files."""
    from ntedit_b200 import lib
    from tests import cases as tc
    from tests import golden_util as gu
    import ctypes as C2
    g = gu.load("m1")
    filt = oracle.OracleFilter.load(g["filter_path"])
    op = oracle.default_params(filt.k, filt.h, **tc.oracle_param_overrides(g["case"]["params"]))
    fa, tsv, vcf = lib.StrBuf(), lib.StrBuf(), lib.StrBuf()
    assert L.ntb_format_tsv_header(filt.k, op.jump, 0, C.byref(tsv)) == 0
    OL = oracle.lib()
    for hdr, seq in g["contigs"]:
        buf = C2.create_string_buffer(seq, len(seq))
        res = oracle.Result()
        assert OL.orc_polish_contig(buf, len(seq), filt.ptr, None, C2.byref(op), C2.byref(res)) == 0
        nodes = (lib.Node * res.n_nodes)()
        for i in range(res.n_nodes):
            n = res.nodes[i]
            nodes[i].node_type, nodes[i].s_pos, nodes[i].e_pos, nodes[i].num_support, nodes[i].c = \
                n.node_type, n.s_pos, n.e_pos, n.num_support, n.c
        srecs = (lib.SRec * max(1, res.n_srecs))()
        for i in range(res.n_srecs):
            r = res.srecs[i]
            for f in ("pos", "num_support", "altsupp1", "altsupp2", "altsupp3", "draft_char", "sub_base", "altbase1",
                      "altbase2", "altbase3"):
                setattr(srecs[i], f, getattr(r, f))
        assert L.ntb_format_contig(hdr, C2.cast(buf, C2.c_void_p), nodes, res.n_nodes, srecs, res.n_srecs, 0,
                                   C.byref(fa), C.byref(tsv), C.byref(vcf)) == 0
        OL.orc_result_free(C2.byref(res))
    assert C.string_at(fa.data, fa.len) == g["fa"]
    assert C.string_at(tsv.data, tsv.len) == g["tsv"]
    assert (C.string_at(vcf.data, vcf.len) if vcf.len else b"") == g["vcf"]
    for b in (fa, tsv, vcf):
        L.ntb_strbuf_free(C.byref(b))
    filt.free()


def test_cli_builds_and_refuses_to_run_without_a_device(tmp_path):
    """ntedit-b200 (cli.cpp) is built next to the library; without a CUDA device it stops with an error, it never
    falls back to a CPU path."""
    import subprocess
    from ntedit_b200 import lib
    lib.build()
    assert os.path.exists(lib.CLI)
    r = subprocess.run([lib.CLI, "--help"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and b"-f" in r.stderr and b"--gpus" in r.stderr
    r = subprocess.run([lib.CLI], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"need to specify assembly draft file" in r.stderr
    import torch
    if not torch.cuda.is_available():
        d = tmp_path / "d.fa"
        d.write_bytes(b">a\nACGT\n")
        f = tmp_path / "f.bf"
        f.write_bytes(b"[BTLKmerBloomFilter_v1]\n")
        r = subprocess.run([lib.CLI, "-f", str(d), "-r", str(f)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode != 0 and b"no CUDA device" in r.stderr
