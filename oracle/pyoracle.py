"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes front-end for oracle/_ref/libntedit_oracle.so (our plain-C restatement of ntEdit's hot path,
oracle/ntedit_oracle.c) and a runner for oracle/_ref/ntedit_ref (the UNMODIFIED reference ntedit.cpp
compiled against oracle/shim).  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package ntedit_b200/.
"""
import ctypes as C
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libntedit_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ntedit_ref")


def build(quiet=True):
    """(Re)build the oracle library (and the reference binary when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class Params(C.Structure):
    _fields_ = [("k", C.c_uint), ("h", C.c_uint), ("jump", C.c_uint), ("mode", C.c_int), ("snv", C.c_int),
                ("mask", C.c_int), ("max_insertions", C.c_uint), ("max_deletions", C.c_uint),
                ("edit_threshold", C.c_float), ("missing_threshold", C.c_float), ("edit_ratio", C.c_float),
                ("missing_ratio", C.c_float), ("use_ratio", C.c_int), ("insertion_cap", C.c_uint),
                ("min_threshold", C.c_uint), ("max_threshold", C.c_uint), ("secbf", C.c_int)]


class Filter(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("bytes", C.c_uint64), ("k", C.c_uint), ("h", C.c_uint),
                ("counting", C.c_int)]


class Node(C.Structure):
    _fields_ = [("node_type", C.c_int32), ("s_pos", C.c_uint32), ("e_pos", C.c_uint32),
                ("num_support", C.c_uint32), ("c", C.c_uint8)]


class SRec(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("draft_char", C.c_uint8), ("sub_base", C.c_uint8),
                ("num_support", C.c_uint32), ("altbase1", C.c_uint8), ("altbase2", C.c_uint8),
                ("altbase3", C.c_uint8), ("altsupp1", C.c_uint32), ("altsupp2", C.c_uint32),
                ("altsupp3", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("nodes", C.POINTER(Node)), ("n_nodes", C.c_size_t), ("srecs", C.POINTER(SRec)),
                ("n_srecs", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        u64, u8p, cp = C.c_uint64, C.POINTER(C.c_uint8), C.c_char_p
        L.orc_srol.restype = u64
        L.orc_srol.argtypes = [u64]
        L.orc_sror.restype = u64
        L.orc_sror.argtypes = [u64]
        L.orc_base_forward_hash.restype = u64
        L.orc_base_forward_hash.argtypes = [cp, C.c_uint]
        L.orc_base_reverse_hash.restype = u64
        L.orc_base_reverse_hash.argtypes = [cp, C.c_uint]
        for name in ("orc_ntmc64_roll", "orc_ntmc64_changelast"):
            getattr(L, name).argtypes = [C.c_ubyte, C.c_ubyte, C.c_uint, C.c_uint, C.POINTER(u64), C.POINTER(u64),
                                         C.POINTER(u64)]
            getattr(L, name).restype = None
        L.orc_ntmc64_seed.argtypes = [cp, C.c_uint, C.c_uint, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.orc_ntmc64_seed.restype = None
        L.orc_filter_new.restype = C.POINTER(Filter)
        L.orc_filter_new.argtypes = [u64, C.c_uint, C.c_uint, C.c_int]
        L.orc_filter_free.argtypes = [C.POINTER(Filter)]
        L.orc_filter_free.restype = None
        L.orc_filter_contains.argtypes = [C.POINTER(Filter), C.POINTER(u64)]
        L.orc_filter_count.argtypes = [C.POINTER(Filter), C.POINTER(u64)]
        L.orc_filter_count.restype = C.c_uint
        L.orc_filter_insert_seq.argtypes = [C.POINTER(Filter), cp, C.c_size_t]
        L.orc_filter_insert_seq.restype = None
        L.orc_filter_save.argtypes = [C.POINTER(Filter), cp]
        L.orc_filter_load.argtypes = [cp]
        L.orc_filter_load.restype = C.POINTER(Filter)
        L.orc_filter_fpr.argtypes = [C.POINTER(Filter)]
        L.orc_filter_fpr.restype = C.c_double
        L.orc_scan_counts.argtypes = [C.POINTER(Filter), cp, C.c_size_t, u8p]
        L.orc_scan_counts.restype = None
        L.orc_params_default.argtypes = [C.POINTER(Params), C.c_uint, C.c_uint]
        L.orc_params_default.restype = None
        L.orc_polish_contig.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(Filter), C.POINTER(Filter),
                                        C.POINTER(Params), C.POINTER(Result)]
        L.orc_result_free.argtypes = [C.POINTER(Result)]
        L.orc_result_free.restype = None
        L.orc_write_contig.argtypes = [cp, cp, C.c_uint32, C.POINTER(Result), C.POINTER(Params),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.orc_tsv_header.argtypes = [C.POINTER(Params), C.c_int, C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def default_params(k, h, **kw):
    p = Params()
    lib().orc_params_default(C.byref(p), k, h)
    for key, val in kw.items():
        if not hasattr(p, key):
            raise KeyError(key)
        setattr(p, key, val)
    return p


def nthash_kmer(kmer: bytes, h: int):
    """(fh, rh, [h hashes]) of one k-mer: NTMC64 seed form, ntedit.cpp:403-416."""
    fh, rh = C.c_uint64(), C.c_uint64()
    hv = (C.c_uint64 * h)()
    lib().orc_ntmc64_seed(kmer, len(kmer), h, C.byref(fh), C.byref(rh), hv)
    return fh.value, rh.value, list(hv)


class OracleFilter:
    def __init__(self, ptr):
        self.ptr = ptr

    @classmethod
    def new(cls, nbytes, k, h, counting=False):
        return cls(lib().orc_filter_new(nbytes, k, h, int(counting)))

    @classmethod
    def load(cls, path):
        p = lib().orc_filter_load(path.encode())
        if not p:
            raise IOError("cannot load filter " + path)
        return cls(p)

    @property
    def k(self):
        return self.ptr.contents.k

    @property
    def h(self):
        return self.ptr.contents.h

    @property
    def nbytes(self):
        return self.ptr.contents.bytes

    @property
    def counting(self):
        return bool(self.ptr.contents.counting)

    def insert_seq(self, seq: bytes):
        lib().orc_filter_insert_seq(self.ptr, seq, len(seq))

    def save(self, path):
        if lib().orc_filter_save(self.ptr, path.encode()) != 0:
            raise IOError("cannot save filter " + path)

    def fpr(self):
        return lib().orc_filter_fpr(self.ptr)

    def data(self):
        import numpy as np
        return np.ctypeslib.as_array(self.ptr.contents.data, shape=(self.nbytes,))

    def scan_counts(self, seq: bytes):
        import numpy as np
        out = np.empty(len(seq), dtype=np.uint8)
        lib().orc_scan_counts(self.ptr, seq, len(seq), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def free(self):
        if self.ptr:
            lib().orc_filter_free(self.ptr)
            self.ptr = None


def tsv_header(params, counting):
    buf = C.create_string_buffer(512)
    n = lib().orc_tsv_header(C.byref(params), int(counting), buf, 512)
    assert n > 0
    return buf.raw[:n]


def polish(contigs, bloom, params, bloomrep=None, min_contig_len=100):
    """Run the C restatement over [(header, seq bytes)], returning (edited_fa, changes_tsv, vcf_body) bytes
    laid out like the reference's output files (ntedit.cpp:2154-2259; VCF header lines omitted)."""
    L = lib()
    fa, tsv, vcf = [], [tsv_header(params, bloom.counting)], []
    for hdr, seq in contigs:
        if len(seq) < min_contig_len:
            continue
        buf = C.create_string_buffer(seq, len(seq))
        res = Result()
        rc = L.orc_polish_contig(buf, len(seq), bloom.ptr, bloomrep.ptr if bloomrep else None, C.byref(params),
                                 C.byref(res))
        assert rc == 0
        outs = [C.c_void_p() for _ in range(3)]
        lens = [C.c_size_t() for _ in range(3)]
        rc = L.orc_write_contig(hdr, buf, len(seq), C.byref(res), C.byref(params), C.byref(outs[0]),
                                C.byref(lens[0]), C.byref(outs[1]), C.byref(lens[1]), C.byref(outs[2]),
                                C.byref(lens[2]))
        assert rc == 0
        for dst, o, n in zip((fa, tsv, vcf), outs, lens):
            dst.append(C.string_at(o.value, n.value))
            _libc.free(o)
        L.orc_result_free(C.byref(res))
    return b"".join(fa), b"".join(tsv), b"".join(vcf)


def have_ref():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(draft_path, filter_path, workdir=None, threads=1, extra=(), rep_path=None, timeout=3600):
    """Run the unmodified reference binary; returns (edited_fa, changes_tsv, variants_vcf) bytes."""
    tmp = workdir or tempfile.mkdtemp(prefix="ntref_")
    prefix = os.path.join(tmp, "ref")
    cmd = [REF_BIN, "-f", draft_path, "-r", filter_path, "-b", prefix, "-t", str(threads)] + [str(x) for x in extra]
    if rep_path:
        cmd += ["-e", rep_path]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("ntedit_ref failed: %s\n%s" % (" ".join(cmd), r.stderr.decode(errors="replace")[-2000:]))
    outs = []
    for suffix in ("_edited.fa", "_changes.tsv", "_variants.vcf"):
        with open(prefix + suffix, "rb") as fh:
            outs.append(fh.read())
    return tuple(outs)
