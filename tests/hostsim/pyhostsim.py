"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU simulator of the device engine (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libhostsim.so")


class NtbParams(C.Structure):
    _fields_ = [("jump", C.c_uint32), ("mode", C.c_int32), ("snv", C.c_int32), ("mask", C.c_int32),
                ("max_insertions", C.c_uint32), ("max_deletions", C.c_uint32), ("edit_threshold", C.c_float),
                ("missing_threshold", C.c_float), ("edit_ratio", C.c_float), ("missing_ratio", C.c_float),
                ("use_ratio", C.c_int32), ("min_threshold", C.c_uint32), ("max_threshold", C.c_uint32),
                ("min_contig_len", C.c_uint32), ("segment_len", C.c_uint32)]


class NtbStats(C.Structure):
    _fields_ = [("bases", C.c_uint64), ("contigs", C.c_uint64), ("sites", C.c_uint64), ("edits", C.c_uint64),
                ("segments", C.c_uint64), ("reruns", C.c_uint64), ("rounds", C.c_uint32),
                ("kernel_launches", C.c_uint32), ("ms_scan", C.c_float), ("ms_walk", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("ms_host", C.c_float), ("ms_pre", C.c_float),
                ("pad_", C.c_uint32)]


def default_params(**kw):
    p = NtbParams(jump=3, mode=0, snv=0, mask=0, max_insertions=5, max_deletions=5, edit_threshold=9.0,
                  missing_threshold=5.0, edit_ratio=0.5, missing_ratio=0.5, use_ratio=0, min_threshold=1,
                  max_threshold=255, min_contig_len=100, segment_len=0)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def build():
    srcs = [os.path.join(HERE, "hostsim.cpp")]
    deps = srcs + [os.path.join(ROOT, "ntedit_b200", "csrc", f) for f in
                   ("engine.h", "site_dense.h", "nthash.h", "ntb_common.h", "replay.hpp", "polish_driver.hpp", "writer.hpp")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", SO] + srcs, check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
    return _lib


def pack_batch(contigs):
    """[(hdr, seq bytes)] -> (buffer bytes, offsets list): NUL-terminated sequences laid end to end."""
    offs = [0]
    parts = []
    for _, s in contigs:
        parts.append(s + b"\0")
        offs.append(offs[-1] + len(s) + 1)
    return b"".join(parts), offs


def polish(contigs, filt_bytes, k, h, counting, params, rep=None):
    """rep = (bytes, h, counting) or None.  Returns (fa, tsv, vcf_body, stats)."""
    L = lib()
    buf, offs = pack_batch(contigs)
    cbuf = C.create_string_buffer(buf, len(buf))
    coffs = (C.c_uint64 * len(offs))(*offs)
    hdrs = (C.c_char_p * len(contigs))(*[h_ for h_, _ in contigs])
    outs = [C.c_void_p() for _ in range(3)]
    lens = [C.c_size_t() for _ in range(3)]
    st = NtbStats()
    err = C.create_string_buffer(512)
    fb = (C.c_uint8 * len(filt_bytes)).from_buffer_copy(filt_bytes)
    if rep:
        rb = (C.c_uint8 * len(rep[0])).from_buffer_copy(rep[0])
        rargs = (rb, len(rep[0]), rep[1], int(rep[2]))
    else:
        rargs = (None, 0, 0, 0)
    L.hostsim_polish.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64,
                                 C.c_uint32, C.c_int, C.POINTER(NtbParams), C.c_void_p, C.POINTER(C.c_uint64),
                                 C.c_uint64, C.POINTER(C.c_char_p)] + [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)] * 3 + \
                                [C.POINTER(NtbStats), C.c_char_p, C.c_size_t]
    rc = L.hostsim_polish(C.cast(fb, C.c_void_p), len(filt_bytes), k, h, int(counting),
                          C.cast(rargs[0], C.c_void_p) if rep else None, rargs[1], rargs[2], rargs[3],
                          C.byref(params), C.cast(cbuf, C.c_void_p), coffs, len(contigs), hdrs,
                          C.byref(outs[0]), C.byref(lens[0]), C.byref(outs[1]), C.byref(lens[1]),
                          C.byref(outs[2]), C.byref(lens[2]), C.byref(st), err, 512)
    if rc != 0:
        raise RuntimeError("hostsim_polish rc=%d: %s" % (rc, err.value.decode()))
    res = []
    L.hostsim_free.argtypes = [C.c_void_p]
    for o, n in zip(outs, lens):
        res.append(C.string_at(o.value, n.value))
        L.hostsim_free(o)
    return res[0], res[1], res[2], st
