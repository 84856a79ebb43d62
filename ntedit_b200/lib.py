"""Build + ctypes binding of libntedit_b200.so (C ABI: include/ntedit_b200.h).

The library is built in-tree (ntedit_b200/_lib/) with nvcc for sm_100a only.  There is no CPU fallback: if the
library is missing or no CUDA device is present, every compute call raises.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
SO = os.path.join(LIBDIR, "libntedit_b200.so")
CLI = os.path.join(LIBDIR, "ntedit-b200")
MAKE_BF = os.path.join(LIBDIR, "ntedit-b200-make-bf")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]


def _sources():
    return [os.path.join(CSRC, f) for f in ("kernels.cu", "capi.cu")]


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(INCLUDE, "ntedit_b200.h"))
    return deps


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile the CUDA library (and the ntedit-b200 command line tool) for sm_100a."""
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    out = None if verbose else subprocess.DEVNULL
    if force or _stale(SO, _deps()):
        cmd = [nvcc] + NVCC_FLAGS + ["-Xcompiler", "-fPIC", "-shared"] + _sources() + ["-o", SO]
        subprocess.run(cmd, check=True, stdout=out)
    for src, exe in (("cli.cpp", CLI), ("make_bf_cli.cpp", MAKE_BF)):
        path = os.path.join(CSRC, src)
        if os.path.exists(path) and (force or _stale(exe, _deps() + [SO])):
            cmd = ["g++", "-O2", "-std=c++17", "-I", INCLUDE, path, "-o", exe, "-L", LIBDIR, "-lntedit_b200",
                   "-Wl,-rpath,$ORIGIN", "-lz", "-pthread"]
            subprocess.run(cmd, check=True, stdout=out)
    return SO


class FilterInfo(C.Structure):
    _fields_ = [("bytes", C.c_uint64), ("k", C.c_uint32), ("hash_num", C.c_uint32), ("counting", C.c_int32),
                ("device", C.c_int32), ("fpr", C.c_double)]


class Params(C.Structure):
    _fields_ = [("jump", C.c_uint32), ("mode", C.c_int32), ("snv", C.c_int32), ("mask", C.c_int32),
                ("max_insertions", C.c_uint32), ("max_deletions", C.c_uint32), ("edit_threshold", C.c_float),
                ("missing_threshold", C.c_float), ("edit_ratio", C.c_float), ("missing_ratio", C.c_float),
                ("use_ratio", C.c_int32), ("min_threshold", C.c_uint32), ("max_threshold", C.c_uint32),
                ("min_contig_len", C.c_uint32), ("segment_len", C.c_uint32)]


class Node(C.Structure):
    _fields_ = [("node_type", C.c_int32), ("s_pos", C.c_uint32), ("e_pos", C.c_uint32), ("num_support", C.c_uint32),
                ("c", C.c_uint8), ("pad_", C.c_uint8 * 3)]


class SRec(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("num_support", C.c_uint32), ("altsupp1", C.c_uint32), ("altsupp2", C.c_uint32),
                ("altsupp3", C.c_uint32), ("draft_char", C.c_uint8), ("sub_base", C.c_uint8), ("altbase1", C.c_uint8),
                ("altbase2", C.c_uint8), ("altbase3", C.c_uint8), ("pad_", C.c_uint8 * 3)]


class Stats(C.Structure):
    _fields_ = [("bases", C.c_uint64), ("contigs", C.c_uint64), ("sites", C.c_uint64), ("edits", C.c_uint64),
                ("segments", C.c_uint64), ("reruns", C.c_uint64), ("rounds", C.c_uint32),
                ("kernel_launches", C.c_uint32), ("ms_scan", C.c_float), ("ms_walk", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("ms_host", C.c_float), ("ms_pre", C.c_float),
                ("pad_", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class StrBuf(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t)]


# every symbol include/ntedit_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_U64P = C.POINTER(C.c_uint64)
SYMBOLS = {
    "ntb_last_error": (C.c_char_p, []),
    "ntb_version": (C.c_char_p, []),
    "ntb_device_count": (C.c_int, []),
    "ntb_host_alloc": (_VP, [C.c_size_t]),
    "ntb_host_free": (None, [_VP]),
    "ntb_filter_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_VP)]),
    "ntb_filter_create": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(_VP)]),
    "ntb_filter_wrap_device": (C.c_int, [_VP, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(_VP)]),
    "ntb_filter_replicate": (C.c_int, [_VP, C.c_int, C.POINTER(_VP)]),
    "ntb_filter_get_info": (C.c_int, [_VP, C.POINTER(FilterInfo)]),
    "ntb_filter_device_ptr": (_VP, [_VP]),
    "ntb_filter_insert": (C.c_int, [_VP, _VP, _U64P, C.c_uint64]),
    "ntb_filter_insert_batch": (C.c_int, [_VP, _VP]),
    "ntb_filter_save": (C.c_int, [_VP, C.c_char_p]),
    "ntb_filter_download": (C.c_int, [_VP, _VP, C.c_uint64]),
    "ntb_filter_free": (None, [_VP]),
    "ntb_params_init": (None, [C.POINTER(Params)]),
    "ntb_batch_upload": (C.c_int, [_VP, _U64P, C.c_uint64, C.c_int, C.POINTER(_VP)]),
    "ntb_batch_wrap_device": (C.c_int, [_VP, _U64P, C.c_uint64, C.c_int, C.POINTER(_VP)]),
    "ntb_batch_total_bases": (C.c_uint64, [_VP]),
    "ntb_batch_free": (None, [_VP]),
    "ntb_scan": (C.c_int, [_VP, _VP, _U64P, C.c_uint64, _VP, _VP]),
    "ntb_polish_batch": (C.c_int, [_VP, _VP, C.POINTER(Params), _VP, _U64P, C.c_uint64, C.POINTER(_VP)]),
    "ntb_polish_device": (C.c_int, [_VP, _VP, C.POINTER(Params), _VP, _VP, C.POINTER(_VP)]),
    "ntb_result_contig": (C.c_int, [_VP, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.POINTER(Node)), _U64P,
                                    C.POINTER(C.POINTER(SRec)), _U64P]),
    "ntb_result_stats": (C.c_int, [_VP, C.POINTER(Stats)]),
    "ntb_result_free": (None, [_VP]),
    "ntb_format_contig": (C.c_int, [C.c_char_p, _VP, C.POINTER(Node), C.c_uint64, C.POINTER(SRec), C.c_uint64, C.c_int,
                                    C.POINTER(StrBuf), C.POINTER(StrBuf), C.POINTER(StrBuf)]),
    "ntb_format_tsv_header": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(StrBuf)]),
    "ntb_strbuf_free": (None, [C.POINTER(StrBuf)]),
}

_lib = None


class NtbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ntedit_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        so = os.environ.get("NTB_LIB", SO)  # tuning aid: load a differently compiled build of the same library
        if not os.path.exists(so):
            raise ImportError("libntedit_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(ntedit_b200 has no CPU fallback)")
        L = C.CDLL(so)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise NtbError(rc, load().ntb_last_error().decode(errors="replace"))
