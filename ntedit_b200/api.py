"""Host-side mirror of ntEdit's interface for the hot path, on top of the C ABI.

Names follow the reference: a `BloomFilter` stands where `BFWrapper` does (ntedit.cpp:350-401: get_k, get_hash_num,
is_counting, print_details), `kmerize_and_correct` replaces the per-contig `kmerizeAndCorrect` calls of
`readAndCorrect` (ntedit.cpp:2154-2259) for a whole batch of contigs, and `write_edits` is `writeEditsToFile`
(ntedit.cpp:925-1213).  All compute runs in the CUDA library; nothing here computes a hash or probes a filter.
"""
import ctypes as C

import numpy as np

from . import lib as _l


def default_params(**kw):
    """ntb_params with the defaults of `namespace opt` (ntedit.cpp:99-133); keyword overrides by field name."""
    p = _l.Params()
    _l.load().ntb_params_init(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def pack_contigs(contigs):
    """[(header, seq bytes)] -> (uint8 array of NUL-terminated sequences laid end to end, uint64 offsets)."""
    offs = np.zeros(len(contigs) + 1, dtype=np.uint64)
    total = 0
    for i, (_, s) in enumerate(contigs):
        total += len(s) + 1
        offs[i + 1] = total
    buf = np.zeros(total, dtype=np.uint8)
    for i, (_, s) in enumerate(contigs):
        o = int(offs[i])
        buf[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
    return buf, offs


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


class BloomFilter:
    """Device-resident Bloom / counting Bloom filter (BFWrapper, ntedit.cpp:350-401)."""

    def __init__(self, handle):
        self._h = handle
        self._info = None

    @classmethod
    def load(cls, path, device=0):
        h = C.c_void_p()
        _l.check(_l.load().ntb_filter_load(path.encode(), device, C.byref(h)))
        return cls(h)

    @classmethod
    def create(cls, nbytes, k, hash_num, counting=False, device=0):
        h = C.c_void_p()
        _l.check(_l.load().ntb_filter_create(nbytes, k, hash_num, int(counting), device, C.byref(h)))
        return cls(h)

    @classmethod
    def wrap_device(cls, dev_ptr, nbytes, k, hash_num, counting=False, device=0):
        h = C.c_void_p()
        _l.check(_l.load().ntb_filter_wrap_device(C.c_void_p(dev_ptr), nbytes, k, hash_num, int(counting), device,
                                                  C.byref(h)))
        return cls(h)

    def replicate(self, device):
        """A copy of this filter on another device (one device-to-device copy)."""
        h = C.c_void_p()
        _l.check(_l.load().ntb_filter_replicate(self._h, device, C.byref(h)))
        return BloomFilter(h)

    def info(self, refresh=False):
        if self._info is None or refresh:
            fi = _l.FilterInfo()
            _l.check(_l.load().ntb_filter_get_info(self._h, C.byref(fi)))
            self._info = fi
        return self._info

    def get_k(self):
        return self.info().k

    def get_hash_num(self):
        return self.info().hash_num

    def is_counting(self):
        return bool(self.info().counting)

    def get_bytes(self):
        return self.info().bytes

    def get_fpr(self):
        return self.info(refresh=True).fpr

    def details(self):
        """The BLOOM:: line of print_details (ntedit.cpp:387-395)."""
        fi = self.info(refresh=True)
        return "BLOOM::\tcounting: %s\tsize: %d\tnumber hash functions: %d\tkmer size: %d\tFPR: %g" % (
            "YES" if fi.counting else "NO", fi.bytes, fi.hash_num, fi.k, fi.fpr)

    def device_ptr(self):
        return _l.load().ntb_filter_device_ptr(self._h)

    def insert(self, contigs):
        buf, offs = pack_contigs(contigs)
        _l.check(_l.load().ntb_filter_insert(self._h, buf.ctypes.data_as(C.c_void_p), _u64p(offs), len(contigs)))
        self._info = None

    def insert_batch(self, batch):
        _l.check(_l.load().ntb_filter_insert_batch(self._h, batch._h))
        self._info = None

    def save(self, path):
        _l.check(_l.load().ntb_filter_save(self._h, path.encode()))

    def download(self):
        n = self.get_bytes()
        out = np.empty(n, dtype=np.uint8)
        _l.check(_l.load().ntb_filter_download(self._h, out.ctypes.data_as(C.c_void_p), n))
        return out

    def free(self):
        if self._h:
            _l.load().ntb_filter_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Batch:
    """A batch of contigs resident on the device."""

    def __init__(self, handle, offsets, keep=None):
        self._h = handle
        self.offsets = offsets
        self._keep = keep

    @classmethod
    def upload(cls, buf, offs, device=0):
        h = C.c_void_p()
        _l.check(_l.load().ntb_batch_upload(buf.ctypes.data_as(C.c_void_p), _u64p(offs), len(offs) - 1, device,
                                            C.byref(h)))
        return cls(h, offs)

    @classmethod
    def upload_ptr(cls, host_ptr, offs, device=0):
        h = C.c_void_p()
        _l.check(_l.load().ntb_batch_upload(C.c_void_p(host_ptr), _u64p(offs), len(offs) - 1, device, C.byref(h)))
        return cls(h, offs)

    @classmethod
    def wrap_device(cls, dev_ptr, offs, device=0):
        h = C.c_void_p()
        _l.check(_l.load().ntb_batch_wrap_device(C.c_void_p(dev_ptr), _u64p(offs), len(offs) - 1, device, C.byref(h)))
        return cls(h, offs)

    def total_bases(self):
        return _l.load().ntb_batch_total_bases(self._h)

    def free(self):
        if self._h:
            _l.load().ntb_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PolishResult:
    def __init__(self, handle, n_contigs):
        self._h = handle
        self.n_contigs = n_contigs

    def stats(self):
        st = _l.Stats()
        _l.check(_l.load().ntb_result_stats(self._h, C.byref(st)))
        return st

    def contig(self, c):
        """(polished, nodes pointer, n_nodes, srecs pointer, n_srecs) of contig c."""
        pol = C.c_int()
        nodes = C.POINTER(_l.Node)()
        srecs = C.POINTER(_l.SRec)()
        nn = C.c_uint64()
        ns = C.c_uint64()
        _l.check(_l.load().ntb_result_contig(self._h, c, C.byref(pol), C.byref(nodes), C.byref(nn), C.byref(srecs),
                                             C.byref(ns)))
        return bool(pol.value), nodes, nn.value, srecs, ns.value

    def free(self):
        if self._h:
            _l.load().ntb_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def scan(bloom, contigs):
    """K1 through the ABI: (counts uint8 per buffer position, valid bool per buffer position, offsets)."""
    buf, offs = pack_contigs(contigs)
    n = len(buf)
    counts = np.zeros(n, dtype=np.uint8)
    valid = np.zeros((n + 31) // 32, dtype=np.uint32)
    _l.check(_l.load().ntb_scan(bloom._h, buf.ctypes.data_as(C.c_void_p), _u64p(offs), len(contigs),
                                counts.ctypes.data_as(C.c_void_p), valid.ctypes.data_as(C.c_void_p)))
    vb = np.unpackbits(valid.view(np.uint8), bitorder="little")[:n].astype(bool)
    return counts, vb, offs


def kmerize_and_correct(buf, offs, bloom, params, bloomrep=None):
    """Polish a packed batch held in HOST memory (`buf` is mutated in place like contigSeq); returns a PolishResult."""
    res = C.c_void_p()
    _l.check(_l.load().ntb_polish_batch(bloom._h, bloomrep._h if bloomrep else None, C.byref(params),
                                        buf.ctypes.data_as(C.c_void_p), _u64p(offs), len(offs) - 1, C.byref(res)))
    return PolishResult(res, len(offs) - 1)


def kmerize_and_correct_device(batch, bloom, params, bloomrep=None, host_buf=None):
    """Polish a batch already resident on the device."""
    res = C.c_void_p()
    hb = host_buf.ctypes.data_as(C.c_void_p) if host_buf is not None else None
    _l.check(_l.load().ntb_polish_device(bloom._h, bloomrep._h if bloomrep else None, C.byref(params), batch._h, hb,
                                         C.byref(res)))
    return PolishResult(res, len(batch.offsets) - 1)


def write_edits(headers, buf, offs, result, params, k, counting):
    """writeEditsToFile for every polished contig: returns (edited_fa, changes_tsv incl. header, vcf rows) bytes."""
    L = _l.load()
    fa, tsv, vcf = _l.StrBuf(), _l.StrBuf(), _l.StrBuf()
    _l.check(L.ntb_format_tsv_header(k, params.jump, int(counting), C.byref(tsv)))
    base = buf.ctypes.data
    for c, hdr in enumerate(headers):
        pol, nodes, nn, srecs, ns = result.contig(c)
        if not pol:
            continue
        _l.check(L.ntb_format_contig(hdr, C.c_void_p(base + int(offs[c])), nodes, nn, srecs, ns, int(params.snv),
                                     C.byref(fa), C.byref(tsv), C.byref(vcf)))
    out = []
    for b in (fa, tsv, vcf):
        out.append(C.string_at(b.data, b.len) if b.len else b"")
        L.ntb_strbuf_free(C.byref(b))
    return tuple(out)


def polish_per_contig(contigs, bloom, params, bloomrep=None):
    """[(header, seq)] -> one (edited_fa, tsv_rows, vcf_rows) per contig, None for contigs below min_contig_len.
    The unit the multi-GPU driver (shard.py) merges in input order."""
    buf, offs = pack_contigs(contigs)
    res = kmerize_and_correct(buf, offs, bloom, params, bloomrep)
    L = _l.load()
    base = buf.ctypes.data
    out = []
    for c, (hdr, _) in enumerate(contigs):
        pol, nodes, nn, srecs, ns = res.contig(c)
        if not pol:
            out.append(None)
            continue
        fa, tsv, vcf = _l.StrBuf(), _l.StrBuf(), _l.StrBuf()
        _l.check(L.ntb_format_contig(hdr, C.c_void_p(base + int(offs[c])), nodes, nn, srecs, ns, int(params.snv),
                                     C.byref(fa), C.byref(tsv), C.byref(vcf)))
        piece = []
        for b in (fa, tsv, vcf):
            piece.append(C.string_at(b.data, b.len) if b.len else b"")
            L.ntb_strbuf_free(C.byref(b))
        out.append(tuple(piece))
    res.free()
    return out


def polish(contigs, bloom, params, bloomrep=None):
    """Convenience: [(header, seq)] -> (edited_fa, changes_tsv, vcf_rows, stats dict)."""
    buf, offs = pack_contigs(contigs)
    res = kmerize_and_correct(buf, offs, bloom, params, bloomrep)
    fi = bloom.info()
    outs = write_edits([h for h, _ in contigs], buf, offs, res, params, fi.k, fi.counting)
    st = res.stats().as_dict()
    res.free()
    return outs + (st,)
