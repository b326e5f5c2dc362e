"""ctypes binding of the CPU oracle (oracle/libac_oracle.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
_lib = C.CDLL(str(ROOT / "oracle" / "libac_oracle.so"))

STANDARD, LEFTMOST_FIRST, LEFTMOST_LONGEST = 0, 1, 2
START_UNANCHORED, START_ANCHORED, START_BOTH = 0, 1, 2
KIND_AUTO, KIND_NFA, KIND_CONTIGUOUS, KIND_DFA = 0, 1, 2, 3
PRE_NONE, PRE_MEMMEM, PRE_START_BYTES, PRE_RARE_BYTES, PRE_PACKED = range(5)
E_OVERFLOW = -21


class Opts(C.Structure):
    _fields_ = [("match_kind", C.c_int), ("start_kind", C.c_int),
                ("ascii_case_insensitive", C.c_int), ("byte_classes", C.c_int),
                ("prefilter", C.c_int), ("kind", C.c_int), ("dense_depth", C.c_int64)]


class Match(C.Structure):
    _fields_ = [("pid", C.c_uint32), ("_pad", C.c_uint32), ("start", C.c_uint64), ("end", C.c_uint64)]


MATCH_DTYPE = np.dtype([("pid", "<u4"), ("_pad", "<u4"), ("start", "<u8"), ("end", "<u8")])


class DfaView(C.Structure):
    _fields_ = [("trans", C.POINTER(C.c_uint32)), ("trans_len", C.c_uint64),
                ("stride2", C.c_uint32), ("alphabet_len", C.c_uint32),
                ("byte_classes", C.POINTER(C.c_uint8)),
                ("max_special_id", C.c_uint32), ("max_match_id", C.c_uint32),
                ("start_unanchored_id", C.c_uint32), ("start_anchored_id", C.c_uint32),
                ("match_offsets", C.POINTER(C.c_uint32)), ("match_pids", C.POINTER(C.c_uint32)),
                ("num_match_states", C.c_uint32),
                ("pattern_lens", C.POINTER(C.c_uint32)), ("n_patterns", C.c_uint32),
                ("match_kind", C.c_uint32),
                ("min_pattern_len", C.c_uint64), ("max_pattern_len", C.c_uint64),
                ("state_len", C.c_uint64)]


class PackedConfig(C.Structure):
    _fields_ = [("kind", C.c_int), ("force", C.c_int), ("only_teddy_fat", C.c_int),
                ("only_teddy_256bit", C.c_int), ("heuristic_pattern_limits", C.c_int)]


_lib.orc_build.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t,
                           C.POINTER(Opts), C.POINTER(C.c_void_p)]
_lib.orc_free.argtypes = [C.c_void_p]
for f in ("orc_kind", "orc_match_kind", "orc_start_kind", "orc_prefilter_kind"):
    getattr(_lib, f).argtypes = [C.c_void_p]
for f in ("orc_patterns_len", "orc_min_pattern_len", "orc_max_pattern_len"):
    getattr(_lib, f).argtypes = [C.c_void_p]
    getattr(_lib, f).restype = C.c_size_t
_lib.orc_packed_variant.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
_lib.orc_dfa_get.argtypes = [C.c_void_p, C.POINTER(DfaView)]
_lib.orc_try_find.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                              C.c_int, C.POINTER(Match), C.POINTER(C.c_int)]
_lib.orc_find_iter.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                               C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
_lib.orc_find_overlapping_iter.argtypes = _lib.orc_find_iter.argtypes
_lib.orc_scan_overlapping_count.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
_lib.orc_packed_build.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t,
                                  C.POINTER(PackedConfig), C.POINTER(C.c_void_p)]
_lib.orc_packed_free.argtypes = [C.c_void_p]
_lib.orc_packed_minimum_len.argtypes = [C.c_void_p]
_lib.orc_packed_minimum_len.restype = C.c_size_t
_lib.orc_packed_find_iter.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_size_t)]


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(f"oracle error {code}")
        self.code = code


def _pack_patterns(patterns):
    pats = [bytes(p) for p in patterns]
    n = len(pats)
    arr = (C.c_char_p * max(n, 1))()
    bufs = []
    for i, p in enumerate(pats):
        b = C.create_string_buffer(p, max(len(p), 1))
        bufs.append(b)
        arr[i] = C.cast(b, C.c_char_p)
    lens = (C.c_size_t * max(n, 1))(*[len(p) for p in pats])
    return arr, lens, n, bufs


def _hay_ptr(hay):
    """Return (keepalive, void*, len) for bytes or a contiguous uint8 numpy array."""
    if isinstance(hay, np.ndarray):
        assert hay.dtype == np.uint8 and hay.flags["C_CONTIGUOUS"]
        return hay, hay.ctypes.data, hay.size
    b = bytes(hay)
    buf = C.create_string_buffer(b, max(len(b), 1))
    return buf, C.addressof(buf), len(b)


class Oracle:
    def __init__(self, patterns, match_kind=STANDARD, start_kind=START_UNANCHORED,
                 ascii_case_insensitive=False, byte_classes=True, prefilter=True,
                 kind=KIND_AUTO, dense_depth=3):
        o = Opts(match_kind, start_kind, int(ascii_case_insensitive), int(byte_classes),
                 int(prefilter), kind, dense_depth)
        arr, lens, n, bufs = _pack_patterns(patterns)
        h = C.c_void_p()
        rc = _lib.orc_build(arr, lens, n, C.byref(o), C.byref(h))
        if rc:
            raise OracleError(rc)
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.orc_free(self._h)
            self._h = None

    kind = property(lambda s: _lib.orc_kind(s._h))
    match_kind = property(lambda s: _lib.orc_match_kind(s._h))
    prefilter_kind = property(lambda s: _lib.orc_prefilter_kind(s._h))
    patterns_len = property(lambda s: _lib.orc_patterns_len(s._h))
    min_pattern_len = property(lambda s: _lib.orc_min_pattern_len(s._h))
    max_pattern_len = property(lambda s: _lib.orc_max_pattern_len(s._h))

    def packed_variant(self):
        fat, ml, vb = C.c_int(), C.c_int(), C.c_int()
        if not _lib.orc_packed_variant(self._h, C.byref(fat), C.byref(ml), C.byref(vb)):
            return None
        return {"fat": bool(fat.value), "mask_len": ml.value, "vector_bytes": vb.value}

    def dfa(self):
        v = DfaView()
        rc = _lib.orc_dfa_get(self._h, C.byref(v))
        if rc:
            raise OracleError(rc)
        # meaningful CSR prefix: rows 2 ..= max_match_id >> stride2 (StartKind::Both over-allocates
        # `matches` when the start states are match states, src/dfa.rs:479-491)
        nms = min(v.num_match_states, (v.max_match_id >> v.stride2) - 1)
        offs = np.ctypeslib.as_array(v.match_offsets, (nms + 1,)).copy()
        return {
            "trans": np.ctypeslib.as_array(v.trans, (v.trans_len,)).copy() if v.trans_len else np.zeros(0, np.uint32),
            "stride2": v.stride2, "alphabet_len": v.alphabet_len,
            "byte_classes": np.ctypeslib.as_array(v.byte_classes, (256,)).copy(),
            "max_special_id": v.max_special_id, "max_match_id": v.max_match_id,
            "start_unanchored_id": v.start_unanchored_id, "start_anchored_id": v.start_anchored_id,
            "match_offsets": offs,
            "match_pids": (np.ctypeslib.as_array(v.match_pids, (int(offs[-1]),)).copy()
                           if int(offs[-1]) else np.zeros(0, np.uint32)),
            "pattern_lens": (np.ctypeslib.as_array(v.pattern_lens, (v.n_patterns,)).copy()
                             if v.n_patterns else np.zeros(0, np.uint32)),
            "match_kind": v.match_kind, "min_pattern_len": v.min_pattern_len,
            "max_pattern_len": v.max_pattern_len, "state_len": v.state_len,
        }

    def try_find(self, hay, span=None, anchored=False, earliest=False):
        keep, ptr, n = _hay_ptr(hay)
        s, e = span if span is not None else (0, n)
        m, found = Match(), C.c_int()
        rc = _lib.orc_try_find(self._h, ptr, n, s, e, int(anchored), int(earliest), C.byref(m), C.byref(found))
        if rc:
            raise OracleError(rc)
        return (m.pid, m.start, m.end) if found.value else None

    def _iter(self, fn, hay, span, anchored):
        keep, ptr, n = _hay_ptr(hay)
        s, e = span if span is not None else (0, n)
        cap = 1024
        while True:
            out = np.zeros(cap, MATCH_DTYPE)
            cnt = C.c_size_t()
            rc = fn(self._h, ptr, n, s, e, int(anchored), out.ctypes.data, cap, C.byref(cnt))
            if rc == E_OVERFLOW:
                cap = cnt.value
                continue
            if rc:
                raise OracleError(rc)
            return out[: cnt.value]

    def find_iter_np(self, hay, span=None, anchored=False):
        return self._iter(_lib.orc_find_iter, hay, span, anchored)

    def find_overlapping_iter_np(self, hay, span=None, anchored=False):
        return self._iter(_lib.orc_find_overlapping_iter, hay, span, anchored)

    def find_iter(self, hay, span=None, anchored=False):
        r = self.find_iter_np(hay, span, anchored)
        return [(int(a), int(b), int(c)) for a, b, c in zip(r["pid"], r["start"], r["end"])]

    def find_overlapping_iter(self, hay, span=None, anchored=False):
        r = self.find_overlapping_iter_np(hay, span, anchored)
        return [(int(a), int(b), int(c)) for a, b, c in zip(r["pid"], r["start"], r["end"])]

    def scan_overlapping_count(self, hay, span=None):
        keep, ptr, n = _hay_ptr(hay)
        s, e = span if span is not None else (0, n)
        cnt, fnv = C.c_uint64(), C.c_uint64()
        rc = _lib.orc_scan_overlapping_count(self._h, ptr, n, s, e, C.byref(cnt), C.byref(fnv))
        if rc:
            raise OracleError(rc)
        return cnt.value, fnv.value


class PackedOracle:
    """packed::Searcher (src/packed/api.rs). `.built` is False when the reference returns None."""

    def __init__(self, patterns, kind=0, force=0, only_teddy_fat=-1, only_teddy_256bit=-1,
                 heuristic_pattern_limits=True):
        cfg = PackedConfig(kind, force, only_teddy_fat, only_teddy_256bit, int(heuristic_pattern_limits))
        arr, lens, n, bufs = _pack_patterns(patterns)
        h = C.c_void_p()
        rc = _lib.orc_packed_build(arr, lens, n, C.byref(cfg), C.byref(h))
        if rc:
            raise OracleError(rc)
        self._h = h
        self.built = bool(h.value)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.orc_packed_free(self._h)
            self._h = None

    @property
    def minimum_len(self):
        return _lib.orc_packed_minimum_len(self._h)

    def find_iter(self, hay):
        keep, ptr, n = _hay_ptr(hay)
        cap = 1024
        while True:
            out = np.zeros(cap, MATCH_DTYPE)
            cnt = C.c_size_t()
            rc = _lib.orc_packed_find_iter(self._h, ptr, n, out.ctypes.data, cap, C.byref(cnt))
            if rc == E_OVERFLOW:
                cap = cnt.value
                continue
            if rc:
                raise OracleError(rc)
            r = out[: cnt.value]
            return [(int(a), int(b), int(c)) for a, b, c in zip(r["pid"], r["start"], r["end"])]
