"""ctypes binding of the CPU oracle (oracle/refcpu.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "librefcpu.so")


class RcError(C.Structure):
    _fields_ = [("code", C.c_int32), ("a", C.c_uint64), ("b", C.c_uint64)]


class RcHeader(C.Structure):
    _fields_ = [("content_checksum_flag", C.c_uint8), ("window_size", C.c_uint64),
                ("has_dict_id", C.c_uint8), ("dictionnary_id", C.c_uint64),
                ("has_content_size", C.c_uint8), ("content_size", C.c_uint64)]


class RcFrameInfo(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("magic", C.c_uint32), ("src_off", C.c_uint64), ("src_len", C.c_uint64),
                ("out_off", C.c_uint64), ("out_len", C.c_uint64), ("n_blocks", C.c_uint32),
                ("has_checksum", C.c_uint8), ("stored_checksum", C.c_uint32), ("computed_xxh64_low32", C.c_uint32),
                ("header", RcHeader)]


# error codes (refcpu.h)
OK = 0
NotEnoughBytes, NotEnoughBits, MaximumReadableBitsExceeded, EmptyInputData, NullByte, EmptySliceError = 1, 2, 3, 4, 5, 6
LargeAccuracyLog, CorruptedTable, SequenceCodeMaxValueExceeded = 10, 11, 12
HuffmanDecoderMissing, CorruptedStreamsSizeTooBig = 20, 21
SeqReservedSet, NoPreviousDecoder = 30, 31
WindowSizeTooBig, NullOffsetError, ImpossibleValue = 40, 41, 42
ReservedBlockType = 50
UnrecognizedMagic, FrameReservedSet, MissingChecksum = 60, 61, 62
Panic = 99


class RefError(Exception):
    def __init__(self, err):
        self.code, self.a, self.b = err.code, err.a, err.b
        super().__init__(f"{lib().rc_strerror(err.code).decode()} (a={err.a}, b={err.b})")


def build(force=False):
    src = os.path.join(ROOT, "oracle", "refcpu.c")
    hdr = os.path.join(ROOT, "oracle", "refcpu.h")
    if (not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    u8p, sz, szp, ep = C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(RcError)
    L.rc_fwd_new.restype = C.c_void_p; L.rc_fwd_new.argtypes = [C.c_char_p, sz, ep]
    L.rc_fwd_free.argtypes = [C.c_void_p]
    L.rc_fwd_take.restype = C.c_uint64; L.rc_fwd_take.argtypes = [C.c_void_p, sz, ep]
    L.rc_fwd_peek.restype = C.c_uint64; L.rc_fwd_peek.argtypes = [C.c_void_p, sz, ep]
    L.rc_fwd_len.restype = sz; L.rc_fwd_len.argtypes = [C.c_void_p]
    L.rc_fwd_bytes_read.restype = sz; L.rc_fwd_bytes_read.argtypes = [C.c_void_p]
    L.rc_bwd_new.restype = C.c_void_p; L.rc_bwd_new.argtypes = [C.c_char_p, sz, ep]
    L.rc_bwd_free.argtypes = [C.c_void_p]
    L.rc_bwd_take.restype = C.c_uint64; L.rc_bwd_take.argtypes = [C.c_void_p, sz, ep]
    L.rc_bwd_len.restype = sz; L.rc_bwd_len.argtypes = [C.c_void_p]
    L.rc_parse_fse_table.argtypes = [C.c_char_p, sz, u8p, C.POINTER(C.c_int16), szp, szp, szp, ep]
    L.rc_fse_from_distribution.argtypes = [C.c_uint8, C.POINTER(C.c_int16), sz, C.POINTER(C.c_uint16), ep]
    L.rc_fse_run.argtypes = [C.POINTER(C.c_uint16), C.c_uint8, C.c_int, C.c_char_p, sz, sz, C.POINTER(C.c_uint16), szp, szp, ep]
    L.rc_huffman_from_weights.argtypes = [C.c_char_p, sz, u8p, C.POINTER(C.c_uint32), ep]
    L.rc_huffman_parse.argtypes = [C.c_char_p, sz, u8p, C.POINTER(C.c_uint32), szp, u8p, szp, ep]
    L.rc_huffman_decode_stream.argtypes = [u8p, C.POINTER(C.c_uint32), C.c_char_p, sz, u8p, sz, szp, ep]
    L.rc_execute_sequences.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), sz, C.c_char_p, sz, C.POINTER(C.c_void_p), szp, ep]
    L.rc_header_parse.argtypes = [C.c_char_p, sz, C.POINTER(RcHeader), szp, ep]
    L.rc_window_descriptor.restype = C.c_uint64; L.rc_window_descriptor.argtypes = [C.c_uint8]
    L.rc_decode_frames.argtypes = [C.c_char_p, sz, C.c_int, C.POINTER(C.c_void_p), szp, C.POINTER(C.c_void_p), szp, ep]
    L.rc_main_decode.argtypes = [C.c_char_p, sz, C.c_int, C.POINTER(C.c_void_p), szp, ep]
    L.rc_main_decode_mt.argtypes = [C.c_char_p, sz, C.c_int, C.c_int, C.POINTER(C.c_void_p), szp, ep]
    L.rc_xxh64.restype = C.c_uint64; L.rc_xxh64.argtypes = [C.c_char_p, sz, C.c_uint64]
    L.rc_free.argtypes = [C.c_void_p]
    L.rc_strerror.restype = C.c_char_p; L.rc_strerror.argtypes = [C.c_int]
    _lib = L
    return L


def _take_buf(ptr, n):
    data = C.string_at(ptr.value, n) if (ptr.value and n) else b""
    if ptr.value:
        lib().rc_free(ptr)
    return data


class FwdBits:
    """ForwardBitParser (parsing.rs:114-189)."""
    def __init__(self, data):
        self._d = bytes(data); e = RcError()
        self._h = lib().rc_fwd_new(self._d, len(self._d), C.byref(e))
        if not self._h:
            raise RefError(e)
    def take(self, n):
        e = RcError(); v = lib().rc_fwd_take(self._h, n, C.byref(e))
        if e.code: raise RefError(e)
        return v
    def peek(self, n):
        e = RcError(); v = lib().rc_fwd_peek(self._h, n, C.byref(e))
        if e.code: raise RefError(e)
        return v
    def len(self): return lib().rc_fwd_len(self._h)
    def is_empty(self): return self.len() == 0
    def bytes_read(self): return lib().rc_fwd_bytes_read(self._h)
    def __del__(self):
        if getattr(self, "_h", None): lib().rc_fwd_free(self._h)


class BwdBits:
    """BackwardBitParser (parsing.rs:191-259)."""
    def __init__(self, data):
        self._d = bytes(data); e = RcError()
        self._h = lib().rc_bwd_new(self._d, len(self._d), C.byref(e))
        if not self._h:
            raise RefError(e)
    def take(self, n):
        e = RcError(); v = lib().rc_bwd_take(self._h, n, C.byref(e))
        if e.code: raise RefError(e)
        return v
    def len(self): return lib().rc_bwd_len(self._h)
    def is_empty(self): return self.len() == 0
    def __del__(self):
        if getattr(self, "_h", None): lib().rc_bwd_free(self._h)


def parse_fse_table(data):
    """parse_fse_table (fse.rs:16-69) -> (al, distribution, bits_left, bytes_read)."""
    data = bytes(data); al = C.c_uint8(); dist = (C.c_int16 * 600)(); nd = C.c_size_t(); bl = C.c_size_t(); br = C.c_size_t(); e = RcError()
    if lib().rc_parse_fse_table(data, len(data), C.byref(al), dist, C.byref(nd), C.byref(bl), C.byref(br), C.byref(e)):
        raise RefError(e)
    return al.value, list(dist[:nd.value]), bl.value, br.value


def fse_from_distribution(al, dist):
    """FseTable::from_distribution (fse.rs:110-202) -> list of (output, baseline, bits_to_read)."""
    d = (C.c_int16 * len(dist))(*dist); out = (C.c_uint16 * (3 << max(al, 0) if al <= 9 else 3))(); e = RcError()
    if lib().rc_fse_from_distribution(al, d, len(dist), out, C.byref(e)):
        raise RefError(e)
    return [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(1 << al)]


def fse_run(table, al, stream, count, alternating=False):
    """Decode with an explicit table: returns (symbols, bits_left, error_or_None)."""
    flat = (C.c_uint16 * (3 * len(table)))(*[x for s in table for x in s]); out = (C.c_uint16 * max(count, 1))()
    no = C.c_size_t(); bl = C.c_size_t(); e = RcError(); stream = bytes(stream)
    rc = lib().rc_fse_run(flat, al, int(alternating), stream, len(stream), count, out, C.byref(no), C.byref(bl), C.byref(e))
    return list(out[:no.value]), bl.value, (RefError(e) if rc else None)


def huffman_from_weights(weights):
    """HuffmanDecoder::from_weights -> {symbol: (nbits, code)}."""
    w = bytes(weights); lens = (C.c_uint8 * 257)(); codes = (C.c_uint32 * 257)(); e = RcError()
    if lib().rc_huffman_from_weights(w, len(w), lens, codes, C.byref(e)):
        raise RefError(e)
    return {s: (lens[s], codes[s]) for s in range(256) if lens[s]}


def huffman_parse(data):
    """HuffmanDecoder::parse -> ({symbol: (nbits, code)}, consumed, weights)."""
    d = bytes(data); lens = (C.c_uint8 * 257)(); codes = (C.c_uint32 * 257)(); cons = C.c_size_t(); w = (C.c_uint8 * 4100)(); nw = C.c_size_t(); e = RcError()
    if lib().rc_huffman_parse(d, len(d), lens, codes, C.byref(cons), w, C.byref(nw), C.byref(e)):
        raise RefError(e)
    return {s: (lens[s], codes[s]) for s in range(256) if lens[s]}, cons.value, bytes(w[:nw.value])


def huffman_decode_stream(table, stream, cap=1 << 20):
    lens = (C.c_uint8 * 257)(); codes = (C.c_uint32 * 257)()
    for s, (n, c) in table.items():
        lens[s] = n; codes[s] = c
    out = (C.c_uint8 * cap)(); ol = C.c_size_t(); e = RcError(); stream = bytes(stream)
    if lib().rc_huffman_decode_stream(lens, codes, stream, len(stream), out, cap, C.byref(ol), C.byref(e)):
        raise RefError(e)
    return bytes(out[:ol.value])


def execute_sequences(window, seqs, literals):
    """DecodingContext::new(window).execute_sequences(seqs, literals) -> decoded bytes."""
    flat = (C.c_uint64 * max(3 * len(seqs), 1))(*[x for s in seqs for x in s]); out = C.c_void_p(); ol = C.c_size_t(); e = RcError(); literals = bytes(literals)
    rc = lib().rc_execute_sequences(window, flat, len(seqs), literals, len(literals), C.byref(out), C.byref(ol), C.byref(e))
    data = _take_buf(out, ol.value)
    if rc:
        raise RefError(e)
    return data


def header_parse(data):
    d = bytes(data); h = RcHeader(); cons = C.c_size_t(); e = RcError()
    if lib().rc_header_parse(d, len(d), C.byref(h), C.byref(cons), C.byref(e)):
        raise RefError(e)
    return h, cons.value


def decode_frames(data, quirks=True):
    """Iterate + decode every frame.  Returns (output, [frame info dict], error_or_None)."""
    d = bytes(data); out = C.c_void_p(); ol = C.c_size_t(); fr = C.c_void_p(); nf = C.c_size_t(); e = RcError()
    rc = lib().rc_decode_frames(d, len(d), int(quirks), C.byref(out), C.byref(ol), C.byref(fr), C.byref(nf), C.byref(e))
    frames = []
    if fr.value:
        arr = C.cast(fr, C.POINTER(RcFrameInfo))
        for i in range(nf.value):
            f = arr[i]
            frames.append(dict(kind=f.kind, magic=f.magic, src_off=f.src_off, src_len=f.src_len, out_off=f.out_off, out_len=f.out_len,
                               n_blocks=f.n_blocks, has_checksum=bool(f.has_checksum), stored_checksum=f.stored_checksum,
                               computed_xxh64_low32=f.computed_xxh64_low32, window_size=f.header.window_size,
                               content_size=(f.header.content_size if f.header.has_content_size else None),
                               dict_id=(f.header.dictionnary_id if f.header.has_dict_id else None)))
        lib().rc_free(fr)
    data_out = _take_buf(out, ol.value)
    return data_out, frames, (RefError(e) if rc else None)


def main_decode(data, print_skippable=False, threads=0):
    """src/main.rs default mode.  Raises RefError on the first error (no partial output)."""
    d = bytes(data); out = C.c_void_p(); ol = C.c_size_t(); e = RcError()
    if threads and threads > 0:
        rc = lib().rc_main_decode_mt(d, len(d), int(print_skippable), threads, C.byref(out), C.byref(ol), C.byref(e))
    else:
        rc = lib().rc_main_decode(d, len(d), int(print_skippable), C.byref(out), C.byref(ol), C.byref(e))
    res = _take_buf(out, ol.value)
    if rc:
        raise RefError(e)
    return res


def xxh64(data, seed=0):
    d = bytes(data)
    return lib().rc_xxh64(d, len(d), seed)
