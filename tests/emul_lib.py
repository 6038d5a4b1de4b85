"""ctypes binding of tests/emul (CPU build of the lane-serial device code).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "tests", "emul", "_build", "libzsb_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True)
        _lib = C.CDLL(_SO)
    return _lib


def fse_parse(desc, type_=3):
    desc = bytes(desc); al = C.c_int(); ns = C.c_int(); dist = (C.c_int16 * 256)(); cells = (C.c_uint32 * 512)(); cons = C.c_uint32()
    rc = lib().emul_fse_parse(desc, len(desc), type_, C.byref(al), C.byref(ns), dist, cells, C.byref(cons))
    if rc:
        return rc, None
    return 0, dict(al=al.value, dist=list(dist[:ns.value]), cells=list(cells[:1 << al.value]), consumed=cons.value)


def fse_build(type_, al, dist, stride=1):
    d = (C.c_int16 * len(dist))(*dist); cells = (C.c_uint32 * (512 * stride))()
    rc = lib().emul_fse_build(type_, al, d, len(dist), cells, stride)
    return rc, [cells[i * stride] for i in range(1 << al)]


def cell_fields(c):
    """(code, base, nb, xb)"""
    return ((c >> 16) & 0x3F, c >> 22, (c >> 8) & 0x1F, c & 0x3F)


def huf_parse(desc):
    desc = bytes(desc); lens = (C.c_uint8 * 256)(); lut = (C.c_uint16 * 2048)(); mb = C.c_int(); cons = C.c_uint32(); w = (C.c_uint8 * 260)(); nw = C.c_int()
    rc = lib().emul_huf_parse(desc, len(desc), lens, lut, C.byref(mb), C.byref(cons), w, C.byref(nw))
    if rc:
        return rc, None
    codes = {}
    for i in range((1 << mb.value) - 1, -1, -1):
        s, nb = lut[i] & 255, lut[i] >> 8
        codes[s] = (nb, i >> (mb.value - nb))
    return 0, dict(lens=list(lens), codes=codes, maxbits=mb.value, consumed=cons.value, weights=bytes(w[:nw.value]))


def decode(data, flags=0, cap=None):
    """Full pipeline on the CPU build.  Returns (scan_rc, output, [(status, off, len)], (err_a, err_b))."""
    data = bytes(data)
    cap = cap if cap is not None else max(64 * len(data), 1 << 20)
    out = C.create_string_buffer(cap); ol = C.c_uint64(); nfc = 1 << 17
    st = (C.c_int32 * nfc)(); fo = (C.c_uint64 * nfc)(); fl = (C.c_uint64 * nfc)(); nf = C.c_size_t(); ea = C.c_uint64(); eb = C.c_uint64()
    L = lib()
    L.emul_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_int32),
                              C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    rc = L.emul_decode(data, len(data), flags, out, cap, C.byref(ol), st, fo, fl, nfc, C.byref(nf), C.byref(ea), C.byref(eb))
    frames = [(st[i], fo[i], fl[i]) for i in range(nf.value)]
    return rc, out.raw[:ol.value], frames, (ea.value, eb.value)


def fast_stats():
    """(ran, same, slow, diff): blocks on which the fast sequence path ran, agreed with the careful decoder,
    asked for the careful decoder on a stream the careful decoder also rejects, or disagreed"""
    v = [C.c_long() for _ in range(4)]
    lib().emul_fast_stats(*[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def hist_assoc(seed, n):
    L = lib()
    L.emul_hist_assoc.argtypes = [C.c_uint64, C.c_int]
    return L.emul_hist_assoc(seed, n)


def huf_stats():
    """like fast_stats, for the fast Huffman stream decode (one count per stream)"""
    v = [C.c_long() for _ in range(4)]
    lib().emul_huf_stats(*[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def huf_pairs(weights):
    """-> (rc, maxbits, one-symbol table [(symbol, bits)] of 2^maxbits cells, two-symbol table built as k_huf builds it, the same derived from the
    one-symbol table) for `weights` (the implied last weight not included)"""
    L = lib()
    w = (C.c_uint8 * len(weights))(*weights)
    lut = (C.c_uint16 * 2048)(); pa = (C.c_uint32 * 1024)(); pb = (C.c_uint32 * 1024)(); mb = C.c_int()
    L.emul_huf_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    rc = L.emul_huf_pairs(w, len(weights), lut, pa, pb, C.byref(mb))
    return rc, mb.value, [(v & 0xFF, v >> 8) for v in lut[:1 << mb.value]] if rc == 0 else [], list(pa), list(pb)
