"""Synthetic corpus generator for the BASELINE.json configs (C1-C5).

Compression is done by the system libzstd 1.5.5 through ctypes (it is only the INPUT generator; the
decode path under test never calls it).  Deterministic: fixed seeds, fixed parameters.  The text
source is the decoded moby-dick fixture, the only sizeable text available offline.
"""
import ctypes as C
import os
import random
import struct
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import zasm  # noqa: E402

FIXTURES = os.path.join(ROOT, "tests", "fixtures")

# ZSTD_cParameter ids (zstd.h 1.5.5)
C_LEVEL, C_WINDOWLOG, C_LDM, C_CONTENTSIZE, C_CHECKSUM, C_DICTID = 100, 101, 160, 200, 201, 202
C_LITMODE = 1002      # ZSTD_c_literalCompressionMode (experimental): 1 = huffman, 2 = uncompressed
C_TARGET_CBLOCK = 130  # ZSTD_c_targetCBlockSize

_z = None


def libzstd():
    global _z
    if _z is None:
        z = C.CDLL("libzstd.so.1")
        z.ZSTD_createCCtx.restype = C.c_void_p
        z.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        z.ZSTD_CCtx_setParameter.restype = C.c_size_t
        z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        z.ZSTD_compress2.restype = C.c_size_t
        z.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        z.ZSTD_compressBound.restype = C.c_size_t
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_isError.restype = C.c_uint
        z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_decompress.restype = C.c_size_t
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        z.ZSTD_getErrorName.restype = C.c_char_p
        z.ZSTD_getErrorName.argtypes = [C.c_size_t]
        z.ZSTD_createDCtx.restype = C.c_void_p
        z.ZSTD_freeDCtx.argtypes = [C.c_void_p]
        z.ZSTD_decompressDCtx.restype = C.c_size_t
        z.ZSTD_decompressDCtx.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        z.ZSTD_DCtx_setParameter.restype = C.c_size_t
        z.ZSTD_DCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _z = z
    return _z


class Compressor:
    def __init__(self, level=3, checksum=True, content_size=True, window_log=None, ldm=False, extra=None):
        z = libzstd()
        self.c = z.ZSTD_createCCtx()
        params = {C_LEVEL: level, C_CHECKSUM: int(checksum), C_CONTENTSIZE: int(content_size), C_LDM: int(ldm)}
        if window_log:
            params[C_WINDOWLOG] = window_log
        params.update(extra or {})
        for k, v in params.items():
            r = z.ZSTD_CCtx_setParameter(self.c, k, v)
            if z.ZSTD_isError(r):
                raise RuntimeError(f"ZSTD_CCtx_setParameter({k},{v}): {z.ZSTD_getErrorName(r).decode()}")

    def compress(self, data: bytes) -> bytes:
        z = libzstd()
        cap = z.ZSTD_compressBound(len(data))
        buf = C.create_string_buffer(cap)
        n = z.ZSTD_compress2(self.c, buf, cap, data, len(data))
        if z.ZSTD_isError(n):
            raise RuntimeError(z.ZSTD_getErrorName(n).decode())
        return buf.raw[:n]

    def __del__(self):
        if getattr(self, "c", None):
            libzstd().ZSTD_freeCCtx(self.c)


def compress(data, **kw):
    return Compressor(**kw).compress(data)


def libzstd_decompress(blob: bytes, out_size: int) -> bytes:
    """Decode ONE-or-more concatenated frames with libzstd (skippable frames are skipped)."""
    z = libzstd()
    buf = C.create_string_buffer(max(out_size, 1))
    d = z.ZSTD_createDCtx()
    z.ZSTD_DCtx_setParameter(d, 100, 31)      # ZSTD_d_windowLogMax
    n = z.ZSTD_decompressDCtx(d, buf, max(out_size, 1), blob, len(blob))
    z.ZSTD_freeDCtx(d)
    if z.ZSTD_isError(n):
        raise RuntimeError(z.ZSTD_getErrorName(n).decode())
    return buf.raw[:n]


_text = None


def moby_text() -> bytes:
    global _text
    if _text is None:
        with open(os.path.join(FIXTURES, "moby-dick.txt.zst"), "rb") as f:
            _text = libzstd_decompress(f.read(), 1276235)
        assert len(_text) == 1276235
    return _text


def text_frames(n_frames, seed, frame_size=131072, level=3, threads=None):
    """C2 / C5: frame i = frame_size-byte window of the text at Random(seed).randrange(len - frame_size)
    (one draw per frame, in order); level 3, content size + checksum on.
    Returns (list of compressed frames, list of window offsets)."""
    text = moby_text()
    r = random.Random(seed)
    offs = [r.randrange(len(text) - frame_size) for _ in range(n_frames)]
    threads = threads or min(32, (os.cpu_count() or 4))
    chunks = [list(range(i, n_frames, threads)) for i in range(threads)]
    out = [None] * n_frames

    def work(idx):
        c = Compressor(level=level)
        for i in idx:
            out[i] = c.compress(text[offs[i]:offs[i] + frame_size])

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, chunks))
    return out, offs


def make_c1():
    with open(os.path.join(FIXTURES, "moby-dick.txt.zst"), "rb") as f:
        return f.read(), moby_text()


def make_c2(n_frames=4096, seed=2, frame_size=131072):
    frames, offs = text_frames(n_frames, seed, frame_size)
    text = moby_text()
    return b"".join(frames), b"".join(text[o:o + frame_size] for o in offs)


def make_c2_range(n_frames, seed, f0, f1, frame_size=131072):
    """Frames [f0, f1) of the n_frames-frame corpus of that seed (the windows are drawn for all frames, only the range is compressed)."""
    text = moby_text()
    r = random.Random(seed)
    offs = [r.randrange(len(text) - frame_size) for _ in range(n_frames)][f0:f1]
    threads = min(32, (os.cpu_count() or 4))
    out = [None] * len(offs)

    def work(t):
        c = Compressor(level=3)
        for i in range(t, len(offs), threads):
            out[i] = c.compress(text[offs[i]:offs[i] + frame_size])

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return b"".join(out), b"".join(text[o:o + frame_size] for o in offs)


def make_c5(n_frames=65536, seed=5, frame_size=131072):
    return make_c2(n_frames, seed, frame_size)


def shuffled_tiles(total, first_tile=0):
    """C3 plaintext: tiles of the text with its lines shuffled by Random(tile_index)."""
    lines = moby_text().split(b"\n")
    parts, size, t = [], 0, first_tile
    while size < total:
        l2 = list(lines)
        random.Random(t).shuffle(l2)
        tile = b"\n".join(l2)
        parts.append(tile); size += len(tile); t += 1
    return b"".join(parts)[:total]


def make_c3(total=1 << 30, window_log=23, level=3):
    """Single multi-segment frame, 128 KiB blocks, matches reaching up to 8 MiB back, no LDM."""
    plain = shuffled_tiles(total)
    return compress(plain, level=level, window_log=window_log, ldm=False), plain


def make_c4(seed=4):
    """Mixed corpus: every block / literal / table mode, skippable frames, checksums.
    Returns (blob, expected_without_skippable, expected_with_skippable, labels).
    Only inputs the REFERENCE accepts (no Q1/Q2 classes); see make_rfc_only()."""
    r = random.Random(seed)
    text = moby_text()
    parts = []       # (bytes, plaintext, is_skippable)

    def add(frame, plain, skip=False):
        parts.append((frame, plain, skip))

    for nib in range(16):
        payload = bytes(r.randrange(256) for _ in range(r.randrange(1, 65)))
        add(zasm.skippable(payload, nib), payload, True)
    for n in (200, 600, 2000, 20000):                           # predefined / small FSE tables, 1-stream huffman
        o = r.randrange(len(text) - n)
        add(compress(text[o:o + n]), text[o:o + n])
    rnd = bytes(r.randrange(256) for _ in range(65536))          # incompressible -> raw block
    add(compress(rnd), rnd)
    zeros = bytes(300 * 1024)                                    # RLE blocks
    add(compress(zeros), zeros)
    mix = b"".join(bytes([r.randrange(4) + 65]) * r.randrange(1, 40) for _ in range(4000))   # runs: rle-ish literals, repeat offsets
    add(compress(mix), mix)
    dna = bytes(r.choice(b"ACGT") for _ in range(40000))         # 4-symbol alphabet -> direct huffman weights
    add(compress(dna), dna)
    o = r.randrange(len(text) - 400000)
    big = text[o:o + 400000]
    add(compress(big, level=19), big)                            # repeat mode, block splitting, treeless
    add(compress(big, level=1), big)                             # level 1: bigger literal sections
    add(compress(big[:150000], level=3, extra={C_LITMODE: 2}), big[:150000])   # raw literals in compressed blocks
    per = (b"0123456789abcdef" * 8 + b"\n") * 600                # periodic: single repeated offset
    per = bytearray(per)
    for _ in range(300):
        per[r.randrange(len(per))] = r.randrange(256)
    add(compress(bytes(per)), bytes(per))
    add(compress(big[:200000], level=3, window_log=16, content_size=False), big[:200000])  # window descriptor, no FCS
    for f, p in (zasm.frame_rle_modes(seed), zasm.frame_huffman_direct(seed, streams=4), zasm.frame_huffman_direct(seed + 1, n=700, streams=1),
                 zasm.frame_treeless(seed)):
        add(f, p)
    for f, p in zasm.frame_header_variants(seed):
        add(f, p)
    with open(os.path.join(FIXTURES, "welcome.zst"), "rb") as fh:
        w = fh.read()
    add(w[:56], w[8:56], True)                                   # its skippable frame
    add(w[56:], libzstd_decompress(w[56:], 126))
    order = list(range(len(parts)))
    r.shuffle(order)
    blob = b"".join(parts[i][0] for i in order)
    exp = b"".join(parts[i][1] for i in order if not parts[i][2])
    exp_skip = b"".join(parts[i][1] for i in order)
    return blob, exp, exp_skip, [parts[i] for i in order]


def make_rfc_only(seed=6):
    """Inputs that are valid RFC 8878 but that the reference rejects: checked against
    libzstd only, never a reference-parity row."""
    r = random.Random(seed)
    out = []
    out.append((compress(b""), b""))                                             # empty frame: size-0 raw block (Q2)
    lit_only = bytes(r.randrange(64) + 32 for _ in range(3000))                   # literal-only block: nseq = 0 (Q1)
    out.append((compress(lit_only), lit_only))
    text = moby_text()
    rep = text[1000:60000] * 6                                                    # LDM: zero-literal blocks (Q2)
    out.append((compress(rep, ldm=True, window_log=20), rep))
    out.append((zasm.skippable(b"", 3), b""))                                     # empty skippable (Q2)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--size", type=int, default=None)
    a = ap.parse_args()
    if a.config == "c2":
        blob, _ = make_c2(a.frames or 4096)
    elif a.config == "c5":
        blob, _ = make_c5(a.frames or 65536)
    elif a.config == "c3":
        blob, _ = make_c3(a.size or (1 << 30))
    else:
        blob = make_c4()[0]
    with open(a.output, "wb") as f:
        f.write(blob)
    print(len(blob))
