"""Hand assembler for Zstandard frames that libzstd will not emit on demand.

Builds, bit by bit, frames that exercise the table / literal modes no reference fixture covers
(SURVEY.md 8c): RLE-mode sequence tables, treeless literals, 1-stream and 4-stream Huffman with
direct (4-bit) weights, RLE and raw literals inside compressed blocks, window-descriptor and
dict-id frame headers, RLE / raw blocks, skippable frames.  Pure Python; independent of the GPU
library and of the CPU oracle.  Every builder returns (frame_bytes, expected_plaintext).
"""
import heapq
import struct
from typing import Dict, List, Sequence, Tuple

# ---------------------------------------------------------------- XXH64 (for the content checksum)
_M = (1 << 64) - 1
_P1, _P2, _P3, _P4, _P5 = 0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5


def _rotl(x, r): return ((x << r) | (x >> (64 - r))) & _M
def _round(acc, v): return (_rotl((acc + v * _P2) & _M, 31) * _P1) & _M
def _merge(h, v): return (((h ^ _round(0, v)) * _P1) + _P4) & _M


def xxh64(data: bytes, seed: int = 0) -> int:
    n, p = len(data), 0
    if n >= 32:
        v = [(seed + _P1 + _P2) & _M, (seed + _P2) & _M, seed, (seed - _P1) & _M]
        while p + 32 <= n:
            for i in range(4):
                v[i] = _round(v[i], int.from_bytes(data[p + 8 * i:p + 8 * i + 8], "little"))
            p += 32
        h = (_rotl(v[0], 1) + _rotl(v[1], 7) + _rotl(v[2], 12) + _rotl(v[3], 18)) & _M
        for i in range(4):
            h = _merge(h, v[i])
    else:
        h = (seed + _P5) & _M
    h = (h + n) & _M
    while p + 8 <= n:
        h ^= _round(0, int.from_bytes(data[p:p + 8], "little")); h = (_rotl(h, 27) * _P1 + _P4) & _M; p += 8
    if p + 4 <= n:
        h ^= (int.from_bytes(data[p:p + 4], "little") * _P1) & _M; h = (_rotl(h, 23) * _P2 + _P3) & _M; p += 4
    while p < n:
        h ^= (data[p] * _P5) & _M; h = (_rotl(h, 11) * _P1) & _M; p += 1
    h ^= h >> 33; h = (h * _P2) & _M; h ^= h >> 29; h = (h * _P3) & _M; h ^= h >> 32
    return h


# ---------------------------------------------------------------- backward bitstream writer
class BackBits:
    """Collects fields in the order the decoder will READ them; emit() lays them out so that the first
    field sits just below the end marker of the last byte (RFC 8878 4.1 / parsing.rs:200-254)."""
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value: int, nbits: int):
        assert 0 <= value < (1 << nbits) or nbits == 0
        self.v = (self.v << nbits) | value
        self.n += nbits

    def emit(self) -> bytes:
        total = (1 << self.n) | self.v            # end marker on top
        return total.to_bytes((self.n + 1 + 7) // 8, "little")


# ---------------------------------------------------------------- Huffman
def huffman_lengths(freq: Dict[int, int], max_bits: int = 11) -> Dict[int, int]:
    """Code lengths of a complete prefix code (>= 2 symbols)."""
    assert len(freq) >= 2
    heap = [(f, s, (s,)) for s, f in freq.items()]
    heapq.heapify(heap)
    depth = {s: 0 for s in freq}
    while len(heap) > 1:
        f1, k1, g1 = heapq.heappop(heap)
        f2, k2, g2 = heapq.heappop(heap)
        for s in g1 + g2:
            depth[s] += 1
        heapq.heappush(heap, (f1 + f2, min(k1, k2), g1 + g2))
    assert max(depth.values()) <= max_bits, "alphabet too skewed for this tiny builder"
    return depth


def huffman_codes(lengths: Dict[int, int]) -> Dict[int, Tuple[int, int]]:
    """Canonical zstd code: longest codes first, ascending symbol within a length, counting from 0."""
    order = sorted(lengths, key=lambda s: (-lengths[s], s))
    codes, code, prev = {}, 0, lengths[order[0]]
    for s in order:
        L = lengths[s]
        code >>= (prev - L)
        codes[s] = (L, code)
        code += 1
        prev = L
    return codes


def huffman_direct_description(lengths: Dict[int, int]) -> bytes:
    """Tree description with direct 4-bit weights (header byte >= 128); last symbol's weight implied."""
    maxbits = max(lengths.values())
    last = max(lengths)
    nw = last                                   # weights for symbols 0 .. last-1
    assert 1 <= nw <= 128
    w = [(maxbits + 1 - lengths[s]) if s in lengths else 0 for s in range(nw)]
    if nw % 2:
        w.append(0)
    return bytes([127 + nw]) + bytes((w[i] << 4) | w[i + 1] for i in range(0, len(w), 2))


def huffman_encode_stream(codes, data: bytes) -> bytes:
    # The decoder reads from the END of the stream and yields literals in forward order, so the first
    # literal's code sits just below the end marker.
    bw = BackBits()
    for s in data:
        L, c = codes[s]
        bw.put(c, L)
    return bw.emit()


def literals_raw(lits: bytes) -> bytes:
    n = len(lits)
    if n < 32:
        return bytes([(n << 3) | 0]) + lits
    if n < 4096:
        return bytes([((n & 15) << 4) | (1 << 2) | 0, n >> 4]) + lits
    return bytes([((n & 15) << 4) | (3 << 2) | 0, (n >> 4) & 255, n >> 12]) + lits


def literals_rle(byte: int, n: int) -> bytes:
    if n < 32:
        return bytes([(n << 3) | 1, byte])
    if n < 4096:
        return bytes([((n & 15) << 4) | (1 << 2) | 1, n >> 4, byte])
    return bytes([((n & 15) << 4) | (3 << 2) | 1, (n >> 4) & 255, n >> 12, byte])


def literals_huffman(lits: bytes, lengths: Dict[int, int], streams: int = 4, treeless: bool = False) -> bytes:
    codes = huffman_codes(lengths)
    desc = b"" if treeless else huffman_direct_description(lengths)
    n = len(lits)
    if streams == 1:
        body = huffman_encode_stream(codes, lits)
    else:
        q = (n + 3) // 4
        parts = [huffman_encode_stream(codes, lits[i * q:(i + 1) * q]) for i in range(3)] + [huffman_encode_stream(codes, lits[3 * q:])]
        body = b"".join(struct.pack("<H", len(p)) for p in parts[:3]) + b"".join(parts)
    csize = len(desc) + len(body)
    lt = 3 if treeless else 2
    if streams == 1:
        assert n < 1024 and csize < 1024
        v = (n >> 4) | (csize << 6)
        hdr = bytes([((n & 15) << 4) | (0 << 2) | lt]) + v.to_bytes(2, "little")
    elif n < 1024 and csize < 1024:
        v = (n >> 4) | (csize << 6)
        hdr = bytes([((n & 15) << 4) | (1 << 2) | lt]) + v.to_bytes(2, "little")
    elif n < 16384 and csize < 16384:
        v = (n >> 4) | (csize << 10)
        hdr = bytes([((n & 15) << 4) | (2 << 2) | lt]) + v.to_bytes(3, "little")
    else:
        v = (n >> 4) | (csize << 14)
        hdr = bytes([((n & 15) << 4) | (3 << 2) | lt]) + v.to_bytes(4, "little")
    return hdr + desc + body


# ---------------------------------------------------------------- sequences with RLE-mode tables
LL_BASE = list(range(16)) + [16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
LL_BITS = [0] * 16 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
ML_BASE = list(range(3, 35)) + [35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539]
ML_BITS = [0] * 32 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]


def nseq_bytes(n: int) -> bytes:
    if n < 128:
        return bytes([n])
    if n < 0x7F00:
        return bytes([(n >> 8) + 128, n & 255])
    return bytes([255]) + struct.pack("<H", n - 0x7F00)


def sequences_rle(ll_code: int, of_code: int, ml_code: int, extras: Sequence[Tuple[int, int, int]]) -> bytes:
    """Sequences section whose three tables are all in RLE mode (one code each).  extras: per sequence
    (of_extra, ml_extra, ll_extra) raw extra-bit values.  State updates read 0 bits."""
    bw = BackBits()
    for ofx, mlx, llx in extras:
        bw.put(ofx, of_code); bw.put(mlx, ML_BITS[ml_code]); bw.put(llx, LL_BITS[ll_code])
    modes = (1 << 6) | (1 << 4) | (1 << 2)
    return nseq_bytes(len(extras)) + bytes([modes, ll_code, of_code, ml_code]) + bw.emit()


def rle_sequence_values(ll_code, of_code, ml_code, extras):
    return [(LL_BASE[ll_code] + llx, (1 << of_code) + ofx, ML_BASE[ml_code] + mlx) for ofx, mlx, llx in extras]


def execute(seqs: List[Tuple[int, int, int]], lits: bytes, history: bytes = b"", rep=None) -> Tuple[bytes, list]:
    """Reference semantics of sequence execution (RFC 8878 3.1.1.4/5) for building expected output."""
    out = bytearray(history)
    rep = list(rep or [1, 4, 8])
    lp = 0
    for ll, ov, ml in seqs:
        if ov > 3:
            off = ov - 3; rep = [off, rep[0], rep[1]]
        elif ll:
            if ov == 1: off = rep[0]
            elif ov == 2: off = rep[1]; rep = [rep[1], rep[0], rep[2]]
            else: off = rep[2]; rep = [rep[2], rep[0], rep[1]]
        else:
            if ov == 1: off = rep[1]; rep = [rep[1], rep[0], rep[2]]
            elif ov == 2: off = rep[2]; rep = [rep[2], rep[0], rep[1]]
            else: off = rep[0] - 1; rep = [off, rep[0], rep[1]]
        out += lits[lp:lp + ll]; lp += ll
        assert 0 < off <= len(out)
        for _ in range(ml):
            out.append(out[-off])
    out += lits[lp:]
    return bytes(out[len(history):]), rep


# ---------------------------------------------------------------- blocks / frames
def block(btype: int, payload: bytes, last: bool, size: int = None) -> bytes:
    size = len(payload) if size is None else size
    return ((size << 3) | (btype << 1) | int(last)).to_bytes(3, "little") + payload


def raw_block(data: bytes, last: bool) -> bytes: return block(0, data, last)
def rle_block(byte: int, n: int, last: bool) -> bytes: return block(1, bytes([byte]), last, n)
def compressed_block(literals: bytes, sequences: bytes, last: bool) -> bytes: return block(2, literals + sequences, last)


def frame(blocks: bytes, content: bytes, *, checksum=True, single_segment=True, window_log=None,
          dict_id=None, fcs=True) -> bytes:
    n = len(content)
    fhd = (4 if checksum else 0)
    hdr = b""
    if single_segment:
        fhd |= 0x20
    else:
        wl = window_log or max(10, (max(n, 1) - 1).bit_length())
        hdr += bytes([(wl - 10) << 3])
    if dict_id is not None:
        dl = 1 if dict_id < 256 else (2 if dict_id < 65536 else 4)
        fhd |= {1: 1, 2: 2, 4: 3}[dl]
        hdr += dict_id.to_bytes(dl, "little")
    if single_segment or fcs:
        if single_segment and n < 256:
            fl, code = 1, 0
        elif n < 256:
            fl, code = 4, 2
        elif 256 <= n < 65536 + 256:
            fl, code = 2, 1
        elif n < (1 << 32):
            fl, code = 4, 2
        else:
            fl, code = 8, 3
        fhd |= code << 6
        hdr += (n - 256 if fl == 2 else n).to_bytes(fl, "little")
    out = struct.pack("<I", 0xFD2FB528) + bytes([fhd]) + hdr + blocks
    if checksum:
        out += struct.pack("<I", xxh64(content) & 0xFFFFFFFF)
    return out


def skippable(payload: bytes, nibble: int = 0) -> bytes:
    return struct.pack("<II", 0x184D2A50 | (nibble & 15), len(payload)) + payload


# ---------------------------------------------------------------- ready-made special frames
def _letters(n, seed):
    import random
    r = random.Random(seed)
    # skewed alphabet so that code lengths differ (1..5 bits)
    alpha = b"etaoinshrdlu"
    weights = [40, 20, 10, 8, 6, 5, 4, 3, 2, 2, 1, 1]
    return bytes(r.choices(alpha, weights)[0] for _ in range(n))


def frame_rle_modes(seed=1, nseq=40) -> Tuple[bytes, bytes]:
    """One compressed block: raw literals + all three sequence tables in RLE mode."""
    import random
    r = random.Random(seed)
    ll_code, of_code, ml_code = 17, 4, 33            # ll 18..19, offset_value 16..31, ml 37..38
    extras = [(r.randrange(1 << of_code) if i else 0, r.randrange(2), r.randrange(2)) for i in range(nseq)]
    seqs = rle_sequence_values(ll_code, of_code, ml_code, extras)
    lits = _letters(sum(s[0] for s in seqs) + 7, seed)
    content, _ = execute(seqs, lits)
    blk = compressed_block(literals_raw(lits), sequences_rle(ll_code, of_code, ml_code, extras), True)
    return frame(blk, content), content


def frame_huffman_direct(seed=2, n=3000, streams=4) -> Tuple[bytes, bytes]:
    """Huffman literals with direct weights (1 or 4 streams) + RLE-mode sequences using repeat offsets."""
    lits = _letters(n, seed)
    freq = {}
    for c in lits:
        freq[c] = freq.get(c, 0) + 1
    lengths = huffman_lengths(freq)
    ll_code, of_code, ml_code = 20, 0, 2             # ll 24..27, offset_value 1 (repeat offset 1), ml 5
    nseq = max(1, (n - 10) // 28)
    import random
    r = random.Random(seed)
    extras = [(0, 0, r.randrange(4)) for _ in range(nseq)]
    seqs = rle_sequence_values(ll_code, of_code, ml_code, extras)
    content, _ = execute(seqs, lits)
    blk = compressed_block(literals_huffman(lits, lengths, streams), sequences_rle(ll_code, of_code, ml_code, extras), True)
    return frame(blk, content), content


def frame_treeless(seed=3, n=2000) -> Tuple[bytes, bytes]:
    """Two compressed blocks; the second reuses the first block's Huffman tree (treeless literals),
    uses RLE literals in a third block and a raw + RLE block in between (multi-block frame, offsets
    reaching into earlier blocks, repeat-offset history carried across blocks)."""
    lits1, lits2 = _letters(n, seed), _letters(n // 2, seed + 100)
    freq = {c: 1 for c in b"etaoinshrdlu"}
    for c in lits1 + lits2:
        freq[c] += 1
    lengths = huffman_lengths(freq)
    ll_code, of_code, ml_code = 22, 5, 40            # ll 32..39, offset_value 32..63, ml 67..82
    import random
    r = random.Random(seed)
    def mk(lits, k):
        extras = [(r.randrange(1 << of_code) if i else 0, r.randrange(16), r.randrange(8)) for i in range(k)]
        return extras, rle_sequence_values(ll_code, of_code, ml_code, extras)
    e1, s1 = mk(lits1, len(lits1) // 40)
    out1, rep = execute(s1, lits1)
    b1 = compressed_block(literals_huffman(lits1, lengths, 4), sequences_rle(ll_code, of_code, ml_code, e1), False)
    raw = bytes(r.randrange(256) for _ in range(300))
    b2 = raw_block(raw, False)
    b3 = rle_block(0x5A, 1000, False)
    hist = out1 + raw + bytes([0x5A]) * 1000
    e4, s4 = mk(lits2, len(lits2) // 40)
    out4, rep = execute(s4, lits2, hist, rep)
    b4 = compressed_block(literals_huffman(lits2, lengths, 4, treeless=True), sequences_rle(ll_code, of_code, ml_code, e4), False)
    hist += out4
    # RLE literals + far offsets (code 10: offset_value 1024..2047) + repeat codes with ll == 0
    ll5, of5, ml5 = 3, 10, 10
    e5 = [(r.randrange(1 << of5), 0, 0) for _ in range(30)]
    s5 = rle_sequence_values(ll5, of5, ml5, e5)
    lits5 = bytes([0x2E]) * (sum(s[0] for s in s5) + 5)
    out5, rep = execute(s5, lits5, hist, rep)
    b5 = compressed_block(literals_rle(0x2E, len(lits5)), sequences_rle(ll5, of5, ml5, e5), False)
    hist += out5
    # repeat-offset codes with literal length 0: ll code 0, of code 1 (offset_value 2..3)
    e6 = [(r.randrange(2), 0, 0) for _ in range(25)]
    s6 = rle_sequence_values(0, 1, 5, e6)
    out6, rep = execute(s6, b"tail!", hist, rep)
    b6 = compressed_block(literals_raw(b"tail!"), sequences_rle(0, 1, 5, e6), True)
    content = hist + out6
    return frame(b1 + b2 + b3 + b4 + b5 + b6, content, single_segment=False, window_log=17), content


def frame_header_variants(seed=4) -> List[Tuple[bytes, bytes]]:
    """Raw-block frames covering FCS widths, the window descriptor and the dict-id field."""
    import random
    r = random.Random(seed)
    out = []
    for n, kw in [(5, {}), (200, {}), (300, {}), (70000, {}),
                  (1000, dict(single_segment=False, window_log=10)),
                  (1000, dict(single_segment=False, window_log=12, fcs=False)),
                  (500, dict(dict_id=0x77)), (500, dict(dict_id=0xABEF, single_segment=False)), (500, dict(dict_id=0x12345678)),
                  (64, dict(checksum=False))]:
        data = bytes(r.randrange(256) for _ in range(n))
        blocks = b"".join(raw_block(data[i:i + 65536], i + 65536 >= n) for i in range(0, n, 65536))
        out.append((frame(blocks, data, **kw), data))
    return out
