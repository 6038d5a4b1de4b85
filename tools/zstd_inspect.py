"""Pure-Python walker over a .zst buffer: frames -> blocks -> section headers (no entropy decoding).

Used by the corpus generator to label which block / literal / sequence-table modes an input
exercises, and by tests to cross-check the host scanner (zsb_scan).  It follows RFC 8878 section 3
and never touches the GPU library or the CPU oracle.
"""
from dataclasses import dataclass, field
from typing import List, Optional

MAGIC_ZSTD = 0xFD2FB528
MAGIC_SKIP = 0x184D2A50

LIT_TYPES = ("raw", "rle", "compressed", "treeless")
SEQ_MODES = ("predefined", "rle", "fse", "repeat")
BLOCK_TYPES = ("raw", "rle", "compressed", "reserved")


@dataclass
class BlockInfo:
    type: str
    last: bool
    size: int                 # Block_Size field (RLE: repeat count)
    src_off: int              # offset of the payload (after the 3-byte header)
    lit_type: Optional[str] = None
    lit_regen: int = 0
    lit_csize: int = 0
    lit_streams: int = 0
    huf_header: Optional[int] = None     # <128 FSE-compressed weights, >=128 direct
    nseq: int = 0
    modes: Optional[tuple] = None        # (LL, OF, ML)


@dataclass
class FrameInfo:
    kind: str                 # "zstd" | "skippable"
    magic: int
    src_off: int
    src_len: int = 0
    window_size: int = 0
    content_size: Optional[int] = None
    dict_id: Optional[int] = None
    single_segment: bool = False
    has_checksum: bool = False
    checksum: Optional[int] = None
    blocks: List[BlockInfo] = field(default_factory=list)
    payload_len: int = 0      # skippable


def _fse_desc_len(b: bytes, pos: int, end: int) -> int:
    """Bytes occupied by an FSE table description starting at pos (RFC 8878 4.1.1)."""
    bitpos = 0

    def take(n):
        nonlocal bitpos
        v = 0
        for i in range(n):
            p = bitpos + i
            byte = b[pos + (p >> 3)] if pos + (p >> 3) < end else 0
            v |= ((byte >> (p & 7)) & 1) << i
        bitpos += n
        return v

    def peek(n):
        nonlocal bitpos
        save = bitpos
        v = take(n)
        bitpos = save
        return v

    al = take(4) + 5
    remaining = 1 << al
    nsym = 0
    while remaining > 0 and nsym < 256:
        nb = (remaining + 1).bit_length()
        pk = peek(nb)
        low = (1 << (nb - 1)) - 1
        thr = (1 << nb) - 1 - (remaining + 1)
        if (pk & low) < thr:
            v = take(nb - 1)
        elif pk > low:
            v = take(nb) - thr
        else:
            v = take(nb)
        proba = v - 1
        remaining -= abs(proba)
        nsym += 1
        if proba == 0:
            while True:
                z = take(2)
                nsym += z
                if z != 3:
                    break
    return (bitpos + 7) // 8


def _parse_compressed_block(b: bytes, pos: int, size: int, blk: BlockInfo):
    end = pos + size
    h = b[pos]
    lt, sf = h & 3, (h >> 2) & 3
    blk.lit_type = LIT_TYPES[lt]
    if lt < 2:
        if sf in (0, 2):
            blk.lit_regen, hl = h >> 3, 1
        elif sf == 1:
            blk.lit_regen, hl = (h >> 4) + (b[pos + 1] << 4), 2
        else:
            blk.lit_regen, hl = (h >> 4) + (b[pos + 1] << 4) + (b[pos + 2] << 12), 3
        blk.lit_streams = 1
        p = pos + hl + (blk.lit_regen if lt == 0 else 1)
    else:
        extra = 2 if sf <= 1 else (3 if sf == 2 else 4)
        v = int.from_bytes(b[pos + 1:pos + 1 + extra], "little")
        rb, cb = (6, 10) if sf <= 1 else ((10, 14) if sf == 2 else (14, 18))
        blk.lit_regen = (h >> 4) + ((v & ((1 << rb) - 1)) << 4)
        blk.lit_csize = (v >> rb) & ((1 << cb) - 1)
        blk.lit_streams = 1 if sf == 0 else 4
        if lt == 2:
            blk.huf_header = b[pos + 1 + extra]
        p = pos + 1 + extra + blk.lit_csize
    b0 = b[p]
    if b0 < 128:
        blk.nseq, p = b0, p + 1
    elif b0 < 255:
        blk.nseq, p = ((b0 - 128) << 8) + b[p + 1], p + 2
    else:
        blk.nseq, p = b[p + 1] + (b[p + 2] << 8) + 0x7F00, p + 3
    if blk.nseq:
        m = b[p]
        blk.modes = (SEQ_MODES[(m >> 6) & 3], SEQ_MODES[(m >> 4) & 3], SEQ_MODES[(m >> 2) & 3])
    assert p <= end


def inspect(b: bytes) -> List[FrameInfo]:
    pos, frames = 0, []
    n = len(b)
    while pos < n:
        magic = int.from_bytes(b[pos:pos + 4], "little")
        start = pos
        pos += 4
        if magic == MAGIC_ZSTD:
            f = FrameInfo("zstd", magic, start)
            fhd = b[pos]; pos += 1
            dflag, f.has_checksum, f.single_segment, csf = fhd & 3, bool(fhd & 4), bool(fhd & 32), fhd >> 6
            if not f.single_segment:
                wd = b[pos]; pos += 1
                base = 1 << (10 + (wd >> 3))
                f.window_size = base + (base // 8) * (wd & 7)
            if dflag:
                dl = 1 << (dflag - 1)
                f.dict_id = int.from_bytes(b[pos:pos + dl], "little"); pos += dl
            fcs = 0 if (csf == 0 and not f.single_segment) else (1 if csf == 0 else 1 << csf)
            if fcs:
                f.content_size = int.from_bytes(b[pos:pos + fcs], "little") + (256 if fcs == 2 else 0); pos += fcs
            if f.single_segment:
                f.window_size = f.content_size
            while True:
                v = int.from_bytes(b[pos:pos + 3], "little"); pos += 3
                blk = BlockInfo(BLOCK_TYPES[(v >> 1) & 3], bool(v & 1), v >> 3, pos)
                if blk.type == "compressed":
                    _parse_compressed_block(b, pos, blk.size, blk)
                pos += 1 if blk.type == "rle" else blk.size
                f.blocks.append(blk)
                if blk.last:
                    break
            if f.has_checksum:
                f.checksum = int.from_bytes(b[pos:pos + 4], "little"); pos += 4
        elif (magic ^ MAGIC_SKIP) <= 0xF:
            f = FrameInfo("skippable", magic, start)
            f.payload_len = int.from_bytes(b[pos:pos + 4], "little"); pos += 4 + f.payload_len
        else:
            raise ValueError(f"unrecognised magic {magic:#x} at {start}")
        f.src_len = pos - start
        frames.append(f)
    return frames


def features(b: bytes) -> set:
    """Set of mode labels exercised by a buffer (used for coverage accounting of C4)."""
    s = set()
    for f in inspect(b):
        if f.kind == "skippable":
            s.add("skippable"); continue
        s.add("single_segment" if f.single_segment else "window_descriptor")
        if f.dict_id is not None: s.add("dict_id")
        if f.has_checksum: s.add("checksum")
        if len(f.blocks) > 1: s.add("multi_block")
        for k in f.blocks:
            s.add("block_" + k.type)
            if k.type != "compressed": continue
            s.add("lit_" + k.lit_type)
            if k.lit_type in ("compressed", "treeless"):
                s.add(f"huf_{k.lit_streams}stream")
            if k.huf_header is not None:
                s.add("huf_weights_fse" if k.huf_header < 128 else "huf_weights_direct")
            if k.nseq == 0:
                s.add("nseq0")
            else:
                for name, m in zip(("ll", "of", "ml"), k.modes):
                    s.add(f"{name}_{m}")
                if k.nseq >= 128: s.add("nseq_2byte")
    return s
