"""What `println!("{:#x?}", frame)` prints for a parsed frame of the reference (src/main.rs:35-40), rebuilt INDEPENDENTLY of the product:
the container and the section headers are parsed here in Python (RFC 8878 as the reference reads it), the Huffman tree and the FSE tables
come from the CPU oracle (tests/refcpu.py: HuffmanDecoder::parse, parse_fse_table, FseTable::from_distribution), and the layout follows
Rust's derived Debug in pretty + hex mode plus the reference's own impl for HuffmanDecoder (huffman.rs:60-77).  Test infrastructure."""
import refcpu as R

I = "    "


def hexlist(data, d):
    if not data:
        return "[]"
    return "[\n" + "".join(f"{I * (d + 1)}{b:#x},\n" for b in data) + I * d + "]"


def opt(name, v, d):
    return f"{I * d}{name}: None,\n" if v is None else f"{I * d}{name}: Some(\n{I * (d + 1)}{v:#x},\n{I * d}),\n"


def mode_str(name, m, d):
    kind, arg = m
    if kind == "predefined":
        return f"{I * d}{name}: PredefinedMode,\n"
    if kind == "repeat":
        return f"{I * d}{name}: RepeatMode,\n"
    if kind == "rle":
        return f"{I * d}{name}: RLEMode(\n{I * (d + 1)}{arg:#x},\n{I * d}),\n"
    al, states = arg
    s = f"{I * d}{name}: FseCompressedMode(\n{I * (d + 1)}FseTable {{\n{I * (d + 2)}table: [\n"
    for o, b, n in states:
        s += f"{I * (d + 3)}State {{\n{I * (d + 4)}output: {o:#x},\n{I * (d + 4)}baseline: {b:#x},\n{I * (d + 4)}bits_to_read: {n:#x},\n{I * (d + 3)}}},\n"
    return s + f"{I * (d + 2)}],\n{I * (d + 2)}al: {al:#x},\n{I * (d + 1)}}},\n{I * d}),\n"


def compressed_block(p, d):
    """p: payload of a compressed block -> its dump at depth d"""
    h = p[0]; lt, sf = h & 3, (h >> 2) & 3
    s = f"{I * d}CompressedBlock {{\n{I * (d + 1)}literals_section: "
    if lt <= 1:
        if sf in (0, 2): regen, q = h >> 3, 1
        elif sf == 1: regen, q = (h >> 4) + (p[1] << 4), 2
        else: regen, q = (h >> 4) + (p[1] << 4) + (p[2] << 12), 3
        if lt == 0:
            s += f"RawLiteralsBlock {{\n{I * (d + 2)}data: {hexlist(p[q:q + regen], d + 2)},\n{I * (d + 1)}}},\n"; q += regen
        else:
            s += f"RLELiteralsBlock {{\n{I * (d + 2)}byte: {p[q]:#x},\n{I * (d + 2)}repeat: {regen:#x},\n{I * (d + 1)}}},\n"; q += 1
    else:
        extra = 2 if sf <= 1 else 3 if sf == 2 else 4
        v = int.from_bytes(p[1:1 + extra], "little"); q = 1 + extra
        rb, cb = (6, 10) if sf <= 1 else (10, 14) if sf == 2 else (14, 18)
        regen = (h >> 4) + ((v & ((1 << rb) - 1)) << 4); csize = (v >> rb) & ((1 << cb) - 1)
        body = p[q:q + csize]; q += csize
        s += f"CompressedLiteralsBlock {{\n{I * (d + 2)}huffman_decoder: "
        if lt == 3:
            s += "None,\n"; used = 0
        else:
            tree, used, _ = R.huffman_parse(body)
            s += f"Some(\n{I * (d + 3)}HuffmanDecoder {{\n"
            for code, sym in sorted((format(c, f"0{n}b"), sy) for sy, (n, c) in tree.items()):
                s += f"{I * (d + 4)} {code}: {sym:#x},\n"
            s += f"{I * (d + 3)}}},\n{I * (d + 2)}),\n"
        rest = body[used:]
        if sf == 0:
            jt, data = [len(rest) & 0xFFFF, 0, 0, 0], rest
        else:
            s1, s2, s3 = (int.from_bytes(rest[2 * k:2 * k + 2], "little") for k in range(3))
            jt, data = [s1, s2, s3, (len(rest) - 6 - s1 - s2 - s3) & 0xFFFF], rest[6:]
        s += f"{I * (d + 2)}regenerated_size: {regen:#x},\n{I * (d + 2)}jump_table: {hexlist(jt, d + 2)},\n{I * (d + 2)}data: {hexlist(data, d + 2)},\n{I * (d + 1)}}},\n"
    # Sequences::parse sequences.rs:52-143
    b0 = p[q]; q += 1
    if b0 < 128: nseq = b0
    elif b0 < 255: nseq = ((b0 - 128) << 8) + p[q]; q += 1
    else: nseq = p[q] + (p[q + 1] << 8) + 0x7F; q += 2
    modes = [("repeat", None)] * 3; bitstream = b""
    if nseq:
        mb = p[q]; q += 1
        modes = []
        for t in range(3):
            m = (mb >> (6 - 2 * t)) & 3
            if m == 0: modes.append(("predefined", None))
            elif m == 3: modes.append(("repeat", None))
            elif m == 1: modes.append(("rle", p[q])); q += 1
            else:
                al, dist, _, used = R.parse_fse_table(p[q:])
                modes.append(("fse", (al, R.fse_from_distribution(al, dist)))); q += used
        bitstream = p[q:]
    s += f"{I * (d + 1)}sequences_section: Sequences {{\n{I * (d + 2)}number_of_sequences: {nseq:#x},\n"
    for name, m in zip(("literal_lengths_mode", "offsets_mode", "match_lengths_mode"), modes):
        s += mode_str(name, m, d + 2)
    return s + f"{I * (d + 2)}bitstream: {hexlist(bitstream, d + 2)},\n{I * (d + 1)}}},\n{I * d}}},\n"


def dump(data):
    """every frame of a well-formed .zst buffer"""
    out, pos = "", 0
    while pos < len(data):
        magic = int.from_bytes(data[pos:pos + 4], "little"); pos += 4
        if (magic ^ 0x184D2A50) <= 0xF:
            n = int.from_bytes(data[pos:pos + 4], "little"); pos += 4
            out += f"SkippableFrame(\n{I}Skippable {{\n{I * 2}magic: {magic:#x},\n{I * 2}data: {hexlist(data[pos:pos + n], 2)},\n{I}}},\n)\n"; pos += n
            continue
        assert magic == 0xFD2FB528
        fhd = data[pos]; pos += 1
        dflag, cks, single, csf = fhd & 3, (fhd >> 2) & 1, (fhd >> 5) & 1, fhd >> 6
        window = None
        if not single:
            wd = data[pos]; pos += 1
            base = 1 << ((wd >> 3) + 10); window = base + (base // 8) * (wd & 7)
        did = None
        if dflag:
            n = 1 << (dflag - 1); did = int.from_bytes(data[pos:pos + n], "little"); pos += n
        fcs_n = 0 if (csf == 0 and not single) else (1 if csf == 0 else 1 << csf)
        fcs = None
        if fcs_n:
            fcs = int.from_bytes(data[pos:pos + fcs_n], "little") + (256 if fcs_n == 2 else 0); pos += fcs_n
        if single: window = fcs
        out += (f"ZStandardFrame(\n{I}ZStandard {{\n{I * 2}header: Header {{\n{I * 3}content_checksum_flag: {'true' if cks else 'false'},\n{I * 3}window_size: {window:#x},\n"
                + opt("dictionnary_id", did, 3) + opt("content_size", fcs, 3) + f"{I * 2}}},\n{I * 2}blocks: [\n")
        while True:
            v = int.from_bytes(data[pos:pos + 3], "little"); pos += 3
            last, bt, size = v & 1, (v >> 1) & 3, v >> 3
            if bt == 0:
                out += f"{I * 3}RawBlock(\n{I * 4}{hexlist(data[pos:pos + size], 4)},\n{I * 3}),\n"; pos += size
            elif bt == 1:
                out += f"{I * 3}RLEBlock {{\n{I * 4}byte: {data[pos]:#x},\n{I * 4}repeat: {size:#x},\n{I * 3}}},\n"; pos += 1
            else:
                out += compressed_block(data[pos:pos + size], 3); pos += size
            if last: break
        ck = None
        if cks:
            ck = int.from_bytes(data[pos:pos + 4], "little"); pos += 4
        out += f"{I * 2}],\n" + opt("checksum", ck, 2) + f"{I}}},\n)\n"
    return out
