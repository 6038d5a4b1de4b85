// zsb_walk.h -- the walk over ONE frame: magic, frame header, block headers, checksum.
//
// == Frame::parse + Header::parse + ZStandard::parse + Block::parse of the reference (frame.rs:61-230, block.rs:43-72), stopping at block
//    extents.  One implementation for both walkers: the host scanner (zsb_scan.cpp, ZsbScanner::next) and the device scanner (zsb_dscan.cu,
//    one lane per candidate frame start), so that the two give the same descriptors by construction.  `emit(block)` is called for every block in
//    order (the host appends to its vector, the device counts in its first pass and writes in its second).
#pragma once
#include "zsb_common.h"

#define ZSB_MAGIC_ZSTD 0xFD2FB528u
#define ZSB_MAGIC_SKIP 0x184D2A50u

struct ZsbWalkErr { int code; uint64_t a, b; };

ZSB_HD uint64_t zsb_rd_le(const uint8_t *p, int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i); return v; }
ZSB_HD bool zsb_is_frame_magic(uint32_t m) { return m == ZSB_MAGIC_ZSTD || (m ^ ZSB_MAGIC_SKIP) <= 0x0Fu; }

// Header::parse frame.rs:111-177.  pos: behind the magic on entry, behind the header on success.
ZSB_HD bool zsb_walk_header(const uint8_t *src, uint64_t n, uint64_t &pos, zsb_frame &f, ZsbWalkErr &e) {
#define ZSB_WALK_NEED(k) do { if (n - pos < (uint64_t)(k)) { e.code = ZSB_E_NOT_ENOUGH_BYTES; e.a = (uint64_t)(k); e.b = n - pos; return false; } } while (0)
    ZSB_WALK_NEED(1);
    const uint8_t b = src[pos++];
    const unsigned dflag = b & 3, checksum = (b >> 2) & 1, reserved = (b >> 3) & 1, single = (b >> 5) & 1, csf = b >> 6;
    if (reserved) { e.code = ZSB_E_FRAME_RESERVED; e.a = e.b = 0; return false; }
    const int fcs = (csf == 0 && !single) ? 0 : (csf == 0 ? 1 : 1 << csf);
    uint64_t window = 0;
    if (!single) {                                                   // parse_window_descriptor frame.rs:179-187
        ZSB_WALK_NEED(1);
        const uint8_t wd = src[pos++];
        const uint64_t base = (uint64_t)1 << ((wd >> 3) + 10);
        window = base + (base / 8) * (wd & 7);
    }
    f.has_dict_id = 0; f.dict_id = 0;
    if (dflag) {
        const uint64_t dl = (uint64_t)1 << (dflag - 1);
        ZSB_WALK_NEED(dl);
        f.dict_id = zsb_rd_le(src + pos, (int)dl); f.has_dict_id = 1; pos += dl;
    }
    f.has_content_size = 0; f.content_size = 0;
    if (fcs) {
        ZSB_WALK_NEED(fcs);
        f.content_size = zsb_rd_le(src + pos, fcs) + (fcs == 2 ? 256 : 0); f.has_content_size = 1; pos += (uint64_t)fcs;
    }
    f.has_checksum = (uint8_t)checksum; f.single_segment = (uint8_t)single;
    f.window_size = single ? f.content_size : window;
    return true;
}

// The frame that starts at `pos`.  On success: f (all fields but first_block / n_blocks / status), `end` = the byte behind the frame, true.
// On failure: e, false; f.kind / f.magic say how far it got, `n_emitted` blocks had been emitted.
template <class Emit>
ZSB_HD bool zsb_walk_frame(const uint8_t *src, uint64_t n, uint64_t pos, uint32_t flags, uint64_t max_window, uint32_t frame_index,
                           zsb_frame &f, uint64_t &end, ZsbWalkErr &e, uint32_t &n_emitted, Emit &&emit) {
    const bool quirks = (flags & ZSB_REFERENCE_QUIRKS) != 0;
    f.src_off = pos; n_emitted = 0;
    ZSB_WALK_NEED(4);                                                // Frame::parse frame.rs:61-77
    const uint32_t magic = (uint32_t)zsb_rd_le(src + pos, 4); pos += 4;
    f.magic = magic;
    if (magic == ZSB_MAGIC_ZSTD) {
        f.kind = 0;
        if (!zsb_walk_header(src, n, pos, f, e)) return false;       // ZStandard::parse frame.rs:198-230
        if (f.window_size > max_window) { e.code = ZSB_E_WINDOW_TOO_BIG; e.a = max_window; e.b = f.window_size; return false; }
        if ((flags & ZSB_STRICT_DICT) && f.has_dict_id && f.dict_id != 0) { e.code = ZSB_E_DICTIONARY; e.a = f.dict_id; e.b = 0; return false; }
        for (;;) {                                                   // Block::parse block.rs:43-72
            if (n - pos < 3) { e.code = ZSB_E_NOT_ENOUGH_BYTES; e.a = 3; e.b = n - pos; return false; }
            const uint32_t v = (uint32_t)zsb_rd_le(src + pos, 3); pos += 3;
            zsb_block b;
            b.src_off = pos; b.size = v >> 3; b.frame = frame_index; b.type = (uint8_t)((v >> 1) & 3); b.last = (uint8_t)(v & 1);
            for (int i = 0; i < 6; i++) b.pad[i] = 0;
            if (b.type == 3) { e.code = ZSB_E_RESERVED_BLOCK; e.a = e.b = 0; return false; }
            if (b.type == ZSB_BT_RLE) {
                if (n - pos < 1) { e.code = ZSB_E_NOT_ENOUGH_BYTES; e.a = 1; e.b = 0; return false; }
                pos += 1;
            } else {
                if (b.size == 0 && quirks) { e.code = ZSB_E_EMPTY_SLICE; e.a = e.b = 0; return false; }   // slice(0), SURVEY Q2
                if (n - pos < b.size) { e.code = ZSB_E_NOT_ENOUGH_BYTES; e.a = b.size; e.b = n - pos; return false; }
                pos += b.size;
            }
            emit(b); n_emitted++;
            if (b.last) break;
        }
        if (f.has_checksum) {
            if (n - pos < 4) { e.code = ZSB_E_MISSING_CHECKSUM; e.a = 4; e.b = n - pos; return false; }
            f.stored_checksum = (uint32_t)zsb_rd_le(src + pos, 4); pos += 4;
        }
    } else if ((magic ^ ZSB_MAGIC_SKIP) <= 0x0Fu) {
        f.kind = 1;
        ZSB_WALK_NEED(4);
        const uint32_t len = (uint32_t)zsb_rd_le(src + pos, 4); pos += 4;
        if (len == 0 && quirks) { e.code = ZSB_E_EMPTY_SLICE; e.a = e.b = 0; return false; }
        ZSB_WALK_NEED(len);
        zsb_block b;
        b.src_off = pos; b.size = len; b.frame = frame_index; b.type = ZSB_BT_SKIPPABLE; b.last = 1;
        for (int i = 0; i < 6; i++) b.pad[i] = 0;
        emit(b); n_emitted++;
        pos += len;
    } else { e.code = ZSB_E_UNRECOGNIZED_MAGIC; e.a = magic; e.b = 0; return false; }
    end = pos;
    return true;
#undef ZSB_WALK_NEED
}
