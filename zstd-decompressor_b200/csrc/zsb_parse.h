// zsb_parse.h -- section-header parsing of one compressed block (one lane per block) and the two
// per-frame chain passes that carry state from block to block.
//
//  parse_block    == LiteralsSection::parse / parse_header (literals.rs:88-206),
//                    Sequences::parse / parse_num_sequences / parse_symbol_compression
//                    (sequences.rs:52-143) -- extents and modes only; tables are built by the
//                    entropy kernels.
//  chain_frame    == the state DecodingContext carries between blocks at decode time:
//                    huffman_decoder (literals.rs:59-66) and the three repeat modes
//                    (sequences.rs:147-187,232-234), resolved to the block that defined them.
//  plan_frame     == output placement (Vec::append order, block.rs:76-86) and the repeat-offset
//                    history (decoding_context.rs:40,50-75) carried across blocks.
#pragma once
#include "zsb_fse.h"
#include "zsb_seq.h"

#define ZSB_FAIL(w, code, a, b) do { (w).status = (code); (w).err_a = (uint32_t)(a); (w).err_b = (uint32_t)(b); return; } while (0)

ZSB_HDN void parse_block(const uint8_t *src, const zsb_block &blk, ZsbBlockWork &w, uint32_t flags) {
    const bool quirks = (flags & ZSB_REFERENCE_QUIRKS) != 0;
    w.status = ZSB_OK; w.lit_status = ZSB_OK; w.err_a = w.err_b = 0; w.seq_rem0 = 0;
    w.nseq = 0; w.lit_type = ZSB_LT_NONE; w.lit_regen = 0; w.n_streams = 0; w.raw_modes = 0; w.parse_stage = 0; w.lit_inexact = 0; w.fused = 0;
    w.lit_used = 0; w.out_off = 0; w.lit_buf = 0; w.seq_buf = 0;
    w.mode[0] = w.mode[1] = w.mode[2] = ZSB_M_REPEAT;
    w.rep_out[0] = ZSB_OFF_SYM | (0u << 25); w.rep_out[1] = ZSB_OFF_SYM | (1u << 25); w.rep_out[2] = ZSB_OFF_SYM | (2u << 25);
    if (blk.type != ZSB_BT_COMPRESSED) { w.out_size = blk.size; return; }
    w.out_size = 0;
    uint64_t p = blk.src_off;
    const uint64_t end = p + blk.size;
    // ---- literals section header (literals.rs:135-206)
    if (p >= end) ZSB_FAIL(w, quirks ? ZSB_E_EMPTY_SLICE : ZSB_E_NOT_ENOUGH_BYTES, 1, 0);
    const uint32_t h = src[p++];
    const uint32_t lt = h & 3, sf = (h >> 2) & 3;
    uint32_t regen, csize = 0, nstreams = 1;
    if (lt <= ZSB_LT_RLE) {
        if (sf == 0 || sf == 2) regen = h >> 3;
        else if (sf == 1) { if (end - p < 1) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, 1, 0); regen = (h >> 4) + ((uint32_t)src[p] << 4); p += 1; }
        else {
            if (end - p < 2) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, 1, 0);
            regen = (h >> 4) + ((uint32_t)src[p] << 4) + ((uint32_t)src[p + 1] << 12); p += 2;
        }
    } else {
        const uint32_t extra = sf <= 1 ? 2 : sf == 2 ? 3 : 4;
        if (end - p < extra) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, extra, end - p);
        uint32_t v = 0;
        for (uint32_t i = 0; i < extra; i++) v |= (uint32_t)src[p + i] << (8 * i);
        p += extra;
        const uint32_t rb = sf <= 1 ? 6 : sf == 2 ? 10 : 14, cb = sf <= 1 ? 10 : sf == 2 ? 14 : 18;
        regen = (h >> 4) + ((v & ((1u << rb) - 1u)) << 4);
        csize = (v >> rb) & ((1u << cb) - 1u);
        nstreams = sf == 0 ? 1 : 4;
    }
    w.lit_type = (uint8_t)lt; w.lit_regen = regen; w.n_streams = (uint8_t)nstreams;
    if (regen > ZSB_BLOCK_MAX) ZSB_FAIL(w, ZSB_E_BLOCK_TOO_LARGE, regen, 0);
    if (lt == ZSB_LT_RAW) {                                            // literals.rs:92-94
        if (regen == 0 && quirks) ZSB_FAIL(w, ZSB_E_EMPTY_SLICE, 0, 0);
        if (end - p < regen) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, regen, end - p);
        w.lit_src = p; p += regen;
    } else if (lt == ZSB_LT_RLE) {                                     // literals.rs:95-98
        if (end - p < 1) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, 1, 0);
        w.lit_src = p; p += 1;
    } else {                                                           // literals.rs:99-131
        if (csize == 0) ZSB_FAIL(w, quirks ? ZSB_E_EMPTY_SLICE : ZSB_E_CORRUPT, 0, 0);
        if (end - p < csize) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, csize, end - p);
        const uint64_t lend = p + csize;
        uint64_t q = p;
        w.huf_desc = 0; w.huf_desc_end = 0;
        if (lt == ZSB_LT_COMPRESSED) {
            const uint32_t hb = src[q];
            const uint32_t dl = hb < 128 ? 1 + hb : 1 + (hb - 127 + 1) / 2;
            if (hb == 0) ZSB_FAIL(w, quirks ? ZSB_E_EMPTY_SLICE : ZSB_E_CORRUPT, 0, 0);
            if (lend - q < dl) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, dl - 1, lend - q - 1);
            w.huf_desc = q; w.huf_desc_end = lend; q += dl;
        }
        const uint64_t total = lend - q;
        if (nstreams == 4) {
            if (total < 6) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, 2, total & 1);
            const uint32_t s1 = src[q] | (src[q + 1] << 8), s2 = src[q + 2] | (src[q + 3] << 8), s3 = src[q + 4] | (src[q + 5] << 8);
            if ((uint64_t)s1 + s2 + s3 > total - 6) ZSB_FAIL(w, ZSB_E_STREAMS_TOO_BIG, 0, 0);   // literals.rs:115-117
            const uint64_t s4 = total - 6 - s1 - s2 - s3;
            if (quirks) {
                // the reference keeps `s4 as u16` (literals.rs:119-120), stops decoding at the first zero entry (:71-73) and never looks at
                // Regenerated_Size (:55): anything but the regular shape is decoded its way (huf_decode_block_ref)
                if (total == 6) ZSB_FAIL(w, ZSB_E_EMPTY_SLICE, 0, 0);                              // slice(0) literals.rs:130
                w.stream_size[0] = s1; w.stream_size[1] = s2; w.stream_size[2] = s3; w.stream_size[3] = (uint32_t)(s4 & 0xFFFFu);
                if (s1 == 0 || s2 == 0 || s3 == 0 || s4 == 0 || s4 > 0xFFFFu || 3 * ((regen + 3) / 4) > regen) w.lit_inexact = 1;
            } else {
                w.stream_size[0] = s1; w.stream_size[1] = s2; w.stream_size[2] = s3; w.stream_size[3] = (uint32_t)s4;
                if (s1 == 0 || s2 == 0 || s3 == 0 || s4 == 0) ZSB_FAIL(w, ZSB_E_CORRUPT, 0, 0);
                if (3 * ((regen + 3) / 4) > regen) ZSB_FAIL(w, ZSB_E_CORRUPT, 0, 0);   // streams 1-3 regenerate (regen+3)/4 symbols each
            }
            w.lit_src = q + 6;
        } else {
            if (total == 0) ZSB_FAIL(w, quirks ? ZSB_E_EMPTY_SLICE : ZSB_E_CORRUPT, 0, 0);
            w.stream_size[0] = quirks ? (uint32_t)(total & 0xFFFFu) : (uint32_t)total; w.stream_size[1] = w.stream_size[2] = w.stream_size[3] = 0;   // `len as u16` literals.rs:122
            if (quirks && total > 0xFFFFu) w.lit_inexact = 1;
            w.lit_src = q;
        }
        p = lend;
    }
    // ---- sequences section header (sequences.rs:52-143)
    w.parse_stage = 1;
    if (p >= end) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, 1, 0);
    uint32_t b0 = src[p++], nseq;
    if (b0 < 128) nseq = b0;
    else if (b0 < 255) { if (p >= end) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, 1, 0); nseq = ((b0 - 128) << 8) + src[p++]; }
    else {
        if (end - p < 2) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, 1, 0);
        nseq = (uint32_t)src[p] + ((uint32_t)src[p + 1] << 8) + (quirks ? 0x7Fu : 0x7F00u); p += 2;   // SURVEY Q3 (sequences.rs:84)
    }
    w.nseq = nseq;
    w.tbl_end = end;
    if (nseq == 0) {
        // RFC 8878: the section ends here and the block is literals only.  The reference cannot decode
        // such a block; with ZSB_REFERENCE_QUIRKS the chain pass assigns its error.
        if (!quirks && p != end) ZSB_FAIL(w, ZSB_E_CORRUPT, 0, 0);
        w.out_size = regen;
        return;
    }
    if (p >= end) ZSB_FAIL(w, quirks ? ZSB_E_EMPTY_SLICE : ZSB_E_NOT_ENOUGH_BYTES, 1, 0);   // input.slice(1) sequences.rs:94
    const uint32_t mb = src[p++];
    if (mb & 3) ZSB_FAIL(w, ZSB_E_SEQ_RESERVED, 0, 0);
    w.raw_modes = (uint8_t)mb;
    w.parse_stage = 2;
    for (int t = 0; t < 3; t++) {                                      // LL, OF, ML in this order (sequences.rs:116-141)
        const uint32_t m = (mb >> (6 - 2 * t)) & 3;
        w.mode[t] = (uint8_t)m;
        if (m == ZSB_M_RLE) {
            if (p >= end) ZSB_FAIL(w, ZSB_E_NOT_ENOUGH_BYTES, 1, 0);
            w.rle_sym[t] = src[p++];
        } else if (m == ZSB_M_FSE) {
            if (p >= end) ZSB_FAIL(w, quirks ? ZSB_E_EMPTY_SLICE : ZSB_E_NOT_ENOUGH_BITS, 0, 0);
            // walk the description to find its length; the counts themselves are re-read by the decoder
            FwdBits f; fwd_init(f, src + p, end - p);
            int16_t dummy[1]; int al = 0, ns = 0;
            int rc = fse_read_ncount(f, dummy, 0, 0x7FFFFFFF, al, ns);   // stride 0: counts are discarded
            if (rc) ZSB_FAIL(w, rc, al, 0);
            w.tbl_desc[t] = p; p += fwd_bytes_read(f);
        }
        w.parse_stage = (uint8_t)(3 + t);
    }
    if (p >= end) ZSB_FAIL(w, quirks ? ZSB_E_EMPTY_SLICE : ZSB_E_CORRUPT, 0, 0);   // empty bitstream: slice(0) sequences.rs:72
    w.bs_off = p; w.bs_len = (uint32_t)(end - p);
    // (a bitstream whose last byte is 0 is NullByte, but only when the block is DECODED -- BackwardBitParser::new runs in
    //  Sequences::decode, sequences.rs:211, behind the literals of the block: the sequence stage reports it, not this pass)
}

// Carries the Huffman table and the three table modes from block to block of one frame.
// Returns the frame status (first failing block's status).
// (chain_frame_from: the same walk from block k0 on with the state the blocks before it left -- the warp-cooperative form in k_plan1 hands a frame
//  over to it at the first block that is not plain)
ZSB_HDN int chain_frame_from(const zsb_frame &fr, const zsb_block *blocks, ZsbBlockWork *work, uint32_t flags, uint32_t &err_a, uint32_t &err_b,
                             uint32_t k0, int64_t huf_src, int64_t t0, int64_t t1, int64_t t2) {
    const bool quirks = (flags & ZSB_REFERENCE_QUIRKS) != 0;
    int64_t tsrc[3] = {t0, t1, t2};
    // Frame::parse reads the sections of every block before anything is decoded (frame.rs:210-223): the first block that failed to PARSE fails the
    // frame, whatever the blocks before it would do when they are decoded
    for (uint32_t k = k0; k < fr.n_blocks; k++) {
        const uint32_t bi = fr.first_block + k;
        if (blocks[bi].type != ZSB_BT_COMPRESSED || work[bi].status == ZSB_OK) continue;
        err_a = work[bi].err_a; err_b = work[bi].err_b;
        for (uint32_t j = k + 1; j < fr.n_blocks; j++)
            if (work[fr.first_block + j].status == ZSB_OK) work[fr.first_block + j].status = ZSB_E_PREVIOUS_FRAME;
        return work[bi].status;
    }
    // What follows are errors of Block::decode (the Huffman tree of literals.rs:59-66, the tables of sequences.rs:147-187), which runs block after block:
    // they stay with their block (k_plan2 reports the first failing block in order, so an earlier block whose entropy decoding fails comes first),
    // the blocks behind are never reached.  A block that fails in the sequence stage still has its literals decoded first (block.rs:83-85): k_plan1
    // lists it for k_huf (ZSB_CHAIN_SEQ_ERROR).
    for (uint32_t k = k0; k < fr.n_blocks; k++) {
        const uint32_t bi = fr.first_block + k;
        ZsbBlockWork &w = work[bi];
        if (blocks[bi].type != ZSB_BT_COMPRESSED) continue;
        if (w.status == ZSB_OK) {
            if (w.lit_type == ZSB_LT_COMPRESSED) huf_src = bi;
            else if (w.lit_type == ZSB_LT_TREELESS) {
                if (huf_src < 0) { w.status = ZSB_E_HUFFMAN_MISSING; }                       // literals.rs:63-66
                else { w.huf_desc = work[huf_src].huf_desc; w.huf_desc_end = work[huf_src].huf_desc_end; }
            }
        }
        if (w.status == ZSB_OK) {
            if (w.nseq == 0) {
                if (quirks) w.status = (tsrc[0] < 0) ? ZSB_E_NO_PREVIOUS_DECODER : ZSB_E_EMPTY_INPUT_DATA;
            } else {
                for (int t = 0; t < 3 && w.status == ZSB_OK; t++) {
                    if (w.mode[t] == ZSB_M_REPEAT) {
                        if (tsrc[t] < 0) { w.status = ZSB_E_NO_PREVIOUS_DECODER; break; }     // sequences.rs:165-171
                        const ZsbBlockWork &s = work[tsrc[t]];
                        w.mode[t] = s.mode[t]; w.rle_sym[t] = s.rle_sym[t]; w.tbl_desc[t] = s.tbl_desc[t];
                    }
                }
                if (w.status == ZSB_OK) {
                    // (an inherited FSE description lies in an earlier block: it was length-checked there,
                    //  so reading it again with this block's larger limit gives the same counts)
                    tsrc[0] = tsrc[1] = tsrc[2] = bi;                                         // sequences.rs:232-234
                }
            }
        }
        if (w.status != ZSB_OK) {
            for (uint32_t j = k + 1; j < fr.n_blocks; j++)                // not reached: the frame fails here at the latest
                if (work[fr.first_block + j].status == ZSB_OK) work[fr.first_block + j].status = ZSB_E_PREVIOUS_FRAME;
            return ZSB_OK;
        }
    }
    return ZSB_OK;
}
// statuses chain_frame leaves on a block whose sequence stage cannot start: its literals are decoded all the same
#define ZSB_CHAIN_SEQ_ERROR(st) ((st) == ZSB_E_NO_PREVIOUS_DECODER || (st) == ZSB_E_EMPTY_INPUT_DATA)

ZSB_HDN int chain_frame(const zsb_frame &fr, const zsb_block *blocks, ZsbBlockWork *work, uint32_t flags, uint32_t &err_a, uint32_t &err_b) {
    return chain_frame_from(fr, blocks, work, flags, err_a, err_b, 0, -1, -1, -1, -1);
}

// Output placement and repeat-offset history of one frame, after the entropy stage.
// Returns the frame status; total = regenerated size of the frame.
ZSB_HDN int plan_frame(const zsb_frame &fr, const zsb_block *blocks, ZsbBlockWork *work, uint64_t &total, uint32_t &err_a, uint32_t &err_b) {
    uint32_t rep[3] = {1, 4, 8};                                       // decoding_context.rs:40
    uint64_t pos = 0;
    for (uint32_t k = 0; k < fr.n_blocks; k++) {
        const uint32_t bi = fr.first_block + k;
        ZsbBlockWork &w = work[bi];
        if (w.lit_status != ZSB_OK) { total = pos; return w.lit_status; }       // Block::decode: literals first (block.rs:83-85)
        if (w.status != ZSB_OK) { total = pos; err_a = w.err_a; err_b = w.err_b; return w.status; }
        w.out_off = pos;
        w.rep_in[0] = rep[0]; w.rep_in[1] = rep[1]; w.rep_in[2] = rep[2];
        if (blocks[bi].type == ZSB_BT_COMPRESSED && w.nseq) {
            uint32_t n0 = seq_real_offset(w.rep_out[0], rep), n1 = seq_real_offset(w.rep_out[1], rep), n2 = seq_real_offset(w.rep_out[2], rep);
            // a zero here means an "offset - 1" reached 0: the executor reports it on the sequence that used it
            rep[0] = n0; rep[1] = n1; rep[2] = n2;
        }
        pos += w.out_size;
    }
    total = pos;
    return ZSB_OK;
}
