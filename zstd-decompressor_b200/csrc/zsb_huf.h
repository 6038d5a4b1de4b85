// zsb_huf.h -- Huffman tree description -> flat decoding LUT -> stream decode, one lane per stream.
//
//  huf_read_weights  == HuffmanDecoder::parse / parse_direct / parse_fse (huffman.rs:80-130)
//                       with AlternatingDecoder (alternating.rs:8-69)
//  huf_build_lut     == HuffmanDecoder::from_weights / from_number_of_bits (huffman.rs:161-203):
//                       the reference builds a boxed binary tree and walks it one bit at a time
//                       (huffman.rs:205-218); the same canonical code (longest codes first, ascending
//                       symbol within a length, counting up from 0) is laid out here as a
//                       2^maxbits-entry LUT {symbol, nbits} indexed by the next maxbits bits.
//  huf_decode_stream == the per-stream loop of LiteralsSection::decode (literals.rs:70-81)
#pragma once
#include "zsb_fse.h"
#include "zsb_stream.h"

#define ZSB_HUF_WEIGHT_SYMS 16   // weights 0..15 can be described (legal ones are <= 11)

// Scratch needed by huf_read_weights: FSE table for the weights (<= 512 cells) and counts (16).
// weights[]: up to 256 entries; returns the number of explicit weights in nw (last one is implied).
// desc = first byte of the description (header byte); limit = bytes available from desc.
ZSB_HDN int huf_read_weights(const uint8_t *desc, uint64_t limit, uint8_t *weights, int ws, int &nw, uint32_t &desc_len,
                             uint32_t *ftbl, int fts, int16_t *cnt, int cs, uint64_t src_end_from_desc, bool quirks) {
    if (limit < 1) return ZSB_E_NOT_ENOUGH_BYTES;
    uint32_t hb = desc[0];
    if (hb >= 128) {                                                     // parse_direct huffman.rs:92-106
        int n = (int)hb - 127;
        uint32_t nb = (uint32_t)(n + 1) / 2;
        if (limit < 1 + (uint64_t)nb) return ZSB_E_NOT_ENOUGH_BYTES;
        for (int i = 0; i < n; i++) {
            uint32_t b = desc[1 + (i >> 1)];
            weights[i * ws] = (uint8_t)((i & 1) ? (b & 15) : (b >> 4));  // high nibble first
        }
        nw = n; desc_len = 1 + nb;
        return ZSB_OK;
    }
    // parse_fse huffman.rs:108-130
    if (hb == 0) return quirks ? ZSB_E_EMPTY_SLICE : ZSB_E_CORRUPT;
    if (limit < 1 + (uint64_t)hb) return ZSB_E_NOT_ENOUGH_BYTES;
    FwdBits f; fwd_init(f, desc + 1, hb);
    int al, nsym;
    int rc = fse_read_ncount(f, cnt, cs, ZSB_HUF_WEIGHT_SYMS, al, nsym);
    if (rc) return rc;
    rc = fse_build_table(cnt, cs, nsym, al, ftbl, fts, 3);
    if (rc) return rc;
    uint32_t br = fwd_bytes_read(f);
    BackWin b;
    rc = back_init(b, desc, 1 + br, 1 + hb, src_end_from_desc);          // BackwardBitParser::new(&data[bytes_read..])
    if (rc) return rc;
    // AlternatingDecoder::initialize: first then second state (alternating.rs:28-34)
    uint32_t s1 = back_take(b, (uint32_t)al);
    uint32_t s2 = back_take(b, (uint32_t)al);
    if (b.rem < 0) return ZSB_E_NOT_ENOUGH_BITS;
    int n = 0;
    // while decoder.expected_bits() <= bitstream.len() { push(symbol); update_bits } (huffman.rs:121-124)
    for (;;) {
        uint32_t e1 = ftbl[s1 * fts];
        if ((int64_t)ZSB_CELL_NB(e1) > b.rem) break;
        if (n >= 255) return ZSB_E_CORRUPT;
        weights[n * ws] = (uint8_t)ZSB_CELL_CODE(e1); n++;
        s1 = ZSB_CELL_BASE(e1) + back_take(b, ZSB_CELL_NB(e1));
        uint32_t t = s1; s1 = s2; s2 = t;                                 // the other state is next
    }
    if (n > 254) return ZSB_E_CORRUPT;
    weights[n * ws] = (uint8_t)ZSB_CELL_CODE(ftbl[s1 * fts]); n++;        // flush both pending symbols (huffman.rs:125-126)
    weights[n * ws] = (uint8_t)ZSB_CELL_CODE(ftbl[s2 * fts]); n++;
    nw = n; desc_len = 1 + hb;
    return ZSB_OK;
}

// weights -> LUT.  lut: 1<<maxbits uint16 entries: symbol | nbits << 8.  rank[16] scratch (strided).
// lens_out (optional, 256 entries, stride 1): code length per symbol for the stage-level API.
// quirks (ZSB_REFERENCE_QUIRKS): where the reference's arithmetic (from_weights huffman.rs:177-203: the implied weight from
// `(2^maxbits - sum) as u8`, no completeness check) builds a tree without panicking, that very tree is laid out -- it may be
// incomplete (cells left ZSB_HUF_ABSENT: the reference panics when a stream reaches such a node) or drop symbols that find no room
// (insert returns false, :161-175); where the reference panics, and always without quirks, RFC 8878 4.2.1.1 decides.
#define ZSB_HUF_ABSENT 0xFFFFu
// *incomplete (optional): the table has ZSB_HUF_ABSENT cells or dropped symbols: only huf_decode_block_ref may read it.
// The arithmetic of huf_build_lut up to the first LUT cell of every weight class (rank[w]); appends the implied weight (n = nw + 1).
struct HufPlan { int mb, n; bool loose; };
ZSB_HDN int huf_lut_plan(uint8_t *weights, int ws, int nw, uint32_t *rank, int rs, HufPlan &P, bool quirks, bool *incomplete) {
    if (incomplete) *incomplete = false;
    uint32_t sum = 0, wmax = 0;
    for (int w = 0; w <= ZSB_HUF_MAX_BITS + 1; w++) rank[w * rs] = 0;     // (first: symbols per weight class, counted in the same pass)
    for (int i = 0; i < nw; i++) {
        uint32_t w = weights[i * ws];
        if (w > ZSB_HUF_MAX_BITS + 1) return ZSB_E_CORRUPT;
        if (w) sum += 1u << (w - 1);
        wmax = w > wmax ? w : wmax;
        rank[w * rs] += 1;
    }
    if (sum == 0) return ZSB_E_CORRUPT;                                   // reference: discrete_log2(0) panics (huffman.rs:184)
    if (nw >= 256) return ZSB_E_CORRUPT;
    int mb = zsb_flog2(sum) + 1;                                          // RFC 8878 4.2.1.1 (reference is off by one when sum is 2^k)
    uint32_t rest = (1u << mb) - sum;
    uint32_t lastw = 0;
    bool loose = false;                                                   // the reference's tree, not necessarily a complete code
    if (quirks) {
        const uint32_t p = (uint32_t)zsb_flog2(sum), pu = p + (((1u << p) < sum) ? 1u : 0u);
        const uint32_t r8 = ((1u << pu) - sum) & 0xFFu;                  // `as u8` huffman.rs:190
        bool panics = r8 == 0;                                            // discrete_log2(0)
        const uint32_t lw = panics ? 0u : (uint32_t)zsb_flog2(r8) + 1u;
        if (wmax > pu + 1 || lw > pu + 1) panics = true;                 // u8 underflow of a width
        if (!panics) {
            if (pu > ZSB_HUF_MAX_BITS || pu == 0) return ZSB_E_CORRUPT;   // deeper than this library's table / a root that is a symbol (the reference never returns)
            if (wmax == pu + 1 || lw == pu + 1) return ZSB_E_CORRUPT;     // width 0: the same
            mb = (int)pu; lastw = lw; loose = true;
        }
    }
    if (!loose) {
        if (mb > ZSB_HUF_MAX_BITS) return ZSB_E_CORRUPT;
        if (rest & (rest - 1)) return ZSB_E_CORRUPT;                      // implied weight must be a power of two
        lastw = (uint32_t)zsb_flog2(rest) + 1;
    }
    weights[nw * ws] = (uint8_t)lastw;
    int n = nw + 1;
    // rank[w] = first LUT cell of weight class w: lowest weights (longest codes) first (huffman.rs:161-175)
    rank[lastw * rs] += 1;
    uint32_t start = 0;
    for (int w = 1; w <= mb; w++) { uint32_t c = rank[w * rs]; rank[w * rs] = start; start += c << (w - 1); }
    if (!loose && start != (1u << mb)) return ZSB_E_CORRUPT;
    if (incomplete) *incomplete = start != (1u << mb);
    P.mb = mb; P.n = n; P.loose = loose;
    return ZSB_OK;
}
ZSB_HDN int huf_build_lut(uint8_t *weights, int ws, int nw, uint16_t *lut, uint32_t *rank, int rs, int &maxbits, uint8_t *lens_out, bool quirks = false,
                          bool *incomplete = nullptr) {
    HufPlan P;
    const int rc = huf_lut_plan(weights, ws, nw, rank, rs, P, quirks, incomplete);
    if (rc) return rc;
    const int mb = P.mb, n = P.n;
    const bool loose = P.loose;
    if (lens_out) for (int i = 0; i < 256; i++) lens_out[i] = 0;
    if (loose) for (uint32_t k = 0; k < (1u << mb); k++) lut[k] = ZSB_HUF_ABSENT;
    for (int i = 0; i < n; i++) {
        uint32_t w = weights[i * ws];
        if (!w) continue;
        uint32_t nbits = (uint32_t)mb + 1 - w, len = 1u << (w - 1), at = rank[w * rs];
        rank[w * rs] = at + len;
        if (at + len > (1u << mb)) continue;                              // (quirks) no room left: the reference's insert returns false and the symbol is dropped
        uint16_t cell = (uint16_t)(i | (nbits << 8));
        for (uint32_t k = 0; k < len; k++) lut[at + k] = cell;
        if (lens_out) lens_out[i] = (uint8_t)nbits;
    }
    maxbits = mb;
    return ZSB_OK;
}

// Decode one backward stream src[start,end) into out[0..expect).  The reference decodes "until the
// reader is empty" and ignores Regenerated_Size (literals.rs:55,78-80); a symbol cut short is
// NotEnoughBits.  Here the stream must additionally regenerate exactly `expect` symbols (RFC 8878
// 3.1.1.3.1.6), otherwise ZSB_E_CORRUPT -- identical on every valid stream.
ZSB_HDN int huf_decode_stream(const uint8_t *src, uint64_t start, uint64_t end, uint64_t src_end,
                              const uint16_t *lut, int maxbits, uint8_t *out, uint32_t expect) {
    BackWin b;
    int rc = back_init(b, src, start, end, src_end);
    if (rc) return rc;
    uint32_t n = 0;
    const uint32_t sh = 64u - (uint32_t)maxbits;
    while (b.rem > 0) {
        back_refill(b);
        // up to 5 symbols per refill: 5 * 11 = 55 <= 64 valid bits
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (int k = 0; k < 5 && b.rem > 0; k++) {
            uint32_t cell = lut[(uint32_t)(b.hi >> sh)];
            uint32_t nb = cell >> 8;
            if ((int64_t)nb > b.rem) return ZSB_E_NOT_ENOUGH_BITS;
            if (n >= expect) return ZSB_E_CORRUPT;
            out[n++] = (uint8_t)cell;
            back_consume(b, nb);
        }
    }
    return n == expect ? ZSB_OK : ZSB_E_CORRUPT;
}


// ZSB_REFERENCE_QUIRKS: the literals of one block as the reference decodes them (LiteralsSection::decode literals.rs:70-81): stream
// after stream in jump-table order, stopping at the first zero-sized entry, each decoded UNTIL ITS BITS RUN OUT (a symbol cut short is
// NotEnoughBits, a stream whose last byte is 0 NullByte), the symbols of all streams concatenated; Regenerated_Size plays no part.
// out == nullptr: count only.  n_out = symbols regenerated.  cap: room at out.
ZSB_HDN int huf_decode_block_ref(const uint8_t *src, uint64_t src_end, uint64_t lit_src, const uint32_t *stream_size, const uint16_t *lut, int maxbits,
                                 uint8_t *out, uint32_t cap, uint32_t &n_out) {
    uint64_t start = lit_src;
    uint32_t n = 0;
    const uint32_t sh = 64u - (uint32_t)maxbits;
    for (int s = 0; s < 4; s++) {
        const uint32_t sz = stream_size[s];
        if (sz == 0) break;
        BackWin b;
        const int rc = back_init(b, src, start, start + sz, src_end);
        if (rc) return rc;
        while (b.rem > 0) {
            back_refill(b);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
            for (int k = 0; k < 5 && b.rem > 0; k++) {
                const uint32_t cell = lut[(uint32_t)(b.hi >> sh)];
                if (cell == ZSB_HUF_ABSENT) return ZSB_E_CORRUPT;         // an Absent node of an incomplete tree: the reference panics (huffman.rs:216)
                const uint32_t nb = cell >> 8;
                if ((int64_t)nb > b.rem) return ZSB_E_NOT_ENOUGH_BITS;
                if (out) { if (n >= cap) return ZSB_E_BLOCK_TOO_LARGE; out[n] = (uint8_t)cell; }
                n++;
                back_consume(b, nb);
            }
        }
        start += sz;
    }
    n_out = n;
    return ZSB_OK;
}


// ---- fast stream decode ------------------------------------------------------------------------------
// The same chain as huf_decode_stream with everything that is not on it removed: the window is reloaded once
// per four symbols (4 x 11 bits fit), four symbols leave as one aligned 32-bit store, and the stream is decoded
// for exactly `expect` symbols and must then be exactly empty.  Anything else (empty stream, missing end mark,
// over-read, bits left over, a stream too close to the buffer start) returns ZSB_NEEDS_SLOW and the caller runs
// huf_decode_stream, which reports what the reference reports.  On the GPU the stream is staged through a
// shared-memory ring (ring_sa, zsb_stream.h); the host build reads it in place.
ZSB_HDN int huf_fast_stream(const uint8_t *src, uint64_t start, uint64_t end, const uint16_t *lut, int maxbits, uint8_t *out, uint32_t expect,
                            uint32_t ring_sa) {
    if (end <= start || start < 16 || end - start > (1u << 24)) return ZSB_NEEDS_SLOW;
    const uint32_t lastb = src[end - 1];
    if (lastb == 0) return ZSB_NEEDS_SLOW;
    FastWin F;
#if defined(__CUDA_ARCH__)
    StreamRing R;
    R.sa = ring_sa;
    R.pl = (const uint8_t *)(((uintptr_t)(src + start) - 16) & ~(uintptr_t)63);
    const uint32_t d0 = (uint32_t)((src + start) - R.pl);                     // 16 .. 79
    int32_t top = (int32_t)((d0 + (uint32_t)(end - start) - 1) * 8) + zsb_flog2(lastb);
    const int32_t startbit = (int32_t)(d0 * 8);
    sr_init<6>(R, top);
#define HUF_LOAD(t_) sr_load<6>(R, F, t_)
#else
    (void)ring_sa;
    const uint32_t mis = (uint32_t)((uintptr_t)src & 7);
    const uint64_t w0 = ((start + mis) & ~7ull) - 8;
    const uint8_t *pw = src - mis + w0;
    int32_t top = (int32_t)((end - 1 + mis - w0) * 8) + zsb_flog2(lastb);
    const int32_t startbit = (int32_t)((start + mis - w0) * 8);
#define HUF_LOAD(t_) fast_win_load(F, pw, t_)
#endif
    const uint32_t sh = 64u - (uint32_t)maxbits;
    uint32_t n = 0;
    // single symbols until the output is 4-byte aligned
    while (n < expect && (((uintptr_t)(out + n)) & 3)) {
        HUF_LOAD(top);
        const uint32_t cell = lut[(uint32_t)(fast_win_get(F) >> sh)];
        out[n++] = (uint8_t)cell; top -= (int32_t)(cell >> 8);
    }
#if defined(__CUDA_ARCH__)
    // eight steps of four symbols consume at most 352 bits, less than a 64-byte line: the ring is topped up once per eight
    // steps, by all lanes of the warp in the same pass
    for (; n + 32 <= expect; n += 32) {
        sr_check<6>(R, top);
#pragma unroll 2
        for (uint32_t k = 0; k < 32; k += 4) {
            sr_load_nocheck<6>(R, F, top);
            uint64_t W = fast_win_get(F);
            const uint32_t c0 = lut[(uint32_t)(W >> sh)]; W <<= (c0 >> 8);
            const uint32_t c1 = lut[(uint32_t)(W >> sh)]; W <<= (c1 >> 8);
            const uint32_t c2 = lut[(uint32_t)(W >> sh)]; W <<= (c2 >> 8);
            const uint32_t c3 = lut[(uint32_t)(W >> sh)];
            top -= (int32_t)((c0 >> 8) + (c1 >> 8) + (c2 >> 8) + (c3 >> 8));
            *reinterpret_cast<uint32_t *>(out + n + k) = (c0 & 0xFFu) | (c1 & 0xFFu) << 8 | (c2 & 0xFFu) << 16 | c3 << 24;
        }
    }
#endif
    for (; n + 4 <= expect; n += 4) {
        HUF_LOAD(top);
        uint64_t W = fast_win_get(F);
        const uint32_t c0 = lut[(uint32_t)(W >> sh)]; W <<= (c0 >> 8);
        const uint32_t c1 = lut[(uint32_t)(W >> sh)]; W <<= (c1 >> 8);
        const uint32_t c2 = lut[(uint32_t)(W >> sh)]; W <<= (c2 >> 8);
        const uint32_t c3 = lut[(uint32_t)(W >> sh)];
        top -= (int32_t)((c0 >> 8) + (c1 >> 8) + (c2 >> 8) + (c3 >> 8));
        *reinterpret_cast<uint32_t *>(out + n) = (c0 & 0xFFu) | (c1 & 0xFFu) << 8 | (c2 & 0xFFu) << 16 | c3 << 24;
    }
    while (n < expect) {
        HUF_LOAD(top);
        const uint32_t cell = lut[(uint32_t)(fast_win_get(F) >> sh)];
        out[n++] = (uint8_t)cell; top -= (int32_t)(cell >> 8);
    }
#undef HUF_LOAD
    return top == startbit ? ZSB_OK : ZSB_NEEDS_SLOW;
}
