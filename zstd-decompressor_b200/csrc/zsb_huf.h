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
// The chain of a stream is "cell -> bits consumed -> next cell"; everything else hangs off it.  The fast path therefore looks up TWO
// symbols at a time where the next ten bits hold two whole codes: a table of 1 024 32-bit cells indexed by the next 10 bits of the stream,
//     cell = s1 | s2 << 8 | l1 << 16 | long << 20 | count << 21 | ltot << 24
//     count 2: two symbols s1 s2 in ltot = l1 + l2 <= 10 bits;  count 1: one symbol s1 in ltot = l1 bits (the next code does not fit; s2 = 0);
//     long (count 1): the ten bits are the prefix of two 11-bit codes: the eleventh bit selects s1 (0) or s2 (1), ltot = l1 = 11
// -- the same symbols as one-symbol lookups, cell by cell (a code is decoded from bits that determine it completely).  Only for
// complete codes (every 10-bit prefix leads somewhere); everything else goes the careful way.  4 KiB like the one-symbol table of
// 2 048 16-bit cells it replaces in k_huf's shared memory.
// The table is built from T1, the first symbol under every 10-bit prefix (1 KiB; for a prefix of 11-bit codes the even child, the odd
// one in `odd`), and the code lengths, which are the weights (len = maxbits + 1 - weight).
#define ZSB_HUF_PAIR_BITS 10
// T1 cells of one symbol: `at` = its first cell in the maxbits-bit table (huf_lut_plan's order), wt > 0 its weight
ZSB_HD void huf_t1_put(uint8_t *t1, uint8_t *odd, int mb, uint32_t sym, uint32_t at, uint32_t wt) {
    uint32_t a, len;
    if (mb > ZSB_HUF_PAIR_BITS) {                                   // mb == 11: two cells of the big table per T1 cell
        if (wt == 1) { if (at & 1u) odd[at >> 1] = (uint8_t)sym; else t1[at >> 1] = (uint8_t)sym; return; }
        a = at >> 1; len = 1u << (wt - 2);
    } else { a = at << (ZSB_HUF_PAIR_BITS - mb); len = (1u << (wt - 1)) << (ZSB_HUF_PAIR_BITS - mb); }
#if defined(__CUDA_ARCH__)
    if (len >= 16) {                                                // a is a multiple of len (a complete code): aligned
        const uint32_t v = sym * 0x01010101u;
        for (uint32_t k = 0; k < len; k += 16) *reinterpret_cast<uint4 *>(t1 + a + k) = make_uint4(v, v, v, v);
    } else
#endif
    for (uint32_t k = 0; k < len; k++) t1[a + k] = (uint8_t)sym;
}
ZSB_HD uint32_t huf_pair_cell(uint32_t x, const uint8_t *t1, const uint8_t *odd, const uint8_t *weights, int ws, int mb) {
    // (no branches: the four loads of a cell are dependent, the cells of one lane are meant to overlap)
    const uint32_t s1 = t1[x], l1 = (uint32_t)mb + 1u - weights[s1 * ws];
    const bool lng = l1 > ZSB_HUF_PAIR_BITS;
    const uint32_t rest = (x << (lng ? 0u : l1)) & ((1u << ZSB_HUF_PAIR_BITS) - 1u);
    const uint32_t s2 = t1[rest], l2 = (uint32_t)mb + 1u - weights[s2 * ws];
    const uint32_t od = odd[x & 127u];
    const bool two = !lng && l1 + l2 <= ZSB_HUF_PAIR_BITS;
    const uint32_t b1 = lng ? od : two ? s2 : 0u, lt = two ? l1 + l2 : l1;
    return s1 | b1 << 8 | l1 << 16 | (lng ? 1u : 0u) << 20 | (two ? 2u : 1u) << 21 | lt << 24;
}
// the same table from a finished one-symbol table (host builds: tests/emul); t1 / odd: 1 024 / 512 bytes of scratch
ZSB_HDN void huf_pairs_from_lut(const uint16_t *lut, int mb, const uint8_t *weights, int ws, uint8_t *t1, uint8_t *odd, uint32_t *pair) {
    for (uint32_t x = 0; x < (1u << ZSB_HUF_PAIR_BITS); x++) {
        if (mb > ZSB_HUF_PAIR_BITS) { t1[x] = (uint8_t)lut[2 * x]; if (x < 512) odd[x] = (uint8_t)lut[2 * x + 1]; }
        else t1[x] = (uint8_t)lut[x >> (ZSB_HUF_PAIR_BITS - mb)];
    }
    for (uint32_t x = 0; x < (1u << ZSB_HUF_PAIR_BITS); x++) pair[x] = huf_pair_cell(x, t1, odd, weights, ws, mb);
}

// The stream is decoded for exactly `expect` symbols and must then be exactly empty.  Anything else (empty stream, missing end mark,
// over-read, bits left over, a stream too close to the buffer start) returns ZSB_NEEDS_SLOW and the caller runs
// huf_decode_stream, which reports what the reference reports.  The window is reloaded once per five cells (5 x 11 bits fit), the
// symbols collect in a register and leave as aligned 32-bit stores.  On the GPU the stream is staged through a
// shared-memory ring (ring_sa, zsb_stream.h); the host build reads it in place.
ZSB_HDN int huf_fast_stream(const uint8_t *src, uint64_t start, uint64_t end, const uint32_t *pair, uint8_t *out, uint32_t expect,
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
    const uint32_t ish = 64u - ZSB_HUF_PAIR_BITS;
    uint32_t n = 0;
    // one symbol: the first of a cell
#define HUF_ONE() do { \
        HUF_LOAD(top); \
        const uint64_t W1 = fast_win_get(F); \
        const uint32_t c = pair[(uint32_t)(W1 >> ish)]; \
        uint32_t sy = c & 0xFFu; \
        if (((c >> 20) & 1u) && ((W1 >> (ish - 1)) & 1u)) sy = (c >> 8) & 0xFFu; \
        out[n++] = (uint8_t)sy; top -= (int32_t)((c >> 16) & 15u); } while (0)
    // five cells from one window: 5 .. 10 symbols into `pend` (np8 bits pending, < 32 between windows), whole words out.
    // The cells are looked up first -- that is the chain --, then the cursor moves and NEXT requests the words of the following window, and
    // only then are the symbols of this window put away: that work (two thirds of the instructions) runs in the shadow of the next loads.
    // lb: 8 when the cell is a prefix of two 11-bit codes and the eleventh bit of the window is set (the odd child, byte 1 of the cell)
#define HUF_LOOKUP(c_, h_) const uint32_t h_ = (uint32_t)(W >> 32); const uint32_t c_ = pair[h_ >> (32u - ZSB_HUF_PAIR_BITS)]; W <<= (c_ >> 24)
#define HUF_PUT(c_, h_) do { \
        const uint32_t lg = (c_ >> 17) & 8u, lb = (h_ >> 18) & lg; \
        const uint32_t sy = (c_ >> lb) & (0xFFFFu >> lg); \
        pend |= (uint64_t)sy << np8; np8 += (c_ >> 18) & 0x18u; } while (0)
#define HUF_FLUSH() do { if (np8 >= 32u) { *reinterpret_cast<uint32_t *>(out + n) = (uint32_t)pend; pend >>= 32; n += 4; np8 -= 32u; } } while (0)
#define HUF_WINDOW(NEXT) do { \
        uint64_t W = fast_win_get(F); \
        HUF_LOOKUP(c0, h0); HUF_LOOKUP(c1, h1); HUF_LOOKUP(c2, h2); HUF_LOOKUP(c3, h3); HUF_LOOKUP(c4, h4); \
        top -= (int32_t)((c0 >> 24) + (c1 >> 24) + (c2 >> 24) + (c3 >> 24) + (c4 >> 24)); \
        NEXT; \
        HUF_PUT(c0, h0); HUF_PUT(c1, h1); HUF_FLUSH(); HUF_PUT(c2, h2); HUF_PUT(c3, h3); HUF_FLUSH(); HUF_PUT(c4, h4); HUF_FLUSH(); } while (0)
    // single symbols until the output is 4-byte aligned
    while (n < expect && (((uintptr_t)(out + n)) & 3)) HUF_ONE();
    uint64_t pend = 0; uint32_t np8 = 0;
#if defined(__CUDA_ARCH__)
    // eight windows of five cells consume at most 440 bits, less than a 64-byte line: the ring is topped up once per eight
    // windows, by all lanes of the warp in the same pass
    while (n + (np8 >> 3) + 80 <= expect) {
        sr_check<6>(R, top);
        sr_load_nocheck<6>(R, F, top);
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) HUF_WINDOW(if (k < 7) sr_load_nocheck<6>(R, F, top));
    }
#endif
    while (n + (np8 >> 3) + 10 <= expect) {
        HUF_LOAD(top);
        HUF_WINDOW((void)0);
    }
    while (np8) { out[n++] = (uint8_t)pend; pend >>= 8; np8 -= 8u; }
    while (n < expect) HUF_ONE();
#undef HUF_LOAD
#undef HUF_ONE
#undef HUF_LOOKUP
#undef HUF_PUT
#undef HUF_FLUSH
#undef HUF_WINDOW
    return top == startbit ? ZSB_OK : ZSB_NEEDS_SLOW;
}
