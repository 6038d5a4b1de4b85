// zsb_seqfast.h -- the fast sequence path: the serial FSE chain stripped to its dependency core
// (phase 1, one lane per block) and everything else moved to a lane-per-sequence pass (phase 2).
//
// Why two phases.  Sequences::decode (sequences.rs:191-237) is one dependent chain per block:
// state -> table cell -> number of bits -> bit position -> next state, three states sharing one bit
// cursor.  Nothing else in the loop of SequenceDecoder::decode (decoders/sequence.rs:41-88) -- the
// extra-bit values, the code -> baseline tables, literal/match positions, the repeat-offset history
// of DecodingContext::decode_offset (decoding_context.rs:50-75) -- feeds back into that chain.  So
//
//   phase 1 (seq_fast_phase1)  walks the chain only and emits one 32-bit word per sequence:
//                              the three codes and how many bits the sequence consumed;
//   phase 2 (seq_fast_values, hist_*) turns 32 words at a time into (ll, ml, offset_value) with one
//                              lane per sequence: bit positions by prefix sum, extra bits by direct
//                              extraction, positions by prefix sum, the repeat-offset history as a
//                              prefix "sum" over composable history transforms.
//
// Both phases only ever produce results for streams on which nothing unusual happened.  Any anomaly
// (illegal code, over-read, a literal or output overrun, an offset that reaches zero or exceeds the
// representable range) returns ZSB_NEEDS_SLOW and the block is decoded again by the careful
// seq_decode (zsb_seq.h), which reproduces the reference's error order exactly.
#pragma once
#include "zsb_seq.h"
#include "zsb_stream.h"


// phase-1 word, one byte per field: LL code | OF code | ML code | state bits consumed (LL+ML+OF).
// The code bytes carry two stray high bits (the low bits of the cell's base field): mask with 63.
#define ZSB_W_CL(w) ((w) & 63u)
#define ZSB_W_CO(w) (((w) >> 8) & 63u)
#define ZSB_W_CM(w) (((w) >> 16) & 63u)
#define ZSB_W_NB(w) (((w) >> 24) & 31u)
ZSB_HD uint32_t seq_fast_word(uint32_t eL, uint32_t eO, uint32_t eM, uint32_t nbs) {
    return zsb_prmt(zsb_prmt(eL, eO, 0x0062), zsb_prmt(eM, nbs, 0x0052), 0x5410);      // nbs: a sum of cells, state bits in byte 1
}

// code -> baseline | extra bits << 24 (sequence.rs:98-191)
ZSB_HD uint32_t zsb_ll_entry(uint32_t c) { return zsb_ll_base(c) | (zsb_ll_bits(c) << 24); }
ZSB_HD uint32_t zsb_ml_entry(uint32_t c) { return zsb_ml_base(c) | (zsb_ml_bits(c) << 24); }

// ---- phase 1 ---------------------------------------------------------------------------------------
// Walks the three-state chain of block w and writes words[0..nseq) (stride ws).  rem0 = unread bits
// after the three initial states, i.e. where sequence 0 starts.  Returns ZSB_OK, an initialisation
// error identical to the careful path's (BackwardBitParser::new parsing.rs:200-220, the initial
// states sequence.rs:59-65), or ZSB_NEEDS_SLOW.
//
// Bit positions are 32-bit and relative to `pw`, the aligned 64-bit word just BELOW the word holding
// the first stream byte, so that the 64-bit window ending at any position inside the stream starts at
// a non-negative bit.  The two window words of the next step are requested as soon as the bit count of
// the current one is known and are only combined after the next step's table cells were requested.
// This is the reference form of phase 1 (the CPU build of the differential tests runs it); k_seq in
// zsb_kernels.cu runs the same steps with the states kept as shared-memory addresses, the window fed from a
// cp.async stream ring and the words handed to the phase-2 warps through shared memory.
ZSB_HDN int seq_fast_phase1(const uint8_t *src, const ZsbBlockWork &w, const SeqTables &T, uint32_t *words, int ws, uint32_t &rem0) {
    const uint64_t start = w.bs_off, end = w.bs_off + w.bs_len;
    if (end <= start) return ZSB_E_EMPTY_INPUT_DATA;
    const uint32_t lastb = src[end - 1];
    if (lastb == 0) return ZSB_E_NULL_BYTE;
    if (start < 16 || w.bs_len > (1u << 24)) return ZSB_NEEDS_SLOW;            // no word below the stream / positions would not fit
    const uint32_t a0 = (uint32_t)T.al[0], a1 = (uint32_t)T.al[1], a2 = (uint32_t)T.al[2];
    const uint32_t nseq = w.nseq;
    FastWin F;
    const uint32_t mis = (uint32_t)((uintptr_t)src & 7);
    const uint64_t w0 = ((start + mis) & ~7ull) - 8;                          // byte offset of pw from the aligned base
    const uint8_t *pw = src - mis + w0;
    int32_t top = (int32_t)((end - 1 + mis - w0) * 8) + zsb_flog2(lastb);
    const int32_t startbit = (int32_t)((start + mis - w0) * 8);
    if (top - startbit < (int32_t)(a0 + a1 + a2)) return ZSB_E_NOT_ENOUGH_BITS;
    fast_win_load(F, pw, top);
    uint64_t W = fast_win_get(F);
    uint32_t sL = (uint32_t)zsb_shr64(W, 64 - a0);
    uint32_t sO = (uint32_t)zsb_shr64(zsb_shl64(W, a0), 64 - a1);
    uint32_t sM = (uint32_t)zsb_shr64(zsb_shl64(W, a0 + a1), 64 - a2);
    top -= (int32_t)(a0 + a1 + a2);
    rem0 = (uint32_t)(top - startbit);
    fast_win_load(F, pw, top);
    const int ts = T.ts;
    const uint32_t *tL = T.tbl[0], *tO = T.tbl[1], *tM = T.tbl[2];
    for (uint32_t i = 0; i < nseq; i++) {
        const uint32_t eL = tL[sL * ts], eO = tO[sO * ts], eM = tM[sM * ts];
        W = fast_win_get(F);
        const uint32_t sum = eL + eO + eM;                        // byte 0: extra bits, byte 1: state bits (no carries: <= 63, <= 27)
        const uint32_t nbs = i + 1 == nseq ? 0u : (sum >> 8) & 0xFFu;    // no state update after the last sequence (sequence.rs:80)
        const uint32_t px = sum & 0xFFu;
        uint32_t skip = px;
        if (px + nbs > 64) { top -= (int32_t)px; fast_win_load(F, pw, top); W = fast_win_get(F); skip = 0; }   // state bits past the window: rare
        const uint32_t t = (uint32_t)(zsb_shl64(W, skip) >> 32);  // the <= 27 state bits, top-aligned
        // the funnel shifts take their 5-bit amounts from the cells (nb in bits 8..12)
        const uint32_t nL = eL >> 8, nM = eM >> 8, nO = eO >> 8;
        const uint32_t bL = zsb_fsl(t, 0, nL), t2 = zsb_fsl(0, t, nL), bM = zsb_fsl(t2, 0, nM), bO = zsb_fsl(zsb_fsl(0, t2, nM), 0, nO);
        top -= (int32_t)(skip + nbs);
        fast_win_load(F, pw, top);
        sL = ZSB_CELL_BASE(eL) + bL; sM = ZSB_CELL_BASE(eM) + bM; sO = ZSB_CELL_BASE(eO) + bO;      // sequence.rs:80-88
        words[i * ws] = seq_fast_word(eL, eO, eM, i + 1 == nseq ? 0u : sum);
    }
    // an over-read shows as a cursor below the stream start; illegal codes are caught by phase 2, which sees every code
    if (top < startbit) return ZSB_NEEDS_SLOW;
    return ZSB_OK;
}

// ---- phase 2 ---------------------------------------------------------------------------------------
// (ll, ml, offset_value) of one sequence from its word and `top`, the absolute bit just above its
// first bit (sequence.rs:50-55: OF extra, then ML extra, then LL extra, MSB first).
ZSB_HD void seq_fast_values(const uint8_t *base8, int64_t top, uint32_t word, const uint32_t *lltab, const uint32_t *mltab,
                            uint32_t &ll, uint32_t &ml, uint32_t &ov, uint32_t &px, int &bad) {
    uint32_t cL = ZSB_W_CL(word), cO = ZSB_W_CO(word), cM = ZSB_W_CM(word);
    if (cL > ZSB_MAX_LL_CODE || cO > ZSB_MAX_OF_CODE || cM > ZSB_MAX_ML_CODE) { bad = 1; cL = cO = cM = 0; }   // sequence.rs:46-48
    const uint32_t eL = lltab[cL], eM = mltab[cM];
    const uint32_t xL = eL >> 24, xM = eM >> 24;
    px = xL + xM + cO;
    const uint64_t W = px ? fast_win_at(base8, top - px, px) : 0ull;
    ll = (eL & 0xFFFFFFu) + ((uint32_t)W & ((1u << xL) - 1u));
    ml = (eM & 0xFFFFFFu) + ((uint32_t)(W >> xL) & ((1u << xM) - 1u));
    ov = (1u << cO) + ((uint32_t)(W >> (xL + xM)) & ((1u << cO) - 1u));
}

// Repeat-offset history as a composable transform.  Each of the three slots is a coded offset in
// the sense of zsb_common.h: a constant, or (ZSB_OFF_SYM) "incoming slot k minus d".
struct Hist { uint32_t h0, h1, h2; };
#define ZSB_HSLOT(k) (ZSB_OFF_SYM | ((uint32_t)(k) << 25))
ZSB_HD Hist hist_identity() { Hist h; h.h0 = ZSB_HSLOT(0); h.h1 = ZSB_HSLOT(1); h.h2 = ZSB_HSLOT(2); return h; }
// what decode_offset (decoding_context.rs:50-75) does to the history for one sequence; h0 of the
// result is the offset the sequence uses
ZSB_HD Hist hist_of_sequence(uint32_t ov, uint32_t ll, int &bad) {
    Hist h;
    if (ov > 3) {
        uint32_t off = ov - 3;
        if (off > ZSB_OFF_MAX) { bad = 1; off = 1; }
        h.h0 = off; h.h1 = ZSB_HSLOT(0); h.h2 = ZSB_HSLOT(1);
        return h;
    }
    const uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);
    if (idx == 0) return hist_identity();
    if (idx == 1) { h.h0 = ZSB_HSLOT(1); h.h1 = ZSB_HSLOT(0); h.h2 = ZSB_HSLOT(2); }
    else if (idx == 2) { h.h0 = ZSB_HSLOT(2); h.h1 = ZSB_HSLOT(0); h.h2 = ZSB_HSLOT(1); }
    else { h.h0 = ZSB_HSLOT(0) + 1u; h.h1 = ZSB_HSLOT(0); h.h2 = ZSB_HSLOT(1); }
    return h;
}
// value of coded slot f once the history it refers to is g
ZSB_HD uint32_t hist_pick(const Hist &g, uint32_t f, int &bad) {
    if (!(f & ZSB_OFF_SYM)) return f;
    const uint32_t k = ZSB_OFF_SLOT(f), d = ZSB_OFF_DEC(f);
    const uint32_t x = k == 0 ? g.h0 : k == 1 ? g.h1 : g.h2;
    if (x & ZSB_OFF_SYM) return x + d;
    if (x <= d) { bad = 1; return 1u; }
    return x - d;
}
// f after g
ZSB_HD Hist hist_compose(const Hist &f, const Hist &g, int &bad) {
    Hist r; r.h0 = hist_pick(g, f.h0, bad); r.h1 = hist_pick(g, f.h1, bad); r.h2 = hist_pick(g, f.h2, bad);
    return r;
}
