// zsb_seq.h -- the three-state interleaved sequence decode of one block, one lane per block.
//
//  seq_build_tables == Sequences::get_decoder x3 (sequences.rs:147-187): predefined / RLE /
//                      FSE-compressed / repeat (already resolved to its source by the chain pass)
//  seq_decode       == Sequences::decode (sequences.rs:191-237) + SequenceDecoder
//                      (decoders/sequence.rs:41-88) + FseDecoder (decoders/fse.rs:279-318)
//                      + DecodingContext::decode_offset (decoding_context.rs:50-75)
//
// Bit order per sequence (sequence.rs:41-55,80-88): OF extra, ML extra, LL extra, then the state
// updates LL, ML, OF (none after the last sequence).  All six field widths are known once the three
// current cells are loaded, so the fast path extracts every field from one 64-bit register window
// with independent shifts; only sequences needing more than 64 bits take the stepwise path.
//
// Repeat offsets: the history a block starts with is only known once the previous block of the
// frame is decoded, and blocks decode concurrently.  The lane therefore tracks the history
// symbolically (ZSB_OFF_SYM: slot of the incoming history + number of "-1" applications) and the
// executor substitutes the real values (rep_in) afterwards.
#pragma once
#include "zsb_fse.h"

struct SeqTables {
    uint32_t *tbl[3];   // LL, OF, ML cells; cell i at tbl[t][i*ts]
    int ts;
    int al[3];
    int max_al[3];      // largest accuracy log each table has room for (0 = ZSB_MAX_AL)
};
#define ZSB_TABLE_TOO_SMALL (-2)   // internal: seq_build_tables met a table larger than the caller's storage

// Builds table t (0 LL, 1 OF, 2 ML) of block `w` (modes already resolved, never ZSB_M_REPEAT) into tbl (stride ts).
// cnt: scratch for max_sym counts (stride cs); max_al: largest accuracy log tbl has room for (0 = ZSB_MAX_AL).
// A description with more symbols than max_sym or a larger accuracy log returns ZSB_TABLE_TOO_SMALL.
ZSB_HDN int seq_build_table(const uint8_t *src, const ZsbBlockWork &w, int t, uint32_t *tbl, int ts, int16_t *cnt, int cs, int max_sym, int max_al,
                            int &al_out) {
    const int mode = w.mode[t];
    if (mode == ZSB_M_RLE) { fse_build_rle(w.rle_sym[t], tbl, t); al_out = 0; return ZSB_OK; }
    int al, nsym, rc;
    if (mode == ZSB_M_PREDEFINED) fse_predefined_counts(t, cnt, cs, al, nsym);
    else if (mode == ZSB_M_FSE) {
        FwdBits f; fwd_init(f, src + w.tbl_desc[t], w.tbl_end - w.tbl_desc[t]);
        rc = fse_read_ncount(f, cnt, cs, max_sym, al, nsym);
        if (rc == ZSB_E_CORRUPT && max_sym < 256) return ZSB_TABLE_TOO_SMALL;      // more symbols than the scratch holds
        if (rc) return rc;
    } else return ZSB_E_NO_PREVIOUS_DECODER;
    if (max_al && al > max_al) return ZSB_TABLE_TOO_SMALL;
    rc = fse_build_table(cnt, cs, nsym, al, tbl, ts, t);
    if (rc) return rc;
    al_out = al;
    return ZSB_OK;
}
// Builds the three tables of block `w` in the reference's order (LL, OF, ML: sequences.rs:116-141).
ZSB_HDN int seq_build_tables(const uint8_t *src, const ZsbBlockWork &w, SeqTables &T, int16_t *cnt, int cs) {
    for (int t = 0; t < 3; t++) {
        const int rc = seq_build_table(src, w, t, T.tbl[t], T.ts, cnt, cs, 256, T.max_al[t], T.al[t]);
        if (rc) return rc;
    }
    return ZSB_OK;
}

// decode_offset (decoding_context.rs:50-75) on the coded history h[0..2]
ZSB_HD uint32_t seq_resolve_offset(uint32_t ov, uint32_t ll, uint32_t &h0, uint32_t &h1, uint32_t &h2, int &err) {
    uint32_t off;
    if (ov > 3) {
        off = ov - 3;
        if (off > ZSB_OFF_MAX) { err = ZSB_E_IMPOSSIBLE_VALUE; off = 1; }
        h2 = h1; h1 = h0; h0 = off;
    } else {
        uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);
        if (idx == 0) off = h0;
        else if (idx == 1) { off = h1; h1 = h0; h0 = off; }
        else if (idx == 2) { off = h2; h2 = h1; h1 = h0; h0 = off; }
        else {
            if (h0 & ZSB_OFF_SYM) off = h0 + 1;          // one more "-1" on an incoming slot
            else { off = h0 - 1; if (off == 0) { err = ZSB_E_IMPOSSIBLE_VALUE; off = 1; } }   // reference: index panic
            h2 = h1; h1 = h0; h0 = off;
        }
    }
    return off;
}

// Decodes w.nseq sequences of one block into packed records rec[0..nseq).
// ll_base / ml_base: the code -> baseline tables (shared memory copies on the GPU).
// Results: w.out_size, w.lit_used, w.rep_out; returns the block status.
ZSB_HDN int seq_decode(const uint8_t *src, uint64_t src_end, ZsbBlockWork &w, const SeqTables &T,
                       const uint32_t *ll_base, const uint32_t *ml_base, uint64_t *rec) {
    BackWin b;
    int rc = back_init(b, src, w.bs_off, w.bs_off + w.bs_len, src_end);
    if (rc) return rc;
    const int ts = T.ts;
    const uint32_t *tL = T.tbl[0], *tO = T.tbl[1], *tM = T.tbl[2];
    back_refill(b);
    // SequenceDecoder::initialize: LL, OF, ML (sequence.rs:59-65)
    uint32_t sL = back_peek(b, 0, (uint32_t)T.al[0]);
    uint32_t sO = back_peek(b, (uint32_t)T.al[0], (uint32_t)T.al[1]);
    uint32_t sM = back_peek(b, (uint32_t)(T.al[0] + T.al[1]), (uint32_t)T.al[2]);
    back_consume(b, (uint32_t)(T.al[0] + T.al[1] + T.al[2]));
    if (b.rem < 0) return ZSB_E_NOT_ENOUGH_BITS;
    uint32_t h0 = ZSB_OFF_SYM | (0u << 25), h1 = ZSB_OFF_SYM | (1u << 25), h2 = ZSB_OFF_SYM | (2u << 25);
    uint32_t out_end = 0, lit_end = 0;
    const uint32_t nseq = w.nseq, regen = w.lit_regen;
    int err = 0, xerr = 0;   // xerr: execution-time error (the reference raises it only after decoding every sequence)
    for (uint32_t i = 0; i < nseq; i++) {
        back_refill(b);
        const uint32_t eL = tL[sL * ts], eO = tO[sO * ts], eM = tM[sM * ts];
        const uint32_t xO = ZSB_CELL_XB(eO), xM = ZSB_CELL_XB(eM), xL = ZSB_CELL_XB(eL);
        const uint32_t cO = ZSB_CELL_CODE(eO), cM = ZSB_CELL_CODE(eM), cL = ZSB_CELL_CODE(eL);
        if (cL > ZSB_MAX_LL_CODE || cM > ZSB_MAX_ML_CODE || cO > ZSB_MAX_OF_CODE) { err = ZSB_E_SEQ_CODE_MAX; break; }  // sequence.rs:46-48
        const bool last = (i + 1 == nseq);
        const uint32_t nL = last ? 0u : ZSB_CELL_NB(eL), nM = last ? 0u : ZSB_CELL_NB(eM), nO = last ? 0u : ZSB_CELL_NB(eO);
        const uint32_t px = xO + xM + xL, total = px + nL + nM + nO;
        uint32_t vO, vM, vL, bL, bM, bO;
        if (total <= 64) {
            vO = back_peek(b, 0, xO); vM = back_peek(b, xO, xM); vL = back_peek(b, xO + xM, xL);
            bL = back_peek(b, px, nL); bM = back_peek(b, px + nL, nM); bO = back_peek(b, px + nL + nM, nO);
            back_consume(b, total);
        } else {
            vO = back_take(b, xO); vM = back_take(b, xM); vL = back_take(b, xL);
            bL = back_take(b, nL); bM = back_take(b, nM); bO = back_take(b, nO);
        }
        if (b.rem < 0) { err = ZSB_E_NOT_ENOUGH_BITS; break; }
        const uint32_t ov = (1u << cO) + vO;                      // sequence.rs:50
        const uint32_t ml = ml_base[cM] + vM, ll = ll_base[cL] + vL;
        const uint32_t off = seq_resolve_offset(ov, ll, h0, h1, h2, xerr);
        if (!xerr) {
            lit_end += ll; out_end += ll + ml;
            if (lit_end > regen) xerr = ZSB_E_IMPOSSIBLE_VALUE;                  // decoding_context.rs:86-90 (ll > literals left)
            else if (out_end + (regen - lit_end) > ZSB_BLOCK_MAX) xerr = ZSB_E_BLOCK_TOO_LARGE;
            else rec[i] = (uint64_t)out_end | ((uint64_t)lit_end << ZSB_REC_POS_BITS) | ((uint64_t)off << (2 * ZSB_REC_POS_BITS));
        }
        sL = ZSB_CELL_BASE(eL) + bL; sM = ZSB_CELL_BASE(eM) + bM; sO = ZSB_CELL_BASE(eO) + bO;   // sequence.rs:80-88
    }
    if (err) return err;
    if (xerr) return xerr;
    w.lit_used = lit_end;
    w.out_size = out_end + (regen - lit_end);
    w.rep_out[0] = h0; w.rep_out[1] = h1; w.rep_out[2] = h2;
    return ZSB_OK;
}

// Substitute the block's incoming history into a coded offset.  Returns 0 if the result is invalid.
ZSB_HD uint32_t seq_real_offset(uint32_t coded, const uint32_t *rep_in) {
    if (!(coded & ZSB_OFF_SYM)) return coded;
    uint32_t base = rep_in[ZSB_OFF_SLOT(coded)], dec = ZSB_OFF_DEC(coded);
    return base > dec ? base - dec : 0u;
}
