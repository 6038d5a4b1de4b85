// zsb_fse.h -- FSE table description parser and decoding-table construction, one lane per table.
//
//  fse_read_ncount  == parse_fse_table              (decoders/fse.rs:16-69)
//  fse_build_table  == FseTable::from_distribution  (decoders/fse.rs:110-202)
//
// Storage is strided so that 32 lanes of a warp can each build their own table into an interleaved
// shared-memory layout (cell i of lane l at word i*32 + l: conflict-free in the decode loop); the
// CPU tests call the same code with stride 1.
#pragma once
#include "zsb_bits.h"

// Reads the normalized counts.  cnt[s*cs] receives the count of symbol s (-1 = "less than one").
// On success: al, nsym (symbols described, trailing zero-runs included), f.pos advanced.
// max_sym: capacity of cnt (the reference accepts up to 255 symbols, fse.rs:14,26,64).
ZSB_HDN int fse_read_ncount(FwdBits &f, int16_t *cnt, int cs, int max_sym, int &al, int &nsym) {
    uint32_t v;
    if (!fwd_take(f, 4, v)) return ZSB_E_NOT_ENOUGH_BITS;
    al = (int)v + 5;
    if (al > ZSB_MAX_AL) return ZSB_E_LARGE_ACCURACY_LOG;            // fse.rs:18-20
    int remaining = 1 << al, n = 0;
    while (remaining > 0 && n < 256) {                                 // fse.rs:26
        uint32_t nb = (uint32_t)zsb_flog2((uint32_t)(remaining + 1)) + 1;
        uint32_t pk;
        if (!fwd_peek(f, nb, pk)) return ZSB_E_NOT_ENOUGH_BITS;
        uint32_t low = (1u << (nb - 1)) - 1u;
        uint32_t thr = (1u << nb) - 1u - (uint32_t)(remaining + 1);
        int dec;
        if ((pk & low) < thr) { dec = (int)(pk & low); f.pos += nb - 1; }   // fse.rs:34-35
        else if (pk > low)    { dec = (int)pk - (int)thr; f.pos += nb; }     // fse.rs:36-37
        else                  { dec = (int)pk; f.pos += nb; }                // fse.rs:38-39
        int proba = dec - 1;
        remaining -= proba < 0 ? -proba : proba;
        if (n < max_sym) cnt[n * cs] = (int16_t)proba;
        n++;
        if (proba == 0) {                                                   // fse.rs:48-58
            for (;;) {
                uint32_t z;
                if (!fwd_take(f, 2, z)) return ZSB_E_NOT_ENOUGH_BITS;
                for (uint32_t k = 0; k < z; k++) { if (n < max_sym) cnt[n * cs] = 0; n++; }
                if (z != 3) break;
            }
        }
    }
    if (remaining != 0 || n >= 256) return ZSB_E_CORRUPTED_TABLE;     // fse.rs:64-66
    if (n > max_sym) return ZSB_E_CORRUPT;
    nsym = n;
    return ZSB_OK;
}

// Builds the decoding table of 1<<al cells.
//   cnt[s*cs]   : normalized counts (overwritten: becomes the per-symbol "next state" counter)
//   tbl[i*ts]   : receives ZSB_CELL(nb, xb, base, code)
//   type        : 0 LL / 1 OF / 2 ML select the extra-bits column and the legal-code limit;
//                 3 = plain table (Huffman weights): xb = 0, code = symbol (< 64 guaranteed by caller)
// Cell contents equal the reference's State{output, baseline, bits_to_read} (fse.rs:72-76): the
// per-symbol grouping of fse.rs:169-189 (lower states read one more bit) is the closed form
//   next = count + rank ; nb = al - floor(log2(next)) ; base = (next << nb) - N
// which tests/test_emul.py checks against the oracle's literal restatement on random distributions.
ZSB_HDN int fse_build_table(int16_t *cnt, int cs, int nsym, int al, uint32_t *tbl, int ts, int type) {
    const int N = 1 << al;
    int high = N - 1;
    for (int s = 0; s < nsym; s++)                                   // fse.rs:120-133
        if (cnt[s * cs] == -1) {
            if (high < 0) return ZSB_E_CORRUPTED_TABLE;
            tbl[high * ts] = (uint32_t)s; high--;
        }
    const int step = (N >> 1) + (N >> 3) + 3, mask = N - 1;           // fse.rs:136-157
    int pos = 0, placed = 0;
    for (int s = 0; s < nsym; s++) {
        int c = cnt[s * cs];
        if (c <= 0) continue;
        for (int k = 0; k < c; k++) {
            if (placed >= high + 1) return ZSB_E_CORRUPTED_TABLE;     // more cells than the table holds
            tbl[pos * ts] = (uint32_t)s; placed++;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    if (placed != high + 1) return ZSB_E_CORRUPTED_TABLE;             // fse.rs:160-166 (unfilled slot)
    for (int s = 0; s < nsym; s++) if (cnt[s * cs] == -1) cnt[s * cs] = 1;   // becomes the "next" counter
    for (int i = 0; i < N; i++) {
        uint32_t s = tbl[i * ts];
        uint32_t nx = (uint32_t)(uint16_t)cnt[s * cs];
        cnt[s * cs] = (int16_t)(nx + 1);
        uint32_t nb = (uint32_t)(al - zsb_flog2(nx));
        uint32_t base = (nx << nb) - (uint32_t)N;
        uint32_t code, xb;
        if (type == 3) { code = s; xb = 0; }
        else { code = s > (uint32_t)zsb_max_code(type) ? 63u : s; xb = zsb_code_xbits(type, s); }
        tbl[i * ts] = ZSB_CELL(nb, xb, base, code);
    }
    return ZSB_OK;
}

// Table of a single symbol (RLE mode, rle.rs:6-34): one cell, zero state bits.
ZSB_HD void fse_build_rle(uint32_t sym, uint32_t *tbl, int type) {
    uint32_t code = sym > (uint32_t)zsb_max_code(type) ? 63u : sym;
    tbl[0] = ZSB_CELL(0, zsb_code_xbits(type, sym), 0, code);
}
// Predefined distribution of table `type` (sequences.rs:29-39,172-185) into cnt.
ZSB_HD void fse_predefined_counts(int type, int16_t *cnt, int cs, int &al, int &nsym) {
    nsym = zsb_predef_nsym(type); al = zsb_predef_al(type);
    for (int s = 0; s < nsym; s++) cnt[s * cs] = (int16_t)zsb_predef_count(type, s);
}
