// emul.cpp -- CPU-only TEST HARNESS for the lane-serial device code (test infrastructure, not product).
//
// The CUDA kernels run one lane per block / per stream over host+device inline functions
// (zsb_parse.h, zsb_fse.h, zsb_huf.h, zsb_seq.h).  This file compiles those same functions with g++
// and drives them serially, so that `pytest -m "not gpu"` can diff every intermediate product
// (tables, weights, literals, sequence records, statuses) against the CPU oracle without a GPU.
// It is never loaded by the product package, bench.py's GPU arm or smoke().
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../zstd-decompressor_b200/csrc/zsb_parse.h"
#include "../../zstd-decompressor_b200/csrc/zsb_huf.h"
#include "../../zstd-decompressor_b200/csrc/zsb_seqfast.h"

// fast-path bookkeeping for tests: blocks where the fast sequence path ran / agreed / asked for the careful path / disagreed
static long g_fast_ran = 0, g_fast_same = 0, g_fast_slow = 0, g_fast_diff = 0;
static long g_huf_ran = 0, g_huf_same = 0, g_huf_slow = 0, g_huf_diff = 0;

// The fast sequence path as the kernels run it (phase 1 per lane, phase 2 folded serially with the same
// per-sequence functions and the same history composition) against the careful decoder's results.
static void check_fast_path(const uint8_t *src, const ZsbBlockWork &w0, const SeqTables &T, int careful_rc, const ZsbBlockWork &wc,
                            const std::vector<uint64_t> &rec_c) {
    const int WS = 3;   // any stride
    std::vector<uint32_t> words((size_t)w0.nseq * WS + 1, 0xABABABAB);
    uint32_t rem0 = 0;
    g_fast_ran++;
    int rc = seq_fast_phase1(src, w0, T, words.data(), WS, rem0);
    if (rc > 0) { if (rc == careful_rc) g_fast_same++; else g_fast_diff++; return; }
    int bad = 0;
    std::vector<uint64_t> rec(w0.nseq, 0);
    uint32_t lit = 0, out = 0; Hist H = hist_identity();
    if (rc == ZSB_OK) {
        uint32_t lltab[36], mltab[53];
        for (uint32_t c = 0; c < 36; c++) lltab[c] = zsb_ll_entry(c);
        for (uint32_t c = 0; c < 53; c++) mltab[c] = zsb_ml_entry(c);
        const uint32_t mis = (uint32_t)((uintptr_t)src & 7);
        int64_t top = (int64_t)(w0.bs_off + mis) * 8 + rem0;
        for (uint32_t i = 0; i < w0.nseq; i++) {
            const uint32_t word = words[(size_t)i * WS];
            uint32_t ll, ml, ov, px;
            seq_fast_values(src - mis, top, word, lltab, mltab, ll, ml, ov, px, bad);
            if (bad) break;
            top -= px + ZSB_W_NB(word);
            Hist F = hist_of_sequence(ov, ll, bad);
            H = hist_compose(F, H, bad);
            lit += ll; out += ll + ml;
            if (lit > w0.lit_regen || out + (w0.lit_regen - lit) > ZSB_BLOCK_MAX) { bad = 1; break; }
            rec[i] = (uint64_t)out | ((uint64_t)lit << ZSB_REC_POS_BITS) | ((uint64_t)H.h0 << (2 * ZSB_REC_POS_BITS));
        }
    }
    if (rc == ZSB_NEEDS_SLOW || bad) { if (careful_rc != ZSB_OK) g_fast_slow++; else g_fast_diff++; return; }
    bool same = careful_rc == ZSB_OK && wc.lit_used == lit && wc.out_size == out + (w0.lit_regen - lit) &&
                wc.rep_out[0] == H.h0 && wc.rep_out[1] == H.h1 && wc.rep_out[2] == H.h2;
    for (uint32_t i = 0; same && i < w0.nseq; i++) same = rec[i] == rec_c[i];
    if (same) g_fast_same++; else g_fast_diff++;
}

extern "C" {

int emul_fse_parse(const uint8_t *desc, size_t n, int type, int *al_out, int *nsym_out, int16_t *dist, uint32_t *cells, uint32_t *consumed) {
    FwdBits f; fwd_init(f, desc, n);
    int16_t cnt[256]; int al = 0, nsym = 0;
    int rc = fse_read_ncount(f, cnt, 1, 256, al, nsym);
    if (rc) return rc;
    *al_out = al; *nsym_out = nsym; *consumed = fwd_bytes_read(f);
    if (dist) memcpy(dist, cnt, sizeof(int16_t) * (size_t)nsym);
    return fse_build_table(cnt, 1, nsym, al, cells, 1, type);
}
int emul_fse_build(int type, int al, const int16_t *dist, int nsym, uint32_t *cells, int stride) {
    std::vector<int16_t> cnt(256 * (size_t)stride, 0);
    for (int i = 0; i < nsym; i++) cnt[(size_t)i * stride] = dist[i];
    return fse_build_table(cnt.data(), stride, nsym, al, cells, stride, type);
}
int emul_huf_parse(const uint8_t *desc, size_t n, uint8_t *lens, uint16_t *lut, int *maxbits, uint32_t *consumed, uint8_t *weights_out, int *nw_out) {
    uint8_t weights[260]; uint32_t ftbl[512]; int16_t cnt[16]; uint32_t rank[16];
    int nw = 0, mb = 0; uint32_t dl = 0;
    std::vector<uint8_t> padded(desc, desc + n); padded.resize(n + 16, 0);
    int rc = huf_read_weights(padded.data(), n, weights, 1, nw, dl, ftbl, 1, cnt, 1, n, false);
    if (rc) return rc;
    if (weights_out) { memcpy(weights_out, weights, (size_t)nw); *nw_out = nw; }
    rc = huf_build_lut(weights, 1, nw, lut, rank, 1, mb, lens);
    *maxbits = mb; *consumed = dl;
    return rc;
}

// k_huf's two tables from one set of weights (nw weights, the last one implied): the one-symbol table (2^maxbits cells) and the two-symbol
// table (1 024 cells) built the way the kernel builds it -- T1 and `odd` filled symbol by symbol from the cell starts (huf_t1_put), then
// huf_pair_cell -- and, in pair_b, the same table derived from the finished one-symbol table (huf_pairs_from_lut).
int emul_huf_pairs(const uint8_t *weights_in, int nw, uint16_t *lut, uint32_t *pair_a, uint32_t *pair_b, int *maxbits) {
    uint8_t weights[260] = {0}; uint32_t rank[16]; uint16_t at[256];
    memcpy(weights, weights_in, (size_t)nw);
    HufPlan P; bool inc = false;
    int rc = huf_lut_plan(weights, 1, nw, rank, 1, P, false, &inc);
    if (rc) return rc;
    if (inc) return ZSB_E_CORRUPT;
    for (int i = 0; i < P.n; i++) {                                        // the cell starts, as in k_huf
        const uint32_t wt = weights[i];
        at[i] = 0xFFFF;
        if (wt) { at[i] = (uint16_t)rank[wt]; rank[wt] += 1u << (wt - 1); }
    }
    static uint8_t t1[1 << ZSB_HUF_PAIR_BITS], odd[512];
    memset(t1, 0xEE, sizeof t1); memset(odd, 0xEE, sizeof odd);
    for (int i = 0; i < P.n; i++) if (weights[i]) huf_t1_put(t1, odd, P.mb, (uint32_t)i, at[i], weights[i]);
    for (uint32_t x = 0; x < (1u << ZSB_HUF_PAIR_BITS); x++) pair_a[x] = huf_pair_cell(x, t1, odd, weights, 1, P.mb);
    uint8_t w2[260] = {0}; memcpy(w2, weights_in, (size_t)nw);
    int mb = 0;
    rc = huf_build_lut(w2, 1, nw, lut, rank, 1, mb, nullptr);
    if (rc) return rc;
    static uint8_t t1b[1 << ZSB_HUF_PAIR_BITS], oddb[512];
    huf_pairs_from_lut(lut, mb, w2, 1, t1b, oddb, pair_b);
    *maxbits = mb;
    return mb == P.mb ? ZSB_OK : ZSB_E_CORRUPT;
}

// Whole pipeline, serially: scan -> parse -> chain -> huffman -> sequences -> plan -> execute.
// out/out_cap: output buffer; per-frame arrays sized n_frames as reported by zsb_scan.
// Intermediates of the LAST compressed block that ran are exported for stage-level diffs when the
// pointers are non-null (lits: literals, recs: packed sequence records).
int emul_decode(const uint8_t *src_in, size_t n, uint32_t flags, uint8_t *out, size_t out_cap, uint64_t *out_len,
                int32_t *status, uint64_t *foff, uint64_t *flen, size_t frames_cap, size_t *n_frames_out, uint64_t *err_a, uint64_t *err_b) {
    std::vector<uint8_t> srcv(src_in, src_in + n); srcv.resize(n + 64, 0);   // same padding as the device copy
    const uint8_t *src = srcv.data();
    zsb_frame *frames = nullptr; zsb_block *blocks = nullptr; size_t nf = 0, nb = 0;
    int scan_rc = zsb_scan(src, n, flags, 0, &frames, &nf, &blocks, &nb, err_a, err_b);
    *n_frames_out = nf;
    if (nf > frames_cap) { zsb_free(frames); zsb_free(blocks); return ZSB_E_ARG; }
    std::vector<ZsbBlockWork> work(nb + 1);
    for (size_t i = 0; i < nb; i++) parse_block(src, blocks[i], work[i], flags);                        // k_parse
    std::vector<int> fstatus(nf, 0);
    for (size_t f = 0; f < nf; f++) {                                                                     // k_plan1 (a)
        fstatus[f] = frames[f].status; uint32_t ea = 0, eb = 0;
        if (!fstatus[f] && frames[f].kind == 0) fstatus[f] = chain_frame(frames[f], blocks, work.data(), flags, ea, eb);
    }
    std::vector<std::vector<uint8_t>> lits(nb);
    std::vector<std::vector<uint64_t>> recs(nb);
    for (size_t i = 0; i < nb; i++) {
        ZsbBlockWork &w = work[i];
        if (blocks[i].type != ZSB_BT_COMPRESSED || (w.status && !ZSB_CHAIN_SEQ_ERROR(w.status))) continue;          // k_plan1 (b): the lists
        const int chain_st = w.status;
        if (w.lit_type >= ZSB_LT_COMPRESSED) {                                                            // k_huf
            uint8_t weights[260]; uint32_t ftbl[512]; int16_t cnt[16]; uint32_t rank[16]; static uint16_t lut[2048];
            int nw = 0, mb = 0; uint32_t dl = 0;
            int rc = huf_read_weights(src + w.huf_desc, w.huf_desc_end - w.huf_desc, weights, 1, nw, dl, ftbl, 1, cnt, 1, n - w.huf_desc,
                                      (flags & ZSB_REFERENCE_QUIRKS) != 0);
            bool incomplete = false;
            if (!rc) rc = huf_build_lut(weights, 1, nw, lut, rank, 1, mb, nullptr, (flags & ZSB_REFERENCE_QUIRKS) != 0, &incomplete);
            lits[i].assign(w.lit_regen + 16, 0);
            const bool quirks = (flags & ZSB_REFERENCE_QUIRKS) != 0;
            bool inexact = quirks && (w.lit_inexact || incomplete);
            uint64_t start = w.lit_src; uint32_t seg = (w.lit_regen + 3) / 4;
            for (uint32_t s = 0; s < w.n_streams && !rc && !inexact; s++) {
                uint32_t expect = w.n_streams == 1 ? w.lit_regen : (s < 3 ? seg : w.lit_regen - 3 * seg);
                uint8_t *o = lits[i].data() + (w.n_streams == 1 ? 0 : s * seg);
                rc = huf_decode_stream(src, start, start + w.stream_size[s], n, lut, mb, o, expect);
                {   // the fast stream decode on the same stream: same bytes, or a request for the careful decoder where that one fails
                    std::vector<uint8_t> fo((size_t)expect + 16, 0xEE);
                    uint8_t *fa = fo.data() + ((4 - ((uintptr_t)fo.data() & 3)) & 3) + ((uintptr_t)o & 3);   // same alignment as the real output
                    int frc = ZSB_NEEDS_SLOW;
                    if (!incomplete) {                                                  // k_huf: the two-symbol table of a complete code
                        static uint32_t pair[1 << ZSB_HUF_PAIR_BITS]; uint8_t t1[1 << ZSB_HUF_PAIR_BITS], odd[512];
                        huf_pairs_from_lut(lut, mb, weights, 1, t1, odd, pair);
                        frc = huf_fast_stream(src, start, start + w.stream_size[s], pair, fa, expect, 0);
                    }
                    g_huf_ran++;
                    if (frc == ZSB_OK) { if (rc == ZSB_OK && memcmp(fa, o, expect) == 0) g_huf_same++; else g_huf_diff++; }
                    else if (frc == ZSB_NEEDS_SLOW) { if (rc != ZSB_OK) g_huf_slow++; else g_huf_diff++; }
                    else g_huf_diff++;
                    if (quirks && frc == ZSB_NEEDS_SLOW) { inexact = true; rc = 0; }      // k_huf: the block goes the reference's way
                }
                start += w.stream_size[s];
            }
            if (inexact && !rc) {                                                          // k_huf: huf_decode_block_ref, count then decode
                uint32_t n1 = 0, n2 = 0;
                rc = huf_decode_block_ref(src, n, w.lit_src, w.stream_size, lut, mb, nullptr, 0, n1);
                if (!rc && n1 > ZSB_BLOCK_MAX) rc = ZSB_E_BLOCK_TOO_LARGE;
                if (!rc) { lits[i].assign(n1 + 16, 0); rc = huf_decode_block_ref(src, n, w.lit_src, w.stream_size, lut, mb, lits[i].data(), n1, n2); }
                if (!rc) w.lit_regen = n1;
            }
            if (rc) { w.status = rc; continue; }
        }
        if (w.nseq && !chain_st) {                                                                        // k_seq (interleaved layout, lane 5 of 32)
            const int TS = 32, LANE = 5;
            std::vector<uint32_t> tbl(3 * 512 * TS, 0xDEADBEEF); std::vector<int16_t> cnt(256 * TS, 0);
            uint32_t bases[89];
            for (uint32_t k = 0; k < 89; k++) bases[k] = k < 36 ? zsb_ll_base(k) : zsb_ml_base(k - 36);
            SeqTables T; T.ts = TS; T.max_al[0] = T.max_al[1] = T.max_al[2] = 0; T.tbl[0] = tbl.data() + LANE; T.tbl[1] = tbl.data() + 512 * TS + LANE; T.tbl[2] = tbl.data() + 2 * 512 * TS + LANE;
            recs[i].assign(w.nseq + 1, 0);
            int rc = seq_build_tables(src, w, T, cnt.data() + LANE, TS);
            const ZsbBlockWork w0 = w;
            const int trc = rc;
            if (!rc) rc = seq_decode(src, n, w, T, bases, bases + 36, recs[i].data());
            if (!trc) check_fast_path(src, w0, T, rc, w, recs[i]);
            if (rc) { w.status = rc; continue; }
        }
    }
    uint64_t pos = 0;
    for (size_t f = 0; f < nf; f++) {                                                                     // k_plan2 + k_rawrle + k_exec
        uint64_t len = 0; int st = fstatus[f];
        if (!st) {
            if (frames[f].kind == 1) len = (flags & ZSB_PRINT_SKIPPABLE) ? blocks[frames[f].first_block].size : 0;
            else {
                uint32_t pa = 0, pb = 0;
                st = plan_frame(frames[f], blocks, work.data(), len, pa, pb);
                if (!st && frames[f].has_content_size && len != frames[f].content_size && !(flags & ZSB_REFERENCE_QUIRKS)) st = ZSB_E_CONTENT_SIZE;
                if (st) len = 0;
            }
        }
        if (!st && pos + len > out_cap) { st = ZSB_E_DST_TOO_SMALL; }
        if (st) len = 0;
        foff[f] = pos; flen[f] = len; status[f] = st;
        if (st || !len) continue;
        uint8_t *fd = out + pos;
        for (uint32_t k = 0; k < frames[f].n_blocks && !st; k++) {
            const uint32_t bi = frames[f].first_block + k; const zsb_block &b = blocks[bi]; ZsbBlockWork &w = work[bi];
            uint8_t *o = fd + w.out_off;
            if (b.type == ZSB_BT_RLE) { memset(o, src[b.src_off], b.size); continue; }
            if (b.type != ZSB_BT_COMPRESSED) { memcpy(o, src + b.src_off, b.size); continue; }
            auto lit = [&](uint32_t i) -> uint8_t { return w.lit_type == ZSB_LT_RAW ? src[w.lit_src + i] : w.lit_type == ZSB_LT_RLE ? src[w.lit_src] : lits[bi][i]; };
            uint32_t oe = 0, le = 0;
            for (uint32_t s = 0; s < w.nseq; s++) {
                const uint64_t r = recs[bi][s];
                const uint32_t out_end = (uint32_t)r & ZSB_REC_POS_MASK, lit_end = (uint32_t)(r >> 18) & ZSB_REC_POS_MASK;
                const uint32_t off = seq_real_offset((uint32_t)(r >> 36), w.rep_in);
                const uint32_t ll = lit_end - le, ml = out_end - oe - ll;
                if (off == 0 || off > w.out_off + oe + ll) { st = ZSB_E_IMPOSSIBLE_VALUE; break; }
                for (uint32_t i = 0; i < ll; i++) o[oe + i] = lit(le + i);
                for (uint32_t i = 0; i < ml; i++) o[oe + ll + i] = o[(int64_t)oe + ll + i - off];
                oe = out_end; le = lit_end;
            }
            for (uint32_t i = 0; !st && i < w.lit_regen - le; i++) o[oe + i] = lit(le + i);
        }
        if (st) { status[f] = st; flen[f] = 0; }
        pos += len;
    }
    *out_len = pos;
    zsb_free(frames); zsb_free(blocks);
    return scan_rc;
}
void emul_huf_stats(long *ran, long *same, long *slow, long *diff) { *ran = g_huf_ran; *same = g_huf_same; *slow = g_huf_slow; *diff = g_huf_diff; }
void emul_fast_stats(long *ran, long *same, long *slow, long *diff) { *ran = g_fast_ran; *same = g_fast_same; *slow = g_fast_slow; *diff = g_fast_diff; }

// associativity of the history composition on random transforms: returns the number of violations
static uint32_t rnd(uint64_t &s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
static Hist rnd_hist(uint64_t &s) {
    int bad = 0;
    uint32_t ov = (rnd(s) % 3 == 0) ? 1 + rnd(s) % 3 : 4 + rnd(s) % 100000;
    return hist_of_sequence(ov, rnd(s) % 2, bad);
}
int emul_hist_assoc(uint64_t seed, int n) {
    int viol = 0;
    for (int i = 0; i < n; i++) {
        Hist a = rnd_hist(seed), b = rnd_hist(seed), c = rnd_hist(seed), d = rnd_hist(seed), e = rnd_hist(seed);
        int b1 = 0, b2 = 0;
        // ((a.b).(c.d)).e  vs  a.(b.(c.(d.e)))
        Hist l = hist_compose(hist_compose(hist_compose(a, b, b1), hist_compose(c, d, b1), b1), e, b1);
        Hist r = hist_compose(a, hist_compose(b, hist_compose(c, hist_compose(d, e, b2), b2), b2), b2);
        if (b1 || b2) continue;      // an offset reached zero: the careful path takes over
        if (l.h0 != r.h0 || l.h1 != r.h1 || l.h2 != r.h2) viol++;
    }
    return viol;
}
}  // extern "C"
