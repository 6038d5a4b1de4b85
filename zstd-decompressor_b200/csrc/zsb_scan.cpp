// zsb_scan.cpp -- zsb_scan: the host walk over frames and blocks (no CUDA in this file).
//
// == ForwardByteParser::iter + Frame::parse + ZStandard::parse + Block::parse of the reference
//    (parsing.rs:29-112, frame.rs:61-230, block.rs:43-72), stopping at block extents: literals and
//    sequences sections are parsed on the GPU (k_parse).
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "zsb_common.h"
#include "zsb_parse.h"
#include "zsb_huf.h"
#include "zsb_scan.h"
#include "zsb_walk.h"

// ======================================================================================= scan
// (the walk over one frame -- magic, header, block headers, checksum -- is zsb_walk.h, shared with the device scanner)

// ZSB_REFERENCE_QUIRKS only.  The reference parses a block's sections -- literals header, Huffman tree, sequences header, FSE tables --
// inside Block::parse, i.e. during the walk (ZStandard::parse is eager, frame.rs:210-223), while this library leaves them to the GPU.
// The difference shows in ONE place: when the walk itself fails (a truncated block, a missing checksum ...), a section error of an
// EARLIER block is what the reference has already returned.  So on a failed walk the blocks read so far are parsed here, on the host,
// in the reference's order (LiteralsSection::parse literals.rs:88-133 incl. HuffmanDecoder::parse, then Sequences::parse
// sequences.rs:52-143 incl. FseTable::parse per table); a well-formed container never pays for it.
static int eager_sections(const uint8_t *src, size_t n, const zsb_block &blk, uint32_t flags, uint64_t &ea, uint64_t &eb) {
    if (blk.type != ZSB_BT_COMPRESSED) return ZSB_OK;
    ZsbBlockWork w;
    parse_block(src, blk, w, flags);
    ea = w.err_a; eb = w.err_b;
    if (w.status != ZSB_OK && w.parse_stage == 0) return w.status >= ZSB_E_CORRUPT ? ZSB_OK : w.status;      // failed inside the literals section
    if (w.lit_type == ZSB_LT_COMPRESSED) {                                           // HuffmanDecoder::parse huffman.rs:80-130, from_weights :177-203
        static thread_local uint8_t weights[260]; static thread_local uint32_t ftbl[512]; static thread_local uint16_t lut[1 << ZSB_HUF_MAX_BITS];
        int16_t cnt[ZSB_HUF_WEIGHT_SYMS]; uint32_t rank[16]; int nw = 0, mb = 0; uint32_t dl = 0;
        int rc = huf_read_weights(src + w.huf_desc, w.huf_desc_end - w.huf_desc, weights, 1, nw, dl, ftbl, 1, cnt, 1, n - w.huf_desc, true);
        if (!rc) rc = huf_build_lut(weights, 1, nw, lut, rank, 1, mb, nullptr, true);
        if (rc && rc < ZSB_E_CORRUPT) { ea = eb = 0; return rc; }      // (codes >= 100 are this library's own limits, e.g. a tree deeper than 11: the reference parses on)
    }
    // FseTable::parse = parse_fse_table + from_distribution, table by table: a table read before the point of failure is also built
    const int tables_read = w.parse_stage >= 3 ? w.parse_stage - 2 : 0;
    for (int t = 0; t < tables_read; t++) {
        if (w.mode[t] != ZSB_M_FSE) continue;
        static thread_local int16_t cnt[256]; static thread_local uint32_t tbl[512];
        FwdBits f; fwd_init(f, src + w.tbl_desc[t], w.tbl_end - w.tbl_desc[t]);
        int al = 0, nsym = 0;
        int rc = fse_read_ncount(f, cnt, 1, 256, al, nsym);
        if (!rc) rc = fse_build_table(cnt, 1, nsym, al, tbl, 1, 3);
        if (rc && rc < ZSB_E_CORRUPT) { ea = eb = 0; return rc; }
    }
    return (w.status == ZSB_E_NULL_BYTE || w.status >= ZSB_E_CORRUPT) ? ZSB_OK : w.status;      // (the end mark of the bitstream is looked at when the block is decoded, parsing.rs:204)
}

// FseTable::from_distribution (fse.rs:110-202) with full-width fields, for zsb_block_sections (the decode's cells keep 6 bits of the symbol)
static int fse_states(const int16_t *cnt, int nsym, int al, zsb_fse_state *st) {
    const int N = 1 << al;
    std::vector<int> sym(N, -1), next(nsym > 0 ? nsym : 1, 0);
    int high = N - 1;
    for (int s = 0; s < nsym; s++) if (cnt[s] == -1) { if (high < 0) return ZSB_E_CORRUPTED_TABLE; sym[high--] = s; next[s] = 1; }
    const int step = (N >> 1) + (N >> 3) + 3, mask = N - 1;
    int pos = 0, placed = 0;
    for (int s = 0; s < nsym; s++) {
        if (cnt[s] <= 0) continue;
        next[s] = cnt[s];
        for (int k = 0; k < cnt[s]; k++) {
            if (placed >= high + 1) return ZSB_E_CORRUPTED_TABLE;
            sym[pos] = s; placed++;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    if (placed != high + 1) return ZSB_E_CORRUPTED_TABLE;
    for (int i = 0; i < N; i++) {
        const int s = sym[i];
        const uint32_t nx = (uint32_t)next[s]++;
        const uint32_t nb = (uint32_t)(al - zsb_flog2(nx));
        st[i].output = (uint16_t)s; st[i].bits_to_read = (uint16_t)nb; st[i].baseline = (uint16_t)((nx << nb) - (uint32_t)N);
    }
    return ZSB_OK;
}

extern "C" int zsb_block_sections(const uint8_t *src, size_t n, const zsb_block *blk, uint32_t flags, zsb_sections *out) {
    if (!src || !blk || !out || blk->type != ZSB_BT_COMPRESSED || blk->src_off + blk->size > n) return ZSB_E_ARG;
    memset(out, 0, sizeof *out);
    auto fail = [&](int rc, uint64_t a, uint64_t b) { out->status = rc; out->err_a = (uint32_t)a; out->err_b = (uint32_t)b; return rc; };
    ZsbBlockWork w;
    parse_block(src, *blk, w, flags);
    if (w.status != ZSB_OK && w.parse_stage == 0) return fail(w.status, w.err_a, w.err_b);
    out->lit_type = w.lit_type; out->regenerated_size = w.lit_regen;
    if (w.lit_type == ZSB_LT_RAW) { out->lit_data_off = w.lit_src; out->lit_data_len = w.lit_regen; }
    else if (w.lit_type == ZSB_LT_RLE) out->rle_byte = src[w.lit_src];
    else {
        for (int k = 0; k < 4; k++) out->jump_table[k] = (uint16_t)w.stream_size[k];
        const uint64_t lend = w.lit_type == ZSB_LT_COMPRESSED ? w.huf_desc_end : 0;
        if (w.lit_type == ZSB_LT_COMPRESSED) {
            static thread_local uint8_t weights[260]; static thread_local uint32_t ftbl[512]; static thread_local uint16_t lut[1 << ZSB_HUF_MAX_BITS];
            int16_t cnt[ZSB_HUF_WEIGHT_SYMS]; uint32_t rank[16]; int nw = 0, mb = 0; uint32_t dl = 0;
            int rc = huf_read_weights(src + w.huf_desc, w.huf_desc_end - w.huf_desc, weights, 1, nw, dl, ftbl, 1, cnt, 1, n - w.huf_desc, (flags & ZSB_REFERENCE_QUIRKS) != 0);
            if (!rc) rc = huf_build_lut(weights, 1, nw, lut, rank, 1, mb, nullptr, (flags & ZSB_REFERENCE_QUIRKS) != 0);
            if (rc) return fail(rc, 0, 0);
            out->max_bits = (uint8_t)mb;
            for (uint32_t i = 0; i < (1u << mb); i++) {
                if (lut[i] == ZSB_HUF_ABSENT) continue;
                const uint32_t sy = lut[i] & 0xFFu, nb = lut[i] >> 8;
                if (!out->code_len[sy]) { out->code_len[sy] = (uint8_t)nb; out->code[sy] = (uint16_t)(i >> (mb - nb)); }
            }
            out->lit_data_off = w.lit_src; out->lit_data_len = lend - w.lit_src;
        } else {
            // treeless: `data` runs to the end of the compressed literals (Compressed_Size bytes from the end of the section header)
            uint64_t total = 0; for (int k = 0; k < 4; k++) total += w.stream_size[k];
            out->lit_data_off = w.lit_src; out->lit_data_len = total;
        }
    }
    if (w.status != ZSB_OK && w.parse_stage < 3) return fail(w.status, w.err_a, w.err_b);
    out->number_of_sequences = w.nseq;
    for (int t = 0; t < 3; t++) { out->mode[t] = w.nseq ? (uint8_t)((w.raw_modes >> (6 - 2 * t)) & 3) : (uint8_t)ZSB_M_REPEAT; out->rle_symbol[t] = w.rle_sym[t]; }
    const int tables_read = w.nseq == 0 ? 0 : w.parse_stage >= 3 ? w.parse_stage - 2 : 0;
    for (int t = 0; t < tables_read; t++) {
        if (out->mode[t] != ZSB_M_FSE) continue;
        static thread_local int16_t cnt[256];
        FwdBits f; fwd_init(f, src + w.tbl_desc[t], w.tbl_end - w.tbl_desc[t]);
        int al = 0, nsym = 0;
        int rc = fse_read_ncount(f, cnt, 1, 256, al, nsym);
        if (!rc) rc = fse_states(cnt, nsym, al, out->table[t]);
        if (rc) return fail(rc, 0, 0);
        out->al[t] = (uint8_t)al;
    }
    if (w.status != ZSB_OK) return fail(w.status, w.err_a, w.err_b);
    if (w.nseq) { out->bitstream_off = w.bs_off; out->bitstream_len = w.bs_len; }
    return ZSB_OK;
}

// One frame (FrameIterator::next frame.rs:94-99): appends its descriptor (and its blocks) and returns true, or appends the
// failed frame (no blocks, status = the error) and returns false; false with nothing appended when the input is exhausted.
bool ZsbScanner::next() {
    if (done) return false;
    if (n - pos == 0) { done = true; return false; }
    const bool quirks = (flags & ZSB_REFERENCE_QUIRKS) != 0;
    ZsbWalkErr e{ZSB_OK, 0, 0};
    zsb_frame f; memset(&f, 0, sizeof f);
    f.first_block = (uint32_t)blocks.size();
    uint64_t end = pos; uint32_t n_emitted = 0;
    const bool ok = zsb_walk_frame(src, (uint64_t)n, (uint64_t)pos, flags, max_window, (uint32_t)frames.size(), f, end, e, n_emitted,
                                   [&](const zsb_block &b) { blocks.push_back(b); });
    if (!ok) {
        if (quirks && f.kind == 0 && blocks.size() > f.first_block) {
            for (size_t i = f.first_block; i < blocks.size(); i++) {
                uint64_t a = 0, b = 0;
                const int rc = eager_sections(src, n, blocks[i], flags, a, b);
                if (rc != ZSB_OK) { e.code = rc; e.a = a; e.b = b; break; }
            }
        }
        blocks.resize(f.first_block);                            // a failed frame contributes no blocks
        f.n_blocks = 0; f.status = e.code; f.src_len = n - f.src_off;
        frames.push_back(f);
        code = e.code; err_a = e.a; err_b = e.b; done = true;
        return false;
    }
    f.n_blocks = (uint32_t)blocks.size() - f.first_block; f.src_len = end - f.src_off; f.status = ZSB_OK;
    frames.push_back(f);
    pos = (size_t)end;
    return true;
}

// the scanner's arrays as malloc'd copies (zsb_free)
int ZsbScanner::release(zsb_frame **frames_out, size_t *n_frames, zsb_block **blocks_out, size_t *n_blocks) const {
    zsb_frame *fo = (zsb_frame *)malloc(sizeof(zsb_frame) * (frames.size() + 1));
    zsb_block *bo = (zsb_block *)malloc(sizeof(zsb_block) * (blocks.size() + 1));
    if (!fo || !bo) { free(fo); free(bo); return ZSB_E_NOMEM; }
    if (!frames.empty()) memcpy(fo, frames.data(), sizeof(zsb_frame) * frames.size());
    if (!blocks.empty()) memcpy(bo, blocks.data(), sizeof(zsb_block) * blocks.size());
    *frames_out = fo; *n_frames = frames.size(); *blocks_out = bo; *n_blocks = blocks.size();
    return ZSB_OK;
}

extern "C" int zsb_scan(const uint8_t *src, size_t n, uint32_t flags, uint64_t max_window,
                        zsb_frame **frames_out, size_t *n_frames, zsb_block **blocks_out, size_t *n_blocks,
                        uint64_t *err_a, uint64_t *err_b) {
    if (!frames_out || !n_frames || !blocks_out || !n_blocks || (!src && n)) return ZSB_E_ARG;
    ZsbScanner sc(src, n, flags, max_window);
    while (sc.next()) {}
    const int rc = sc.release(frames_out, n_frames, blocks_out, n_blocks);
    if (rc) return rc;
    if (err_a) *err_a = sc.err_a;
    if (err_b) *err_b = sc.err_b;
    return sc.code;
}
extern "C" void zsb_free(void *p) { free(p); }


// ======================================================================================= sharding by frame
// Frames are independent (a fresh DecodingContext per frame, frame.rs:233), so a buffer shards over GPUs by contiguous
// frame ranges with no exchange step.  Ranges are balanced on decompressed bytes: Frame_Content_Size where the header
// declares it, else 2.4 x the compressed size (text at level 3); skippable frames weigh their payload.
extern "C" int zsb_shard_plan(const zsb_frame *frames, size_t n_frames, int n_shards, size_t *first) {
    if ((!frames && n_frames) || n_shards <= 0 || !first) return ZSB_E_ARG;
    std::vector<double> cum(n_frames + 1, 0.0);
    for (size_t f = 0; f < n_frames; f++) {
        const zsb_frame &fr = frames[f];
        const double wgt = fr.kind == 0 ? (fr.has_content_size ? (double)fr.content_size : 2.4 * (double)fr.src_len) : (double)fr.src_len;
        cum[f + 1] = cum[f] + wgt + 1.0;     // +1: empty frames still cost a descriptor
    }
    first[0] = 0;
    size_t f = 0;
    for (int s = 1; s < n_shards; s++) {
        const double target = cum[n_frames] * (double)s / (double)n_shards;
        while (f < n_frames && cum[f + 1] <= target) f++;
        // the frame that crosses the target goes to the side it mostly lies on
        if (f < n_frames && target - cum[f] > cum[f + 1] - target) f++;
        if (f < first[s - 1]) f = first[s - 1];
        first[s] = f;
    }
    first[n_shards] = n_frames;
    return ZSB_OK;
}

// Descriptors of frames [f0, f1) rebased so that they describe the sub-buffer src[*src_off, *src_off + *src_len) on its own:
// what one rank uploads and hands to zsb_decode.  Arrays are malloc'd; free with zsb_free.
extern "C" int zsb_shard_extract(const zsb_frame *frames, size_t n_frames, const zsb_block *blocks, size_t n_blocks, size_t f0, size_t f1,
                                 zsb_frame **out_frames, zsb_block **out_blocks, size_t *out_n_blocks, uint64_t *src_off, uint64_t *src_len) {
    if (!frames || f0 > f1 || f1 > n_frames || !out_frames || !out_blocks || !out_n_blocks || !src_off || !src_len) return ZSB_E_ARG;
    *out_frames = nullptr; *out_blocks = nullptr; *out_n_blocks = 0; *src_off = 0; *src_len = 0;
    if (f0 == f1) return ZSB_OK;
    const uint64_t base = frames[f0].src_off, end = frames[f1 - 1].src_off + frames[f1 - 1].src_len;
    const size_t b0 = frames[f0].first_block, b1 = (size_t)frames[f1 - 1].first_block + frames[f1 - 1].n_blocks;
    if (b1 > n_blocks || b0 > b1) return ZSB_E_ARG;
    zsb_frame *of = (zsb_frame *)malloc(sizeof(zsb_frame) * (f1 - f0));
    zsb_block *ob = (zsb_block *)malloc(sizeof(zsb_block) * (b1 - b0 + 1));
    if (!of || !ob) { free(of); free(ob); return ZSB_E_NOMEM; }
    for (size_t f = f0; f < f1; f++) { of[f - f0] = frames[f]; of[f - f0].src_off -= base; of[f - f0].first_block -= (uint32_t)b0; }
    for (size_t b = b0; b < b1; b++) { ob[b - b0] = blocks[b]; ob[b - b0].src_off -= base; ob[b - b0].frame -= (uint32_t)f0; }
    *out_frames = of; *out_blocks = ob; *out_n_blocks = b1 - b0; *src_off = base; *src_len = end - base;
    return ZSB_OK;
}
