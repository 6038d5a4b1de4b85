// zsb_kernels.cu -- the sm_100a kernels of the Zstandard decode path.
//
// Pipeline of one batch (all frames of the batch move through each stage together):
//
//   k_parse   1 lane / block      section headers                       (literals.rs:135-206, sequences.rs:52-143)
//   k_plan1   1 CTA               per-frame table/tree chaining + scratch placement (scan) + work lists
//   k_huf     1 lane / stream     Huffman weights -> LUT (smem) -> 4-stream literal decode   (huffman.rs, literals.rs:49-86)
//   k_seq     1 lane / block      FSE tables (interleaved smem) + 3-state sequence decode     (fse.rs, sequence.rs, sequences.rs:191-237)
//   k_plan2   1 CTA               block/frame output offsets (scan), repeat-offset history, size checks
//   k_rawrle  1 CTA / block       raw / RLE block expansion, skippable payloads              (block.rs:76-79)
//   k_exec    1 CTA / frame       sequence execution in a 128 KiB shared-memory block image   (decoding_context.rs:78-106)
//   k_xxh     4 lanes / frame     XXH64 content checksum                                      (frame.rs:239-259)
//
// Nothing here is a dense contraction: no tensor cores.  The entropy stages are serial per stream,
// so they run lane-per-stream with all tables in shared memory; the execution stage is the only one
// that moves bulk data and keeps the whole block on chip, writing HBM once with 16-byte stores.
#include <cuda_runtime.h>
#include "zsb_kernels.h"
#include "zsb_parse.h"
#include "zsb_huf.h"

#define FULL 0xFFFFFFFFu

// ======================================================================================= k_parse
__global__ void __launch_bounds__(128) k_parse(const uint8_t *__restrict__ src, const zsb_block *__restrict__ blocks,
                                               ZsbBlockWork *__restrict__ work, uint32_t nb, uint32_t flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    ZsbBlockWork w;
    parse_block(src, blocks[i], w, flags);
    work[i] = w;
}

// ======================================================================================= CTA scan helper
// exclusive scan of one value per thread over a 1024-thread CTA; returns the exclusive prefix and the CTA total
__device__ __forceinline__ uint64_t cta_scan_excl(uint64_t v, uint64_t *s_warp, uint64_t &total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(FULL, inc, d); if (lane >= (uint32_t)d) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint64_t x = (lane < (blockDim.x >> 5)) ? s_warp[lane] : 0, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(FULL, xi, d); if (lane >= (uint32_t)d) xi += t; }
        s_warp[lane] = xi - x;            // exclusive warp offsets
        if (lane == 31) s_warp[32] = xi;  // total
    }
    __syncthreads();
    uint64_t r = s_warp[warp] + inc - v;
    total = s_warp[32];
    __syncthreads();
    return r;
}

// ======================================================================================= k_plan1
__global__ void __launch_bounds__(1024) k_plan1(const zsb_frame *__restrict__ frames, uint32_t nf, const zsb_block *__restrict__ blocks,
                                                uint32_t nb, ZsbBlockWork *work, ZsbFrameOut *fout, uint32_t *huf_list, uint32_t *seq_list,
                                                ZsbCounters *cnt, uint64_t lit_cap, uint64_t seq_cap, uint32_t flags) {
    __shared__ uint64_t s_warp[33];
    __shared__ uint64_t s_base[4];
    const uint32_t tid = threadIdx.x;
    // (a) per-frame chaining of Huffman tables and table modes
    for (uint32_t f = tid; f < nf; f += blockDim.x) {
        ZsbFrameOut o; o.dst_off = 0; o.dst_len = 0; o.xxh64 = 0; o.err_a = 0; o.err_b = 0; o.pad = 0;
        o.status = frames[f].status;
        if (o.status == ZSB_OK && frames[f].kind == 0) o.status = chain_frame(frames[f], blocks, work, flags, o.err_a, o.err_b);
        fout[f] = o;
    }
    if (tid < 4) s_base[tid] = 0;
    __syncthreads();
    // (b) scratch placement and work lists: four scans over the blocks
    for (uint32_t i0 = 0; i0 < nb; i0 += blockDim.x) {
        const uint32_t i = i0 + tid;
        uint64_t lit_need = 0, seq_need = 0, hf = 0, sf = 0;
        if (i < nb && blocks[i].type == ZSB_BT_COMPRESSED && work[i].status == ZSB_OK) {
            if (work[i].lit_type >= ZSB_LT_COMPRESSED) { lit_need = ((uint64_t)work[i].lit_regen + 15) & ~15ull; hf = 1; }
            if (work[i].nseq) { seq_need = work[i].nseq; sf = 1; }
        }
        uint64_t t0, t1, t2, t3;
        uint64_t a = cta_scan_excl(lit_need, s_warp, t0), b = cta_scan_excl(seq_need, s_warp, t1);
        uint64_t c = cta_scan_excl(hf, s_warp, t2), d = cta_scan_excl(sf, s_warp, t3);
        if (i < nb) {
            if (hf) { work[i].lit_buf = s_base[0] + a; huf_list[s_base[2] + c] = i; }
            if (sf) { work[i].seq_buf = s_base[1] + b; seq_list[s_base[3] + d] = i; }
        }
        __syncthreads();
        if (tid == 0) { s_base[0] += t0; s_base[1] += t1; s_base[2] += t2; s_base[3] += t3; }
        __syncthreads();
    }
    if (tid == 0) {
        cnt->lit_total = s_base[0]; cnt->seq_total = s_base[1];
        cnt->n_huf = (uint32_t)s_base[2]; cnt->n_seq = (uint32_t)s_base[3];
        cnt->overflow = (s_base[0] > lit_cap || s_base[1] > seq_cap) ? 1u : 0u;
    }
}

// ======================================================================================= k_huf
// One warp per CTA, 8 blocks per warp: lane = 4 * slot + stream.  Per slot: LUT (4 KiB, aliased with the
// weight FSE table while the weights are being decoded), weights, counts, ranks.
#define HUF_SLOTS 8
#define HUF_LUT_BYTES (2u << ZSB_HUF_MAX_BITS)   // 4096
struct HufSlot {
    union { uint16_t lut[1 << ZSB_HUF_MAX_BITS]; uint32_t ftbl[512]; } u;
    uint8_t weights[260];
    int16_t cnt[ZSB_HUF_WEIGHT_SYMS];
    uint32_t rank[16];
    int maxbits;
    int status;
};
__global__ void __launch_bounds__(32) k_huf(const uint8_t *__restrict__ src, uint64_t src_len, ZsbBlockWork *work,
                                            const uint32_t *__restrict__ huf_list, const ZsbCounters *__restrict__ cnt,
                                            uint8_t *lit_pool, uint32_t flags) {
    __shared__ HufSlot slots[HUF_SLOTS];
    if (cnt->overflow) return;
    const uint32_t lane = threadIdx.x, slot = lane >> 2, stream = lane & 3;
    const uint32_t n = cnt->n_huf, idx = blockIdx.x * HUF_SLOTS + slot;
    if (blockIdx.x * HUF_SLOTS >= n) return;
    const bool active = idx < n;
    const uint32_t bi = active ? huf_list[idx] : 0;
    HufSlot &S = slots[slot];
    if (active && stream == 0) {
        const ZsbBlockWork &w = work[bi];
        int nw = 0; uint32_t dl = 0;
        int rc = huf_read_weights(src + w.huf_desc, w.huf_desc_end - w.huf_desc, S.weights, 1, nw, dl, S.u.ftbl, 1, S.cnt, 1,
                                  src_len - w.huf_desc, (flags & ZSB_REFERENCE_QUIRKS) != 0);
        int mb = 0;
        if (!rc) rc = huf_build_lut(S.weights, 1, nw, S.u.lut, S.rank, 1, mb, nullptr);
        S.maxbits = mb; S.status = rc;
    }
    __syncwarp();
    int rc = 0;
    if (active) {
        rc = S.status;
        const ZsbBlockWork &w = work[bi];
        if (!rc && stream < w.n_streams) {
            const uint32_t regen = w.lit_regen;
            uint32_t seg, expect, ooff; uint64_t start = w.lit_src;
            if (w.n_streams == 1) { expect = regen; ooff = 0; }
            else {
                seg = (regen + 3) / 4; ooff = stream * seg; expect = stream < 3 ? seg : regen - 3 * seg;
                for (uint32_t k = 0; k < stream; k++) start += w.stream_size[k];
            }
            rc = huf_decode_stream(src, start, start + w.stream_size[stream], src_len, S.u.lut, S.maxbits, lit_pool + w.lit_buf + ooff, expect);
        }
    }
    // first failing stream of the block decides its status
    const int r1 = __shfl_sync(FULL, rc, (lane & ~3u) + 1), r2 = __shfl_sync(FULL, rc, (lane & ~3u) + 2), r3 = __shfl_sync(FULL, rc, (lane & ~3u) + 3);
    if (active && stream == 0) {
        int st = rc ? rc : r1 ? r1 : r2 ? r2 : r3;
        if (st) work[bi].status = st;
    }
}

// ======================================================================================= k_seq
// One warp per CTA, one block per lane.  Tables are interleaved across lanes (cell i of lane l at word
// i*32 + l), so the three table reads of every decode step are bank-conflict free.
#define SEQ_TBL_CELLS 512
#define SEQ_SMEM_BYTES (3 * SEQ_TBL_CELLS * 32 * 4 + 256 * 32 * 2 + 96 * 4)
__global__ void __launch_bounds__(32, 1) k_seq(const uint8_t *__restrict__ src, uint64_t src_len, ZsbBlockWork *work,
                                               const uint32_t *__restrict__ seq_list, const ZsbCounters *__restrict__ cnt,
                                               uint64_t *seq_pool) {
    extern __shared__ __align__(128) uint8_t smem[];
    if (cnt->overflow) return;
    uint32_t *tbl = reinterpret_cast<uint32_t *>(smem);
    int16_t *counts = reinterpret_cast<int16_t *>(smem + 3 * SEQ_TBL_CELLS * 32 * 4);
    uint32_t *bases = reinterpret_cast<uint32_t *>(smem + 3 * SEQ_TBL_CELLS * 32 * 4 + 256 * 32 * 2);   // [0..35] LL, [36..88] ML
    const uint32_t lane = threadIdx.x;
    const uint32_t n = cnt->n_seq, idx = blockIdx.x * 32 + lane;
    if (blockIdx.x * 32 >= n) return;
    for (uint32_t k = lane; k < 36 + 53; k += 32) bases[k] = k < 36 ? zsb_ll_base(k) : zsb_ml_base(k - 36);
    __syncwarp();
    if (idx >= n) return;
    const uint32_t bi = seq_list[idx];
    ZsbBlockWork w = work[bi];
    if (w.status != ZSB_OK) return;            // a literal stream of this block already failed
    SeqTables T;
    T.ts = 32;
    T.tbl[0] = tbl + lane; T.tbl[1] = tbl + SEQ_TBL_CELLS * 32 + lane; T.tbl[2] = tbl + 2 * SEQ_TBL_CELLS * 32 + lane;
    int rc = seq_build_tables(src, w, T, counts + lane, 32);
    if (!rc) rc = seq_decode(src, src_len, w, T, bases, bases + 36, seq_pool + w.seq_buf);
    if (rc) { work[bi].status = rc; return; }
    ZsbBlockWork &g = work[bi];
    g.out_size = w.out_size; g.lit_used = w.lit_used;
    g.rep_out[0] = w.rep_out[0]; g.rep_out[1] = w.rep_out[1]; g.rep_out[2] = w.rep_out[2];
}

// ======================================================================================= k_plan2
__global__ void __launch_bounds__(1024) k_plan2(const zsb_frame *__restrict__ frames, uint32_t nf, const zsb_block *__restrict__ blocks,
                                                ZsbBlockWork *work, ZsbFrameOut *fout, ZsbCounters *cnt, uint64_t dst_cap, uint32_t flags) {
    __shared__ uint64_t s_warp[33];
    __shared__ uint64_t s_base, s_max;
    const uint32_t tid = threadIdx.x;
    if (cnt->overflow) return;
    if (tid == 0) { s_base = 0; s_max = 0; }
    __syncthreads();
    for (uint32_t f0 = 0; f0 < nf; f0 += blockDim.x) {
        const uint32_t f = f0 + tid;
        uint64_t len = 0; int st = ZSB_OK;
        if (f < nf) {
            st = fout[f].status;
            if (st == ZSB_OK) {
                if (frames[f].kind == 1) len = (flags & ZSB_PRINT_SKIPPABLE) ? blocks[frames[f].first_block].size : 0;   // main.rs:45-49
                else {
                    st = plan_frame(frames[f], blocks, work, len);
                    if (st == ZSB_OK && frames[f].has_content_size && len != frames[f].content_size && !(flags & ZSB_REFERENCE_QUIRKS)) st = ZSB_E_CONTENT_SIZE;
                    if (st != ZSB_OK) len = 0;
                }
            }
        }
        uint64_t tot;
        uint64_t off = cta_scan_excl(len, s_warp, tot) + s_base;
        if (f < nf) {
            if (st == ZSB_OK && off + len > dst_cap) { st = ZSB_E_DST_TOO_SMALL; }
            fout[f].dst_off = off; fout[f].dst_len = (st == ZSB_OK) ? len : 0; fout[f].status = st;
            if (st == ZSB_OK && len) atomicMax((unsigned long long *)&s_max, (unsigned long long)(off + len));
        }
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    if (tid == 0) cnt->dst_total = s_max;
}

// ======================================================================================= copy helpers
// CTA-cooperative copy global -> global of n bytes, any alignment
__device__ __forceinline__ void cta_copy_g2g(uint8_t *dst, const uint8_t *src, uint64_t n) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    if ((((uintptr_t)dst ^ (uintptr_t)src) & 15) == 0) {
        uint64_t head = (16 - ((uintptr_t)dst & 15)) & 15; if (head > n) head = n;
        for (uint64_t i = tid; i < head; i += nt) dst[i] = src[i];
        const uint64_t nv = (n - head) >> 4;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head); uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
        for (uint64_t i = tid; i < nv; i += nt) d4[i] = __ldg(s4 + i);
        for (uint64_t i = head + (nv << 4) + tid; i < n; i += nt) dst[i] = src[i];
    } else {
        for (uint64_t i = tid; i < n; i += nt) dst[i] = src[i];
    }
}
__device__ __forceinline__ void cta_fill_g(uint8_t *dst, uint8_t b, uint64_t n) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    uint64_t head = (16 - ((uintptr_t)dst & 15)) & 15; if (head > n) head = n;
    for (uint64_t i = tid; i < head; i += nt) dst[i] = b;
    const uint64_t nv = (n - head) >> 4;
    const uint32_t w = b * 0x01010101u; const uint4 v = make_uint4(w, w, w, w);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
    for (uint64_t i = tid; i < nv; i += nt) d4[i] = v;
    for (uint64_t i = head + (nv << 4) + tid; i < n; i += nt) dst[i] = b;
}

// ======================================================================================= k_rawrle
__global__ void __launch_bounds__(256) k_rawrle(const uint8_t *__restrict__ src, const zsb_block *__restrict__ blocks,
                                                const ZsbBlockWork *__restrict__ work, const ZsbFrameOut *__restrict__ fout,
                                                const uint32_t *__restrict__ list, const ZsbCounters *__restrict__ cnt, uint8_t *dst) {
    if (cnt->overflow) return;
    const uint32_t bi = list[blockIdx.x];
    const zsb_block b = blocks[bi];
    const ZsbFrameOut fo = fout[b.frame];
    if (fo.status != ZSB_OK || fo.dst_len == 0 || b.size == 0) return;
    uint8_t *d = dst + fo.dst_off + work[bi].out_off;
    if (b.type == ZSB_BT_RLE) cta_fill_g(d, src[b.src_off], b.size);       // block.rs:77-79
    else cta_copy_g2g(d, src + b.src_off, b.size);                         // block.rs:76 ; skippable payload frame.rs:81
}

// ======================================================================================= k_exec
#define EXEC_THREADS 512
#define EXEC_WARPS (EXEC_THREADS / 32)
#define EXEC_LIT_STAGE 65536u
#define EXEC_LONG 32u
#define EXEC_OUT_BYTES (ZSB_BLOCK_MAX + 16)
#define EXEC_BM_WORDS (ZSB_BLOCK_MAX / 32)
#define EXEC_SMEM_BYTES (EXEC_OUT_BYTES + EXEC_BM_WORDS * 4 + EXEC_LIT_STAGE + 16 + 16)

struct LitSrc { const uint8_t *s; const uint8_t *g; uint32_t rle; int mode; };   // mode 0 smem, 1 global, 2 rle byte
__device__ __forceinline__ uint8_t lit_at(const LitSrc &L, uint32_t i) {
    return L.mode == 0 ? L.s[i] : (L.mode == 1 ? __ldg(L.g + i) : (uint8_t)L.rle);
}
__device__ __forceinline__ uint32_t range_mask(uint32_t w, uint32_t a, uint32_t e) {
    const uint32_t w0 = w << 5;
    const uint32_t lo = a > w0 ? a - w0 : 0u, hi = e < w0 + 32 ? e - w0 : 32u;
    return zsb_shl32(0xFFFFFFFFu, lo) & ~zsb_shl32(0xFFFFFFFFu, hi);
}
// all bytes of [a, e) already written?
__device__ __forceinline__ bool range_ready(const volatile uint32_t *bm, uint32_t a, uint32_t e) {
    bool ok = true;
    for (uint32_t w = a >> 5; w <= ((e - 1) >> 5); w++) { const uint32_t m = range_mask(w, a, e); ok = ok && ((bm[w] & m) == m); }
    return ok;
}
__device__ __forceinline__ void range_publish(uint32_t *bm, uint32_t a, uint32_t e) {
    for (uint32_t w = a >> 5; w <= ((e - 1) >> 5); w++) atomicOr(&bm[w], range_mask(w, a, e));
}

// One batch of 32 consecutive sequences, one lane per sequence.
__device__ __forceinline__ void exec_batch(uint32_t batch, uint32_t nseq, const uint64_t *__restrict__ seqs, uint8_t *o, uint32_t *bm,
                                           const LitSrc &L, const uint32_t *rep_in, uint64_t P0, const uint8_t *gblk, int *s_err) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t s = batch * 32 + lane;
    const bool valid = s < nseq;
    const uint64_t rec = valid ? __ldg(seqs + s) : 0ull;
    uint64_t prev = __shfl_up_sync(FULL, rec, 1);
    if (lane == 0) prev = batch ? __ldg(seqs + s - 1) : 0ull;
    const uint32_t out_start = (uint32_t)prev & ZSB_REC_POS_MASK, lit_start = (uint32_t)(prev >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK;
    const uint32_t out_end = (uint32_t)rec & ZSB_REC_POS_MASK, lit_end = (uint32_t)(rec >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK;
    const uint32_t ll = valid ? lit_end - lit_start : 0u;
    const uint32_t ml = valid ? out_end - out_start - ll : 0u;
    const uint32_t off = valid ? seq_real_offset((uint32_t)(rec >> (2 * ZSB_REC_POS_BITS)), rep_in) : 1u;
    const uint32_t dstm = out_start + ll;
    const bool bad = valid && (off == 0 || (uint64_t)off > P0 + dstm);            // decoding_context.rs:86-90
    if (bad) *s_err = ZSB_E_IMPOSSIBLE_VALUE;
    const int src = (int)dstm - (int)off;
    const bool longL = ll > EXEC_LONG, longM = ml > EXEC_LONG;

    // ---- literals: no dependency on earlier output (decoding_context.rs:92-93)
    if (ll && !longL)
        for (uint32_t k = 0; k < ll; k++) o[out_start + k] = lit_at(L, lit_start + k);
    for (uint32_t m = __ballot_sync(FULL, longL); m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        const uint32_t a = __shfl_sync(FULL, out_start, j), ls = __shfl_sync(FULL, lit_start, j), n = __shfl_sync(FULL, ll, j);
        for (uint32_t k = lane; k < n; k += 32) o[a + k] = lit_at(L, ls + k);
    }
    __syncwarp();
    // ---- matches: copy as soon as every source byte is known to be written (decoding_context.rs:95-98)
    bool pend = valid && ml && !bad;
    bool first = true;
    // Source bytes at or above out_start are this sequence's own literals (written above) or its own
    // match output; only [a0, e) below out_start comes from earlier sequences and must be awaited.
    const int e = min(src + (int)ml, (int)out_start);
    const uint32_t a0 = src > 0 ? (uint32_t)src : 0u;
    for (uint32_t spins = 0;; spins++) {
        if (spins > (1u << 20)) { *s_err = ZSB_E_CORRUPT; break; }   // watchdog: a dependency that never resolves is a bug, not a hang
        bool didm = false;
        if (pend && !longM) {
            const bool ready = (e <= (int)a0) || range_ready(bm, a0, (uint32_t)e);
            if (ready) {
                __threadfence_block();
                for (uint32_t k = 0; k < ml; k++) {
                    const int sp = src + (int)k;
                    o[dstm + k] = sp < 0 ? __ldcg(gblk + sp) : o[sp];
                }
                pend = false; didm = true;
            }
        }
        for (uint32_t m = __ballot_sync(FULL, pend && longM); m; m &= m - 1) {
            const int j = __ffs(m) - 1;
            const int js = __shfl_sync(FULL, src, j), je = __shfl_sync(FULL, e, j);
            const uint32_t jml = __shfl_sync(FULL, ml, j), jd = __shfl_sync(FULL, dstm, j), joff = __shfl_sync(FULL, off, j);
            const uint32_t ja = js > 0 ? (uint32_t)js : 0u;
            bool ok = true;
            if (je > (int)ja)
                for (uint32_t w = (ja >> 5) + lane; w <= (((uint32_t)je - 1) >> 5); w += 32) {
                    const uint32_t mk = range_mask(w, ja, (uint32_t)je);
                    ok = ok && ((((volatile uint32_t *)bm)[w] & mk) == mk);
                }
            ok = __all_sync(FULL, ok);
            if (ok) {
                __threadfence_block();
                // the copy is periodic with period off when the match overlaps its own output
                for (uint32_t k = lane; k < jml; k += 32) {
                    const int sp = js + (int)(joff >= jml ? k : k % joff);
                    o[jd + k] = sp < 0 ? __ldcg(gblk + sp) : o[sp];
                }
                if ((int)lane == j) { pend = false; didm = true; }
            }
        }
        // ---- publish what this round completed
        __syncwarp();
        __threadfence_block();
        uint32_t pa = 0, pe = 0;
        if (valid) {
            if (first) { pa = out_start; pe = (didm || bad) ? out_end : dstm; }
            else if (didm) { pa = dstm; pe = out_end; }
        }
        const bool plong = (pe - pa) > 4 * EXEC_LONG;
        if (pe > pa && !plong) range_publish(bm, pa, pe);
        for (uint32_t m = __ballot_sync(FULL, plong); m; m &= m - 1) {
            const int j = __ffs(m) - 1;
            const uint32_t ja = __shfl_sync(FULL, pa, j), je = __shfl_sync(FULL, pe, j);
            for (uint32_t w = (ja >> 5) + lane; w <= ((je - 1) >> 5); w += 32) atomicOr(&bm[w], range_mask(w, ja, je));
        }
        first = false;
        if (!__any_sync(FULL, pend)) break;
    }
}

__global__ void __launch_bounds__(EXEC_THREADS, 1) k_exec(const uint8_t *__restrict__ src, const zsb_frame *__restrict__ frames,
                                                          const zsb_block *__restrict__ blocks, const ZsbBlockWork *__restrict__ work,
                                                          ZsbFrameOut *fout, const uint32_t *__restrict__ exec_list,
                                                          const ZsbCounters *__restrict__ cnt, const uint64_t *__restrict__ seq_pool,
                                                          const uint8_t *__restrict__ lit_pool, uint8_t *dst) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_err;
    if (cnt->overflow) return;
    uint8_t *out_s = smem;
    uint32_t *bm = reinterpret_cast<uint32_t *>(smem + EXEC_OUT_BYTES);
    uint8_t *lit_s = smem + EXEC_OUT_BYTES + EXEC_BM_WORDS * 4;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t f = exec_list[blockIdx.x];
    const ZsbFrameOut fo = fout[f];
    if (fo.status != ZSB_OK) return;
    const zsb_frame fr = frames[f];
    uint8_t *fdst = dst + fo.dst_off;
    if (tid == 0) s_err = 0;
    __syncthreads();
    for (uint32_t k = 0; k < fr.n_blocks; k++) {
        const uint32_t bi = fr.first_block + k;
        if (blocks[bi].type != ZSB_BT_COMPRESSED) continue;
        const ZsbBlockWork &W = work[bi];
        const uint32_t out_size = W.out_size, nseq = W.nseq, regen = W.lit_regen;
        uint8_t *gblk = fdst + W.out_off;
        const uint32_t shift = (uint32_t)((uintptr_t)gblk & 15);
        uint8_t *o = out_s + shift;
        // literal source
        LitSrc L; L.rle = 0; L.s = nullptr;
        L.g = W.lit_type == ZSB_LT_RAW ? src + W.lit_src : lit_pool + W.lit_buf;
        if (W.lit_type == ZSB_LT_RLE) { L.mode = 2; L.rle = src[W.lit_src]; }
        else if (regen <= EXEC_LIT_STAGE) {
            L.mode = 0;
            const uint32_t ls = (uint32_t)((uintptr_t)L.g & 15);
            L.s = lit_s + ls;
            // stage the literals: aligned 16-byte loads once past the head
            uint32_t head = (16 - ls) & 15; if (head > regen) head = regen;
            for (uint32_t i = tid; i < head; i += EXEC_THREADS) lit_s[ls + i] = __ldg(L.g + i);
            const uint32_t nv = (regen - head) >> 4;
            const uint4 *g4 = reinterpret_cast<const uint4 *>(L.g + head); uint4 *s4 = reinterpret_cast<uint4 *>(lit_s + ls + head);
            for (uint32_t i = tid; i < nv; i += EXEC_THREADS) s4[i] = __ldg(g4 + i);
            for (uint32_t i = head + (nv << 4) + tid; i < regen; i += EXEC_THREADS) lit_s[ls + i] = __ldg(L.g + i);
        } else L.mode = 1;
        if (nseq) for (uint32_t i = tid; i < (out_size + 31) / 32; i += EXEC_THREADS) bm[i] = 0;
        __syncthreads();
        if (nseq == 0) {
            for (uint32_t i = tid; i < regen; i += EXEC_THREADS) o[i] = lit_at(L, i);        // literals-only block (RFC; reference: Q1)
        } else {
            const uint64_t *seqs = seq_pool + W.seq_buf;
            const uint64_t lastrec = __ldg(seqs + nseq - 1);
            const uint32_t oe = (uint32_t)lastrec & ZSB_REC_POS_MASK, le = (uint32_t)(lastrec >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK;
            for (uint32_t i = tid; i < regen - le; i += EXEC_THREADS) o[oe + i] = lit_at(L, le + i);   // decoding_context.rs:101-103
            const uint32_t nbatch = (nseq + 31) / 32;
            for (uint32_t b = warp; b < nbatch; b += EXEC_WARPS)
                exec_batch(b, nseq, seqs, o, bm, L, W.rep_in, fr.kind == 0 ? W.out_off : 0, gblk, &s_err);
        }
        __syncthreads();
        // flush the block image to HBM: congruent alignment -> 16-byte stores
        {
            uint32_t head = (16 - shift) & 15; if (head > out_size) head = out_size;
            for (uint32_t i = tid; i < head; i += EXEC_THREADS) gblk[i] = o[i];
            const uint32_t nv = (out_size - head) >> 4;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(o + head); uint4 *g4 = reinterpret_cast<uint4 *>(gblk + head);
            for (uint32_t i = tid; i < nv; i += EXEC_THREADS) g4[i] = s4[i];
            for (uint32_t i = head + (nv << 4) + tid; i < out_size; i += EXEC_THREADS) gblk[i] = o[i];
        }
        __syncthreads();
        if (s_err) { if (tid == 0) { fout[f].status = s_err; fout[f].dst_len = 0; } break; }
    }
}

// ======================================================================================= k_xxh
#define XP1 0x9E3779B185EBCA87ull
#define XP2 0xC2B2AE3D27D4EB4Full
#define XP3 0x165667B19E3779F9ull
#define XP4 0x85EBCA77C2B2AE63ull
#define XP5 0x27D4EB2F165667C5ull
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xround(uint64_t acc, uint64_t in) { return rotl64(acc + in * XP2, 31) * XP1; }
__device__ __forceinline__ uint64_t xmerge(uint64_t h, uint64_t v) { return (h ^ xround(0, v)) * XP1 + XP4; }
// 8 bytes at any alignment from two aligned words
__device__ __forceinline__ uint64_t ld64_any(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const unsigned long long *q = reinterpret_cast<const unsigned long long *>(a & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(a & 7) * 8;
    uint64_t w0 = __ldcg(q);
    if (sh == 0) return w0;
    uint64_t w1 = __ldcg(q + 1);
    return (w0 >> sh) | (w1 << (64 - sh));
}
// Four lanes per frame: lane q owns accumulator v(q+1) and reads 8 of every 32 bytes.
__global__ void __launch_bounds__(128) k_xxh(const uint8_t *__restrict__ dst, ZsbFrameOut *fout, const uint32_t *__restrict__ list,
                                             uint32_t n, const ZsbCounters *__restrict__ cnt) {
    if (cnt->overflow) return;
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x, gq = gt >> 2, q = gt & 3, lane = threadIdx.x & 31;
    const bool active = gq < n;
    const uint32_t f = active ? list[gq] : 0;
    const bool ok = active && fout[f].status == ZSB_OK;
    const uint8_t *p = dst + (ok ? fout[f].dst_off : 0);
    const uint64_t len = ok ? fout[f].dst_len : 0;
    uint64_t v = q == 0 ? XP1 + XP2 : q == 1 ? XP2 : q == 2 ? 0ull : 0ull - XP1;
    const uint64_t nstripes = len >> 5;
    const uint8_t *pp = p + 8 * q;
    uint64_t i = 0;
    for (; i + 4 <= nstripes; i += 4) {
        const uint64_t x0 = ld64_any(pp), x1 = ld64_any(pp + 32), x2 = ld64_any(pp + 64), x3 = ld64_any(pp + 96);
        v = xround(v, x0); v = xround(v, x1); v = xround(v, x2); v = xround(v, x3);
        pp += 128;
    }
    for (; i < nstripes; i++) { v = xround(v, ld64_any(pp)); pp += 32; }
    __syncwarp();
    const uint32_t qb = lane & ~3u;
    const uint64_t v1 = __shfl_sync(FULL, v, qb), v2 = __shfl_sync(FULL, v, qb + 1), v3 = __shfl_sync(FULL, v, qb + 2), v4 = __shfl_sync(FULL, v, qb + 3);
    if (ok && q == 0) {
        uint64_t h;
        if (len >= 32) {
            h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
            h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
        } else h = XP5;
        h += len;
        const uint8_t *t = p + (nstripes << 5), *end = p + len;
        while (t + 8 <= end) { h ^= xround(0, ld64_any(t)); h = rotl64(h, 27) * XP1 + XP4; t += 8; }
        if (t + 4 <= end) {
            uint32_t x = (uint32_t)__ldcg(t) | ((uint32_t)__ldcg(t + 1) << 8) | ((uint32_t)__ldcg(t + 2) << 16) | ((uint32_t)__ldcg(t + 3) << 24);
            h ^= (uint64_t)x * XP1; h = rotl64(h, 23) * XP2 + XP3; t += 4;
        }
        while (t < end) { h ^= (uint64_t)__ldcg(t) * XP5; h = rotl64(h, 11) * XP1; t++; }
        h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
        fout[f].xxh64 = h;
    }
}

// ======================================================================================= stage kernels (one lane)
__global__ void k_stage_fse(const uint8_t *desc, uint32_t n, int max_sym, const int16_t *dist_in, int ndist_in, int al_in,
                            int *res /* rc, al, nsym, consumed */, uint32_t *cells, int16_t *dist_out) {
    __shared__ int16_t cntbuf[256];
    int al = al_in, nsym = ndist_in, rc = 0; uint32_t consumed = 0;
    if (desc) {
        FwdBits f; fwd_init(f, desc, n);
        rc = fse_read_ncount(f, cntbuf, 1, max_sym, al, nsym);
        consumed = fwd_bytes_read(f);
    } else for (int i = 0; i < nsym; i++) cntbuf[i] = dist_in[i];
    if (!rc && dist_out) for (int i = 0; i < nsym && i < 256; i++) dist_out[i] = cntbuf[i];
    if (!rc) { if (al > ZSB_MAX_AL) rc = ZSB_E_LARGE_ACCURACY_LOG; else rc = fse_build_table(cntbuf, 1, nsym, al, cells, 1, 3); }
    res[0] = rc; res[1] = al; res[2] = nsym; res[3] = (int)consumed;
}
__global__ void k_stage_huf(const uint8_t *desc, uint32_t n, int *res /* rc, maxbits, consumed */, uint8_t *lens, uint16_t *lut_out) {
    __shared__ HufSlot S;
    int nw = 0, mb = 0; uint32_t dl = 0;
    int rc = huf_read_weights(desc, n, S.weights, 1, nw, dl, S.u.ftbl, 1, S.cnt, 1, n, false);
    if (!rc) rc = huf_build_lut(S.weights, 1, nw, S.u.lut, S.rank, 1, mb, lens);
    if (!rc) for (int i = 0; i < (1 << mb); i++) lut_out[i] = S.u.lut[i];
    res[0] = rc; res[1] = mb; res[2] = (int)dl;
}
// triples (ll, offset_value, ml) -> packed records for one block, history [1,4,8] handled by rep_in
__global__ void k_stage_records(const uint32_t *tri, uint32_t nseq, uint32_t nlit, uint64_t *rec, ZsbBlockWork *w) {
    uint32_t h0 = ZSB_OFF_SYM | (0u << 25), h1 = ZSB_OFF_SYM | (1u << 25), h2 = ZSB_OFF_SYM | (2u << 25);
    uint32_t out_end = 0, lit_end = 0; int err = 0;
    for (uint32_t i = 0; i < nseq && !err; i++) {
        const uint32_t ll = tri[3 * i], ov = tri[3 * i + 1], ml = tri[3 * i + 2];
        if (ov == 0) { err = ZSB_E_NULL_OFFSET; break; }                       // decoding_context.rs:52
        const uint32_t off = seq_resolve_offset(ov, ll, h0, h1, h2, err);
        lit_end += ll; out_end += ll + ml;
        if (lit_end > nlit) { err = ZSB_E_IMPOSSIBLE_VALUE; break; }
        if (out_end + (nlit - lit_end) > ZSB_BLOCK_MAX) { err = ZSB_E_BLOCK_TOO_LARGE; break; }
        rec[i] = (uint64_t)out_end | ((uint64_t)lit_end << ZSB_REC_POS_BITS) | ((uint64_t)off << (2 * ZSB_REC_POS_BITS));
    }
    w->status = err; w->lit_used = lit_end; w->out_size = out_end + (nlit - lit_end);
    w->rep_in[0] = 1; w->rep_in[1] = 4; w->rep_in[2] = 8;
}

// ======================================================================================= launchers
static cudaError_t set_smem(const void *fn, size_t bytes) {
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
cudaError_t zsbk_init() {
    cudaError_t e = set_smem((const void *)k_seq, SEQ_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return set_smem((const void *)k_exec, EXEC_SMEM_BYTES);
}
void zsbk_parse(cudaStream_t st, const uint8_t *src, const zsb_block *blocks, ZsbBlockWork *work, uint32_t nb, uint32_t flags) {
    if (nb) k_parse<<<(nb + 127) / 128, 128, 0, st>>>(src, blocks, work, nb, flags);
}
void zsbk_plan1(cudaStream_t st, const zsb_frame *frames, uint32_t nf, const zsb_block *blocks, uint32_t nb, ZsbBlockWork *work,
                ZsbFrameOut *fout, uint32_t *huf_list, uint32_t *seq_list, ZsbCounters *cnt, uint64_t lit_cap, uint64_t seq_cap, uint32_t flags) {
    k_plan1<<<1, 1024, 0, st>>>(frames, nf, blocks, nb, work, fout, huf_list, seq_list, cnt, lit_cap, seq_cap, flags);
}
void zsbk_huf(cudaStream_t st, uint32_t ncomp, const uint8_t *src, uint64_t src_len, ZsbBlockWork *work, const uint32_t *huf_list,
              const ZsbCounters *cnt, uint8_t *lit_pool, uint32_t flags) {
    if (ncomp) k_huf<<<(ncomp + HUF_SLOTS - 1) / HUF_SLOTS, 32, 0, st>>>(src, src_len, work, huf_list, cnt, lit_pool, flags);
}
void zsbk_seq(cudaStream_t st, uint32_t ncomp, const uint8_t *src, uint64_t src_len, ZsbBlockWork *work, const uint32_t *seq_list,
              const ZsbCounters *cnt, uint64_t *seq_pool) {
    if (ncomp) k_seq<<<(ncomp + 31) / 32, 32, SEQ_SMEM_BYTES, st>>>(src, src_len, work, seq_list, cnt, seq_pool);
}
void zsbk_plan2(cudaStream_t st, const zsb_frame *frames, uint32_t nf, const zsb_block *blocks, ZsbBlockWork *work, ZsbFrameOut *fout,
                ZsbCounters *cnt, uint64_t dst_cap, uint32_t flags) {
    k_plan2<<<1, 1024, 0, st>>>(frames, nf, blocks, work, fout, cnt, dst_cap, flags);
}
void zsbk_rawrle(cudaStream_t st, uint32_t n, const uint8_t *src, const zsb_block *blocks, const ZsbBlockWork *work, const ZsbFrameOut *fout,
                 const uint32_t *list, const ZsbCounters *cnt, uint8_t *dst) {
    if (n) k_rawrle<<<n, 256, 0, st>>>(src, blocks, work, fout, list, cnt, dst);
}
void zsbk_exec(cudaStream_t st, uint32_t n, const uint8_t *src, const zsb_frame *frames, const zsb_block *blocks, const ZsbBlockWork *work,
               ZsbFrameOut *fout, const uint32_t *exec_list, const ZsbCounters *cnt, const uint64_t *seq_pool, const uint8_t *lit_pool, uint8_t *dst) {
    if (n) k_exec<<<n, EXEC_THREADS, EXEC_SMEM_BYTES, st>>>(src, frames, blocks, work, fout, exec_list, cnt, seq_pool, lit_pool, dst);
}
void zsbk_xxh(cudaStream_t st, uint32_t n, const uint8_t *dst, ZsbFrameOut *fout, const uint32_t *list, const ZsbCounters *cnt) {
    if (n) k_xxh<<<(n * 4 + 127) / 128, 128, 0, st>>>(dst, fout, list, n, cnt);
}
void zsbk_stage_fse(cudaStream_t st, const uint8_t *desc, uint32_t n, int max_sym, const int16_t *dist_in, int ndist_in, int al_in,
                    int *res, uint32_t *cells, int16_t *dist_out) {
    k_stage_fse<<<1, 1, 0, st>>>(desc, n, max_sym, dist_in, ndist_in, al_in, res, cells, dist_out);
}
void zsbk_stage_huf(cudaStream_t st, const uint8_t *desc, uint32_t n, int *res, uint8_t *lens, uint16_t *lut) {
    k_stage_huf<<<1, 1, 0, st>>>(desc, n, res, lens, lut);
}
void zsbk_stage_records(cudaStream_t st, const uint32_t *tri, uint32_t nseq, uint32_t nlit, uint64_t *rec, ZsbBlockWork *w) {
    k_stage_records<<<1, 1, 0, st>>>(tri, nseq, nlit, rec, w);
}
