// zsb_kernels.cu -- the sm_100a kernels of the Zstandard decode path.
//
// Pipeline of one batch (all frames of the batch move through each stage together):
//
//   k_parse   1 lane / block      section headers                       (literals.rs:135-206, sequences.rs:52-143)
//   k_plan1   nf/1024 CTAs        per-frame table/tree chaining; the last CTA: scratch placement (scans) + work lists
//   k_huf     1 lane / stream     Huffman weights -> LUT (smem) -> 4-stream literal decode   (huffman.rs, literals.rs:49-86)
//   k_seq     1 lane / block      FSE tables (interleaved smem) + the serial 3-state chain      (fse.rs, sequence.rs, sequences.rs:191-237)
//           + 1 lane / 4 sequences extra bits, positions, repeat-offset history -> packed records (sequence.rs:41-55, decoding_context.rs:50-75)
//   (k_seqx   opt-in: k_seq whose phase-2 warps also execute the blocks whose place is known beforehand)
//   k_seq_slow 1 lane / block     careful decoder for blocks the fast path handed over (exact error order)
//   k_plan2   nf/1024 CTAs        repeat-offset history, frame sizes, size checks; the last CTA: output offsets (scan)
//   k_rawrle  1 CTA / block       raw / RLE block expansion, skippable payloads              (block.rs:76-79)
//   k_exec2   1 warp / frame      sequence execution through a 2 KiB shared-memory ring      (decoding_context.rs:78-106)
//   k_exec    1-3 CTAs / frame    the same for frames of many blocks: a 128 KiB block image in shared memory; with checksums one CTA of the
//                                 frame only hashes (XXH64) behind the bytes the others have committed
//   k_xxh     4 lanes / frame     XXH64 content checksum of the frames k_exec2 executed      (frame.rs:239-259)
//   k_publish (pipelined host path only) counters and per-frame results into page-locked host memory
//
// Nothing here is a dense contraction: no tensor cores.  The entropy stages are serial per stream, so they run
// lane-per-stream with all tables in shared memory and everything that is not on the serial chain moved to other warps;
// the execution stage runs a warp per frame and writes HBM with 16-byte stores.  DESIGN.md section 4 has the measurements
// behind each of these choices.
#include <cuda_runtime.h>
#include <stdlib.h>
#include "zsb_kernels.h"
#include "zsb_parse.h"
#include "zsb_huf.h"
#include "zsb_seqfast.h"
#include "zsb_bulk.cuh"

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

// ---- frames of many blocks in the two plan kernels.  chain_frame / plan_frame walk a frame block after block, one lane per frame: a DRAM round
// trip per block (1.2 us), 10 ms each for the 8 192 blocks of a 1 GiB frame.  A frame of more than PLAN_WARP_BLOCKS blocks is walked by the whole
// warp instead: 32 blocks are fetched at once, the state the reference carries from block to block (literals.rs:59-66, sequences.rs:147-187,
// decoding_context.rs:40,50-75) goes through them in registers, every lane writes its own block.
#define PLAN_WARP_BLOCKS 64u
__device__ __forceinline__ int warp_scan_max(int v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, v, d); if ((int)lane >= d) v = max(v, t); }
    return v;
}
__device__ __noinline__ int chain_frame_warp(const zsb_frame &fr, const zsb_block *blocks, ZsbBlockWork *work, uint32_t flags, uint32_t &err_a, uint32_t &err_b, uint32_t lane) {
    const bool quirks = (flags & ZSB_REFERENCE_QUIRKS) != 0;
    int huf_src = -1, ts[3] = {-1, -1, -1};        // last block with a Huffman description / with a table of its own per LL, OF, ML
    for (uint32_t k0 = 0; k0 < fr.n_blocks; k0 += 32) {
        const uint32_t k = k0 + lane, bi = fr.first_block + k;
        bool comp = false; int st = ZSB_OK; uint32_t lt = ZSB_LT_NONE, nseq = 0, m[3] = {0, 0, 0};
        if (k < fr.n_blocks && blocks[bi].type == ZSB_BT_COMPRESSED) {
            const ZsbBlockWork &w = work[bi];
            comp = true; st = w.status; lt = w.lit_type; nseq = w.nseq; m[0] = w.mode[0]; m[1] = w.mode[1]; m[2] = w.mode[2];
        }
        const int hinc = warp_scan_max((comp && lt == ZSB_LT_COMPRESSED) ? (int)bi : -1, lane);
        const int hsrc = max(hinc, huf_src);
        int tinc[3], tsrc[3];
        bool odd = comp && (st != ZSB_OK || (lt == ZSB_LT_TREELESS && hsrc < 0) || (nseq == 0 && quirks));
#pragma unroll
        for (int t = 0; t < 3; t++) {
            tinc[t] = warp_scan_max((comp && nseq && m[t] != ZSB_M_REPEAT) ? (int)bi : -1, lane);
            int ex = __shfl_up_sync(FULL, tinc[t], 1); if (lane == 0) ex = -1;
            tsrc[t] = max(ex, ts[t]);
            odd = odd || (comp && nseq && m[t] == ZSB_M_REPEAT && tsrc[t] < 0);
        }
        if (__any_sync(FULL, odd)) {
            // something the reference would report (or a block that failed to parse) among these 32: the lane-serial walk takes over from here,
            // with the same state, and gives the reference's first error
            int rc = 0;
            if (lane == 0) rc = chain_frame_from(fr, blocks, work, flags, err_a, err_b, k0, huf_src, ts[0], ts[1], ts[2]);
            err_a = __shfl_sync(FULL, err_a, 0); err_b = __shfl_sync(FULL, err_b, 0);
            return __shfl_sync(FULL, rc, 0);
        }
        if (comp) {
            ZsbBlockWork &w = work[bi];
            if (lt == ZSB_LT_TREELESS) { w.huf_desc = work[hsrc].huf_desc; w.huf_desc_end = work[hsrc].huf_desc_end; }
            if (nseq) {
#pragma unroll
                for (int t = 0; t < 3; t++)
                    if (m[t] == ZSB_M_REPEAT) { const ZsbBlockWork &sw = work[tsrc[t]]; w.mode[t] = sw.mode[t]; w.rle_sym[t] = sw.rle_sym[t]; w.tbl_desc[t] = sw.tbl_desc[t]; }
            }
        }
        huf_src = max(huf_src, __shfl_sync(FULL, hinc, 31));
#pragma unroll
        for (int t = 0; t < 3; t++) ts[t] = max(ts[t], __shfl_sync(FULL, tinc[t], 31));
    }
    return ZSB_OK;
}
__device__ __noinline__ int plan_frame_warp(const zsb_frame &fr, const zsb_block *blocks, ZsbBlockWork *work, uint64_t &total, uint32_t &err_a, uint32_t &err_b, uint32_t lane) {
    uint32_t rep[3] = {1, 4, 8};                                       // decoding_context.rs:40
    uint64_t pos = 0;
    for (uint32_t k0 = 0; k0 < fr.n_blocks; k0 += 32) {
        const uint32_t k = k0 + lane, bi = fr.first_block + k;
        const bool in = k < fr.n_blocks;
        int lst = ZSB_OK, st = ZSB_OK; uint32_t size = 0, ro[3] = {0, 0, 0}, ea = 0, eb = 0; bool has = false;
        if (in) {
            const ZsbBlockWork &w = work[bi];
            lst = w.lit_status; st = w.status; size = w.out_size; ea = w.err_a; eb = w.err_b;
            has = blocks[bi].type == ZSB_BT_COMPRESSED && w.nseq;
            ro[0] = w.rep_out[0]; ro[1] = w.rep_out[1]; ro[2] = w.rep_out[2];
        }
        const uint32_t fm = __ballot_sync(FULL, in && (lst != ZSB_OK || st != ZSB_OK));
        const uint32_t nrun = min(fm ? (uint32_t)__ffs(fm) - 1u : 32u, fr.n_blocks - k0);     // blocks of this group that are placed
        uint32_t inc = lane < nrun ? size : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, d); if ((int)lane >= d) inc += t; }
        uint32_t mine[3] = {rep[0], rep[1], rep[2]};
        for (uint32_t i = 0; i < nrun; i++) {                            // the repeat offsets go through the group in order, the same in every lane
            if (lane == i) { mine[0] = rep[0]; mine[1] = rep[1]; mine[2] = rep[2]; }
            const bool h = __shfl_sync(FULL, (int)has, i) != 0;
            const uint32_t o0 = __shfl_sync(FULL, ro[0], i), o1 = __shfl_sync(FULL, ro[1], i), o2 = __shfl_sync(FULL, ro[2], i);
            if (h) { const uint32_t n0 = seq_real_offset(o0, rep), n1 = seq_real_offset(o1, rep), n2 = seq_real_offset(o2, rep); rep[0] = n0; rep[1] = n1; rep[2] = n2; }
        }
        if (lane < nrun) {
            ZsbBlockWork &w = work[bi];
            w.out_off = pos + (inc - size);
            w.rep_in[0] = mine[0]; w.rep_in[1] = mine[1]; w.rep_in[2] = mine[2];
        }
        pos += __shfl_sync(FULL, inc, 31);
        if (fm) {                                                        // Block::decode of the first block that failed: literals first (block.rs:83-85)
            const int j = __ffs(fm) - 1;
            const int jl = __shfl_sync(FULL, lst, j), js = __shfl_sync(FULL, st, j);
            total = pos;
            if (jl != ZSB_OK) return jl;
            err_a = __shfl_sync(FULL, ea, j); err_b = __shfl_sync(FULL, eb, j);
            return js;
        }
    }
    total = pos;
    return ZSB_OK;
}

// ======================================================================================= k_plan1
// Several CTAs do the per-frame part side by side; the last one to finish (ticket) runs the scans, four blocks per thread.
__global__ void __launch_bounds__(1024) k_plan1(const zsb_frame *__restrict__ frames, uint32_t nf, const zsb_block *__restrict__ blocks,
                                                uint32_t nb, ZsbBlockWork *work, ZsbFrameOut *fout, uint32_t *huf_list, uint32_t *seq_list,
                                                ZsbCounters *cnt, uint64_t lit_cap, uint64_t seq_cap, uint32_t flags) {
    __shared__ uint64_t s_warp[33];
    __shared__ uint64_t s_base[4];
    __shared__ uint32_t s_last;
    const uint32_t tid = threadIdx.x;
    // (a) per-frame chaining of Huffman tables and table modes
    for (uint32_t fb = blockIdx.x * blockDim.x + (tid & ~31u); fb < nf; fb += gridDim.x * blockDim.x) {
        const uint32_t f = fb + (tid & 31u);
        ZsbFrameOut o; o.dst_off = 0; o.dst_len = 0; o.xxh64 = 0; o.err_a = 0; o.err_b = 0; o.pad = 0;
        o.status = f < nf ? frames[f].status : ZSB_OK;
        const bool z = f < nf && o.status == ZSB_OK && frames[f].kind == 0;
        const bool big = z && frames[f].n_blocks > PLAN_WARP_BLOCKS;
        if (z && !big) o.status = chain_frame(frames[f], blocks, work, flags, o.err_a, o.err_b);
        for (uint32_t m = __ballot_sync(FULL, big); m; m &= m - 1) {              // frames of many blocks: by the whole warp
            const uint32_t j = __ffs(m) - 1;
            uint32_t ea = 0, eb = 0;
            const int rc = chain_frame_warp(frames[fb + j], blocks, work, flags, ea, eb, tid & 31u);
            if ((tid & 31u) == j) { o.status = rc; o.err_a = ea; o.err_b = eb; }
        }
        if (f < nf) fout[f] = o;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&cnt->ticket1, 1u) == gridDim.x - 1 ? 1u : 0u;
    if (tid < 4) s_base[tid] = 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // (b) scratch placement and work lists: four scans over the blocks, four consecutive blocks per thread and pass
    for (uint32_t i0 = 0; i0 < nb; i0 += 4 * blockDim.x) {
        uint64_t lit_need[4], seq_need[4]; uint32_t hf[4], sf[4];
        uint64_t sl = 0, ss = 0, sh = 0, sq = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t i = i0 + 4 * tid + j;
            lit_need[j] = 0; seq_need[j] = 0; hf[j] = 0; sf[j] = 0;
            const int bst = i < nb && blocks[i].type == ZSB_BT_COMPRESSED ? __ldcg(&work[i].status) : ZSB_E_ARG;
            if (bst == ZSB_OK || ZSB_CHAIN_SEQ_ERROR(bst)) {
                if (__ldcg(&work[i].lit_type) >= ZSB_LT_COMPRESSED) { lit_need[j] = ((uint64_t)__ldcg(&work[i].lit_regen) + 15) & ~15ull; hf[j] = 1; }
                const uint32_t ns = __ldcg(&work[i].nseq);
                if (ns && bst == ZSB_OK) { seq_need[j] = ((uint64_t)ns + 1) & ~1ull; sf[j] = 1; }     // even: the records of a block start 16-byte aligned
            }
            sl += lit_need[j]; ss += seq_need[j]; sh += hf[j]; sq += sf[j];
        }
        uint64_t t0, t1, t2, t3;
        uint64_t a = cta_scan_excl(sl, s_warp, t0) + s_base[0], b2 = cta_scan_excl(ss, s_warp, t1) + s_base[1];
        uint64_t c = cta_scan_excl(sh, s_warp, t2) + s_base[2], d = cta_scan_excl(sq, s_warp, t3) + s_base[3];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t i = i0 + 4 * tid + j;
            if (i < nb) {
                if (hf[j]) { work[i].lit_buf = a; huf_list[c] = i; }
                if (sf[j]) { work[i].seq_buf = b2; seq_list[d] = i; }
            }
            a += lit_need[j]; b2 += seq_need[j]; c += hf[j]; d += sf[j];
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
// Half a warp per CTA, 4 blocks per CTA: lane = 4 * slot + stream.  Per slot: the two-symbol table (4 KiB: 1 024 cells indexed by the next ten
// bits of the stream, zsb_huf.h; aliased with the weight FSE table while the weights are being decoded), weights, counts, ranks.  Per lane: a
// 256-byte ring through which cp.async feeds its stream (zsb_stream.h); before the streams start, the slot's four rings hold T1, the first
// symbol under every 10-bit prefix, from which the four lanes build the table.  The decode is one dependent chain per stream (cell -> bits
// consumed -> next cell), so like k_seq the kernel is bound by the latency of that chain: two symbols per cell where ten bits hold two codes
// is what shortens it.  A stream the fast decode refuses and a block decoded the reference's way (quirks) need the one-symbol table of
// 2 048 16-bit cells: it is laid out over the two-symbol table once the slot's fast streams are through.
#define HUF_SLOTS 8
#define HUF_THREADS (4 * HUF_SLOTS)
#define HUF_MASK (HUF_THREADS == 32 ? 0xFFFFFFFFu : ((1u << HUF_THREADS) - 1u))
#define HUF_LUT_BYTES (2u << ZSB_HUF_MAX_BITS)   // 4096
struct __align__(16) HufSlot {
    union { uint16_t lut[1 << ZSB_HUF_MAX_BITS]; uint32_t pair[1 << ZSB_HUF_PAIR_BITS]; uint32_t ftbl[512]; } u;
    uint8_t weights[260];
    uint8_t odd[128];   // maxbits 11: the odd child under a 10-bit prefix of two 11-bit codes (huf_t1_put)
    uint16_t at[256];   // first LUT cell of every symbol (HUF_NO_ROOM: none); the four lanes of the slot fill the LUT from it
    int16_t cnt[ZSB_HUF_WEIGHT_SYMS];
    uint32_t rank[16];
    int maxbits;
    int status;
    int n, loose;       // symbols with the implied one; the LUT starts out as ZSB_HUF_ABSENT (huf_lut_plan)
    int incomplete;     // ZSB_REFERENCE_QUIRKS: the reference's tree for these weights is not a complete code (huf_build_lut)
    uint32_t stagger[4]; // the size is 16 mod 128: the same field of the eight slots of a CTA lies in eight different groups of four banks (the slots'
                        // serial set-up loops run side by side in one warp and would otherwise collide on every access: measured 8-way)
};
static_assert(sizeof(HufSlot) % 128 == 16, "HufSlot: slots must be staggered over the shared-memory banks");
#define HUF_NO_ROOM 0xFFFFu
#define HUF_SMEM_BYTES (HUF_THREADS * 256 + HUF_SLOTS * sizeof(HufSlot))
// `len` cells from lut[at] on: 16-byte stores where the alignment allows
__device__ __forceinline__ void huf_fill(uint16_t *lut, uint32_t at, uint32_t len, uint32_t cell) {
    uint32_t a = at; const uint32_t e = at + len;
    while (a < e && (a & 7u)) lut[a++] = (uint16_t)cell;
    const uint32_t c2 = cell | (cell << 16);
    for (; a + 8 <= e; a += 8) *reinterpret_cast<uint4 *>(lut + a) = make_uint4(c2, c2, c2, c2);
    while (a < e) lut[a++] = (uint16_t)cell;
}
// (rare: out of line, so that the stream decode keeps its registers)
__device__ __noinline__ int huf_block_ref_device(const uint8_t *src, uint64_t src_len, ZsbBlockWork &w, const uint16_t *lut, int maxbits, ZsbCounters *cnt, uint8_t *lit_pool,
                                                 uint64_t lit_cap, uint64_t over_cap) {
    // == LiteralsSection::decode (literals.rs:70-81): count, find room (the block's slot, else the overflow region), decode
    uint32_t n1 = 0, n2 = 0;
    int rc = huf_decode_block_ref(src, src_len, w.lit_src, w.stream_size, lut, maxbits, nullptr, 0, n1);
    if (!rc && n1 > ZSB_BLOCK_MAX) rc = ZSB_E_BLOCK_TOO_LARGE;
    if (rc) return rc;
    uint64_t at = w.lit_buf;
    if (n1 > ((w.lit_regen + 15u) & ~15u)) {
        const uint64_t need = ((uint64_t)n1 + 15) & ~15ull;
        const uint64_t o = atomicAdd((unsigned long long *)&cnt->lit_over, (unsigned long long)need);
        if (o + need > over_cap) return ZSB_E_CORRUPT;      // more corrupted blocks than the overflow region holds
        at = lit_cap + o;
    }
    rc = huf_decode_block_ref(src, src_len, w.lit_src, w.stream_size, lut, maxbits, lit_pool + at, n1, n2);
    if (!rc) { w.lit_buf = at; w.lit_regen = n1; }
    return rc;
}

#ifdef ZSB_SEQ_TIMING
__device__ long long g_huf_timing[1024][8];     // per CTA: 0 start, 3 weights read, 4 planned, 5 cell starts known, 6 T1 filled, 1 tables built, 2 streams decoded
extern "C" int zsb_debug_huf_timing(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_huf_timing, sizeof g_huf_timing); }
#define HUF_T(i) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_huf_timing[blockIdx.x][i] = clock64(); } while (0)
#else
#define HUF_T(i) do {} while (0)
#endif
__global__ void __launch_bounds__(HUF_THREADS) k_huf(const uint8_t *__restrict__ src, uint64_t src_len, ZsbBlockWork *work,
                                            const uint32_t *__restrict__ huf_list, ZsbCounters *cnt,
                                            uint8_t *lit_pool, uint64_t lit_cap, uint64_t over_cap, uint32_t flags) {
    extern __shared__ __align__(128) uint8_t huf_smem[];                        // HUF_SMEM_BYTES: the rings, then the slots
    uint8_t (*rings)[256] = reinterpret_cast<uint8_t (*)[256]>(huf_smem);
    HufSlot *slots = reinterpret_cast<HufSlot *>(huf_smem + HUF_THREADS * 256);
    if (cnt->overflow) return;
    const bool quirks = (flags & ZSB_REFERENCE_QUIRKS) != 0;
    const uint32_t lane = threadIdx.x, slot = lane >> 2, stream = lane & 3;
    const uint32_t n = cnt->n_huf, idx = blockIdx.x * HUF_SLOTS + slot;
    if (blockIdx.x * HUF_SLOTS >= n) return;
    const bool active = idx < n;
    const uint32_t bi = active ? huf_list[idx] : 0;
    HufSlot &S = slots[slot];
    HUF_T(0);
    if (active && stream == 0) {
        const ZsbBlockWork &w = work[bi];
        int nw = 0; uint32_t dl = 0;
        int rc = huf_read_weights(src + w.huf_desc, w.huf_desc_end - w.huf_desc, S.weights, 1, nw, dl, S.u.ftbl, 1, S.cnt, 1,
                                  src_len - w.huf_desc, quirks);
        HUF_T(3);
        // == huf_build_lut (zsb_huf.h), the filling left to all four lanes of the slot: one lane on its own spends 159 k cycles there
        // (2 048 cells a cell at a time, the four slots' loops diverging), against 44 k for the weights and 520 k for the streams
        HufPlan P; P.mb = 0; P.n = 0; P.loose = false;
        bool inc = false;
        if (!rc) rc = huf_lut_plan(S.weights, 1, nw, S.rank, 1, P, quirks, &inc);
        HUF_T(4);
        if (!rc) {
            for (int i = 0; i < P.n; i++) {                            // where every symbol's cells start: the classes fill up in symbol order
                const uint32_t wt = S.weights[i];
                uint32_t at = HUF_NO_ROOM;
                if (wt) {
                    const uint32_t len = 1u << (wt - 1);
                    at = S.rank[wt]; S.rank[wt] = at + len;
                    if (at + len > (1u << P.mb)) at = HUF_NO_ROOM;     // (quirks) no room left: the reference drops the symbol
                }
                S.at[i] = (uint16_t)at;
            }
        }
        S.maxbits = P.mb; S.n = P.n; S.loose = P.loose ? 1 : 0; S.status = rc; S.incomplete = inc ? 1 : 0;
    }
    __syncwarp(HUF_MASK);
    HUF_T(5);
    // the two-symbol table of a complete code: T1 in the slot's four rings (1 KiB, contiguous), then the cells
    uint8_t *t1 = rings[4 * slot];
    const bool pairs = active && !S.status && !S.incomplete;
    if (pairs) {
        for (int i = (int)stream; i < S.n; i += 4) {
            const uint32_t wt = S.weights[i];
            if (wt) huf_t1_put(t1, S.odd, S.maxbits, (uint32_t)i, S.at[i], wt);
        }
    }
    __syncwarp(HUF_MASK);
    HUF_T(6);
    if (pairs) {
        for (uint32_t x0 = stream + 4 * slot; x0 < (1u << ZSB_HUF_PAIR_BITS) + 4 * slot; x0 += 32) {      // four dependent loads per cell: eight cells side by side, stored
            uint32_t cell[8];                                                                                 // together; the slots start at different words of T1 (bank-aligned)
#pragma unroll
            for (int k = 0; k < 8; k++) cell[k] = huf_pair_cell((x0 + 4 * k) & ((1u << ZSB_HUF_PAIR_BITS) - 1u), t1, S.odd, S.weights, 1, S.maxbits);
#pragma unroll
            for (int k = 0; k < 8; k++) S.u.pair[(x0 + 4 * k) & ((1u << ZSB_HUF_PAIR_BITS) - 1u)] = cell[k];
        }
    }
    __syncwarp(HUF_MASK);
    HUF_T(1);
    int rc = 0;
    bool inexact = false;                       // ZSB_REFERENCE_QUIRKS: this block's literals are decoded the reference's way
    bool slow = false;                          // this stream goes through huf_decode_stream
    uint32_t expect = 0, ooff = 0; uint64_t start = 0;
    if (active) {
        rc = S.status;
        const ZsbBlockWork &w = work[bi];
        if (!rc && quirks && (w.lit_inexact || S.incomplete)) inexact = true;
        else if (!rc && stream < w.n_streams) {
            const uint32_t regen = w.lit_regen;
            start = w.lit_src;
            if (w.n_streams == 1) { expect = regen; ooff = 0; }
            else {
                const uint32_t seg = (regen + 3) / 4; ooff = stream * seg; expect = stream < 3 ? seg : regen - 3 * seg;
                for (uint32_t k = 0; k < stream; k++) start += w.stream_size[k];
            }
            rc = pairs ? huf_fast_stream(src, start, start + w.stream_size[stream], S.u.pair, lit_pool + w.lit_buf + ooff, expect,
                                         (uint32_t)__cvta_generic_to_shared(rings[lane]))
                       : ZSB_NEEDS_SLOW;
            if (rc == ZSB_NEEDS_SLOW) {
                rc = ZSB_OK;
                if (quirks) inexact = true;      // not the shape Regenerated_Size promises: the reference does not care (literals.rs:55)
                else slow = true;
            }
        }
    }
    __syncwarp(HUF_MASK);
    HUF_T(2);
    const uint32_t q0 = lane & ~3u;
    const bool blk_inexact = __shfl_sync(HUF_MASK, (int)inexact, q0) || __shfl_sync(HUF_MASK, (int)inexact, q0 + 1) || __shfl_sync(HUF_MASK, (int)inexact, q0 + 2) ||
                             __shfl_sync(HUF_MASK, (int)inexact, q0 + 3);
    const bool blk_slow = __shfl_sync(HUF_MASK, (int)slow, q0) || __shfl_sync(HUF_MASK, (int)slow, q0 + 1) || __shfl_sync(HUF_MASK, (int)slow, q0 + 2) ||
                          __shfl_sync(HUF_MASK, (int)slow, q0 + 3);
    // (rare) the one-symbol table == huf_build_lut (zsb_huf.h), laid out by the four lanes of the slot over the two-symbol table
    if ((blk_inexact || blk_slow) && active && !S.status) {
        const int mb = S.maxbits;
        if (S.loose) {
            for (uint32_t k = 8 * stream; k < (1u << mb); k += 32) {
                if (k + 8 <= (1u << mb)) *reinterpret_cast<uint4 *>(S.u.lut + k) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                else for (uint32_t j = k; j < (1u << mb); j++) S.u.lut[j] = ZSB_HUF_ABSENT;
            }
        }
    }
    __syncwarp(HUF_MASK);
    if ((blk_inexact || blk_slow) && active && !S.status) {
        const int mb = S.maxbits;
        for (int i = (int)stream; i < S.n; i += 4) {
            const uint32_t at = S.at[i];
            if (at == HUF_NO_ROOM) continue;
            const uint32_t wt = S.weights[i];
            huf_fill(S.u.lut, at, 1u << (wt - 1), (uint32_t)i | (((uint32_t)mb + 1 - wt) << 8));
        }
    }
    __syncwarp(HUF_MASK);
    if (slow) rc = huf_decode_stream(src, start, start + work[bi].stream_size[stream], src_len, S.u.lut, S.maxbits, lit_pool + work[bi].lit_buf + ooff, expect);
    if (blk_inexact && active && stream == 0 && !S.status) {
        rc = huf_block_ref_device(src, src_len, work[bi], S.u.lut, S.maxbits, cnt, lit_pool, lit_cap, over_cap);
    }
    // first failing stream of the block decides its status
    const int r1 = __shfl_sync(HUF_MASK, rc, q0 + 1), r2 = __shfl_sync(HUF_MASK, rc, q0 + 2), r3 = __shfl_sync(HUF_MASK, rc, q0 + 3);
    if (active && stream == 0) {
        int st = rc ? rc : r1 ? r1 : r2 ? r2 : r3;
        if (st) work[bi].lit_status = st;
    }
}

// ======================================================================================= FSE table build, one warp per table
// == FseTable::from_distribution (fse.rs:110-202), lane parallel.  The serial form (fse_build_table, zsb_fse.h) walks the
// table three times; here
//   (1) the "less than one" symbols take the cells N-1, N-2, ... in symbol order: a ballot rank (fse.rs:120-133);
//   (2) the spread visits position (k*step) & (N-1) at step k -- step is odd, so the N steps are a permutation -- and
//       skips positions above the low-probability zone: the j-th visited position gets occurrence j, whose symbol is
//       found by binary search in the prefix sums of the counts (fse.rs:136-157);
//   (3) the next-state numbers go to a symbol's cells in index order: per 32 cells, __match_any_sync groups equal symbols,
//       the rank inside the group plus a per-symbol running counter gives next = count + rank, then
//       nb = AL - floor(log2(next)), base = (next << nb) - N  (fse.rs:169-189 in closed form, see zsb_fse.h).
// Results are identical to fse_build_table cell for cell (tests: the reference's vectors and random distributions
// through zsb_fse_table_from_distribution, which runs this code).
struct FseWarpScratch { int16_t cnt[64]; uint16_t cum[66]; uint16_t next[64]; uint8_t desc[128]; };
// tab: code -> baseline | extra bits << 24 for LL ([0..35]) and ML ([36..88]); type 3: plain table (xb = 0, code = symbol)
// stage (optional, N bytes of shared memory): the symbol of every cell is kept there until the cells are computed, and only the finished
// cells go to tbl.  k_seq builds into columns of an interleaved table (ts = 32 words: every access of the warp hits ONE bank, 32 wavefronts
// each) -- staged, a table costs 16 such stores instead of 16 stores, 16 loads and 16 more stores (tried and measured: one LANE per table,
// 32 tables per warp, conflict free but serial and divergent: 1.2 M cycles of set-up per CTA instead of 0.19 M).
__device__ __forceinline__ int fse_build_table_warp(FseWarpScratch &X, int nsym, int al, uint32_t *tbl, int ts, int type, const uint32_t *tab, uint32_t lane,
                                                    uint8_t *stage = nullptr) {
    const int N = 1 << al, step = (N >> 1) + (N >> 3) + 3, mask = N - 1;
    const uint32_t lt = (1u << lane) - 1u;
    int nlow = 0, total = 0;
    bool overflow = false;
    for (int s0 = 0; s0 < nsym; s0 += 32) {
        const int sy = s0 + (int)lane;
        const int c = sy < nsym ? X.cnt[sy] : 0;
        const bool low = c == -1;
        const uint32_t lm = __ballot_sync(FULL, low);
        const int p = c > 0 ? c : 0;
        int inc = p;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, inc, d); if ((int)lane >= d) inc += t; }
        if (sy < nsym) { X.cum[sy] = (uint16_t)(total + inc - p); X.next[sy] = (uint16_t)(low ? 1 : p); }
        if (low) {
            const int cell = N - 1 - (nlow + __popc(lm & lt));
            if (cell < 0) overflow = true; else if (stage) stage[cell] = (uint8_t)sy; else tbl[cell * ts] = (uint32_t)sy;
        }
        nlow += __popc(lm); total += __shfl_sync(FULL, inc, 31);
    }
    const int high = N - 1 - nlow;
    if (__any_sync(FULL, overflow) || total != high + 1) return ZSB_E_CORRUPTED_TABLE;       // fse.rs:160-166 and the guards of the serial form
    if (lane == 0) X.cum[nsym] = (uint16_t)total;
    __syncwarp();
    int jbase = 0;
    for (int k0 = 0; k0 < N; k0 += 32) {
        const int pos = ((k0 + (int)lane) * step) & mask;
        const bool valid = pos <= high;
        const uint32_t b = __ballot_sync(FULL, valid);
        const int j = jbase + __popc(b & lt);
        jbase += __popc(b);
        if (valid) {
            int lo = 0, hi = nsym;                      // cum[lo] <= j < cum[hi]
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int)X.cum[mid] <= j) lo = mid; else hi = mid; }
            if (stage) stage[pos] = (uint8_t)lo; else tbl[pos * ts] = (uint32_t)lo;
        }
    }
    __syncwarp();
    for (int i0 = 0; i0 < N; i0 += 32) {
        const int i = i0 + (int)lane;
        const uint32_t sy = stage ? (uint32_t)stage[i] : tbl[i * ts];
        const uint32_t g = __match_any_sync(FULL, sy);
        const uint32_t r = __popc(g & lt);
        const uint32_t nx = (uint32_t)X.next[sy] + r;
        __syncwarp();
        if (r == 0) X.next[sy] = (uint16_t)(nx + __popc(g));
        __syncwarp();
        const uint32_t nb = (uint32_t)(al - zsb_flog2(nx));
        const uint32_t base = (nx << nb) - (uint32_t)N;
        uint32_t code = sy, xb = 0;
        if (type != 3) {
            const uint32_t mx = type == 0 ? ZSB_MAX_LL_CODE : type == 1 ? ZSB_MAX_OF_CODE : ZSB_MAX_ML_CODE;
            if (sy > mx) code = 63u;
            else xb = type == 1 ? sy : tab[(type == 2 ? 36 : 0) + sy] >> 24;
        }
        tbl[i * ts] = ZSB_CELL(nb, xb, base, code);
    }
    return ZSB_OK;
}
// Table t (0 LL, 1 OF, 2 ML) of block w by one warp: == seq_build_table (zsb_seq.h) with max_sym = 64.  All lanes return the same status.
__device__ __forceinline__ int seq_build_table_warp(const uint8_t *src, const ZsbBlockWork &w, int t, uint32_t *tbl, int ts, FseWarpScratch &X,
                                                    const uint32_t *tab, int max_al, int &al_out, uint32_t lane, uint8_t *stage = nullptr) {
    const int mode = w.mode[t];
    al_out = 0;
    if (mode == ZSB_M_RLE) { if (lane == 0) fse_build_rle(w.rle_sym[t], tbl, t); return ZSB_OK; }
    int al = 0, nsym = 0, rc = ZSB_OK;
    if (mode == ZSB_M_PREDEFINED) {
        nsym = zsb_predef_nsym(t); al = zsb_predef_al(t);
        for (int sy = (int)lane; sy < nsym; sy += 32) X.cnt[sy] = (int16_t)zsb_predef_count(t, sy);
    } else if (mode == ZSB_M_FSE) {
        // the description is read by one lane, a few bits per symbol: from a shared-memory copy that the warp fetches in one go -- read
        // where it lies, every symbol costs a dependent round trip to L2/HBM (measured: 31 k cycles per table, 95 us of set-up per CTA)
        const uint8_t *dp = src + w.tbl_desc[t];
        const uint64_t avail = w.tbl_end - w.tbl_desc[t];
        const uint32_t nb = avail < sizeof X.desc ? (uint32_t)avail : (uint32_t)sizeof X.desc;
        for (uint32_t i = lane; i < nb; i += 32) X.desc[i] = __ldg(dp + i);
        __syncwarp();
        if (lane == 0) {
            FwdBits f; fwd_init(f, X.desc, nb);
            rc = fse_read_ncount(f, X.cnt, 1, 64, al, nsym);
            if (rc == ZSB_E_NOT_ENOUGH_BITS && avail > nb) { fwd_init(f, dp, avail); rc = fse_read_ncount(f, X.cnt, 1, 64, al, nsym); }     // longer than the copy
            if (rc == ZSB_E_CORRUPT) rc = ZSB_TABLE_TOO_SMALL;            // more than 64 symbols described
        }
        rc = __shfl_sync(FULL, rc, 0); al = __shfl_sync(FULL, al, 0); nsym = __shfl_sync(FULL, nsym, 0);
        if (rc) return rc;
    } else return ZSB_E_NO_PREVIOUS_DECODER;
    if (max_al && al > max_al) return ZSB_TABLE_TOO_SMALL;
    __syncwarp();
    rc = fse_build_table_warp(X, nsym, al, tbl, ts, t, tab, lane, stage);
    __syncwarp();
    if (!rc) al_out = al;
    return rc;
}

// ======================================================================================= k_seq / k_seq_slow
// The sequence stage in two phases (zsb_seqfast.h), fused in one CTA:
//
//   warp 0 (producer)  the serial three-state FSE chain, one lane per block, SEQ_CHAINS = 32 blocks per CTA (so that a SM
//                      hosts one producer warp and the phase-2 warps mostly run on the other sub-cores).  The chain is
//                      latency bound (table cells -> their sum -> the state bits out of a register window -> next cell
//                      addresses: one shared-memory load and ~6 dependent ALU operations per sequence, SEQ_STEP), so it carries
//                      as little else as possible: tables interleaved across the lanes (cell i of lane l at word i*32 + l:
//                      bank = l, conflict free), the bit window in registers, refilled from a cp.async stream ring, one 32-bit
//                      word per sequence into a shared-memory ring.
//   warps 1..H (phase 2) k_seq: four sequences per lane, 128 per step and block: bit positions, extra-bit values,
//                      literal/output positions and the repeat-offset history by warp prefix operations; packed
//                      records to HBM.  They run in the issue slots the producers leave empty (~70 %).
//
// Hand-over: the chains advance in lockstep, one window per chain (k_seq: 128 sequences, two windows in the ring) at a time;
// named barriers (window written / window consumed, one pair per ring slot) pass the batches on, so a waiting warp costs no
// issue slot.
#define SEQ_TBL_CELLS 512
#define SEQ_CHAINS 32         // table columns / producer lanes of a CTA
// (how many of them carry a block is chosen per launch so that the CTAs fill whole waves of one CTA per SM: 4 096 blocks on
//  148 SMs run 28 chains per CTA in 147 CTAs; 32 would leave 20 SMs idle and load the others' phase-2 warps more)
#define SEQ_HELPERS 16        // k_seq: phase-2 warps, two chains each
#define SEQ_OF_CELLS 256      // offset tables have accuracy log <= 8 (RFC 8878); a log-9 one (the reference accepts it) takes the careful path
#define SEQ_TBL_BYTES ((2 * SEQ_TBL_CELLS + SEQ_OF_CELLS) * SEQ_CHAINS * 4)
#define SEQ_WIN 128           // k_seq: sequences per hand-over and chain (SEQ_PER_LANE = 4 per phase-2 lane), two windows in the ring
#define SEQ_PF 8              // lines (of 128 bytes) the stream rings ask L2 for ahead of their own requests
// k_seqx (sequence decoding + execution in one kernel): one consumer warp per chain, windows of 32 sequences, four in the ring
#define SEQX_WARPS 28
#define SEQX_WIN 32
#define SEQX_NBUF 4
#define SEQX_RING 1024u       // bytes of recent output a consumer warp keeps in shared memory (k_exec2: 2 KiB)
// WSTRIDE: words per chain in the ring (the windows) + 4: rows stay 16-byte aligned, banks spread
template <int WSTRIDE>
struct SeqSharedT {
    uint32_t words[SEQ_CHAINS][WSTRIDE];
    uint32_t tab[36 + 53];                       // code -> baseline | extra bits << 24
    uint32_t nseq[SEQ_CHAINS];                   // 0: chain not running (no block, or it failed before the first sequence)
    uint32_t regen[SEQ_CHAINS], bi[SEQ_CHAINS];
    unsigned long long top0[SEQ_CHAINS], rec[SEQ_CHAINS];   // absolute bit position of sequence 0, record pointer
    int final_rc[SEQ_CHAINS];
    int tal[SEQ_CHAINS][3], trc[SEQ_CHAINS][3];  // accuracy log / build status of each table (built by the phase-2 warps during set-up)
    unsigned long long gdst[SEQ_CHAINS];         // k_seqx: where the block's output goes (0: the chain only writes records)
    uint32_t limit[SEQ_CHAINS];                  // k_seqx: bytes the block may produce there
};
// named barriers of the hand-over (0 is __syncthreads): 1 + (window % NBUF): window written, 1 + NBUF + (window % NBUF): window
// consumed; every thread of the CTA takes part.  1 + 2 * NBUF: the phase-2 warps among themselves.
template <int NT> __device__ __forceinline__ void seq_bar_sync(uint32_t id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(NT) : "memory"); }
template <int NT> __device__ __forceinline__ void seq_bar_arrive(uint32_t id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(NT) : "memory"); }
// tables | counts, then stream rings | hand-over (| the consumer warps' output rings)
#define SEQ_SMEM_BYTES (SEQ_TBL_BYTES + 256 * SEQ_CHAINS * 2 + sizeof(SeqSharedT<SEQ_WIN * 2 + 4>))
#define SEQX_SMEM_BYTES (SEQ_TBL_BYTES + 256 * SEQ_CHAINS * 2 + ((sizeof(SeqSharedT<SEQX_WIN * SEQX_NBUF + 4>) + 15) & ~15ull) + SEQX_WARPS * SEQX_RING)

__device__ __forceinline__ Hist hist_shfl_up(const Hist &h, int d) {
    Hist r; r.h0 = __shfl_up_sync(FULL, h.h0, d); r.h1 = __shfl_up_sync(FULL, h.h1, d); r.h2 = __shfl_up_sync(FULL, h.h2, d); return r;
}
__device__ __forceinline__ Hist hist_bcast(const Hist &h, int l) {
    Hist r; r.h0 = __shfl_sync(FULL, h.h0, l); r.h1 = __shfl_sync(FULL, h.h1, l); r.h2 = __shfl_sync(FULL, h.h2, l); return r;
}
// phase 2 for 32 consecutive sequences of one block; C carries the block-level state from batch to batch (warp uniform, but `bad`)
struct Seq2Carry { int64_t top; uint32_t lit_acc, out_acc; Hist H; int bad; };
// codes -> (ll, ml, offset_value, bits consumed) of one sequence whose first bit lies just below absolute bit `top`
struct Seq2One { uint32_t ll, ml, ov; };
__device__ __forceinline__ void seq2_codes(uint32_t tab_sa, uint32_t word, bool valid, uint32_t &eL, uint32_t &eM, uint32_t &cO, uint32_t &px,
                                           uint32_t &tot, int &bad) {
    uint32_t cL = ZSB_W_CL(word), cM = ZSB_W_CM(word);
    cO = ZSB_W_CO(word);
    if (cL > ZSB_MAX_LL_CODE || cO > ZSB_MAX_OF_CODE || cM > ZSB_MAX_ML_CODE) { bad = 1; cL = cO = cM = 0; }    // sequence.rs:46-48
    // small codes carry no extra bits (sequence.rs:98-191: LL codes 0..15 are the length itself, ML codes 0..31 the length - 3):
    // the table is only consulted for the rare larger ones, which keeps these loads off the SM's shared-memory path
    eL = cL; eM = cM + 3u;
    if (cL >= 16) eL = zsb_lds32(tab_sa + 4u * cL);
    if (cM >= 32) eM = zsb_lds32(tab_sa + 4u * (36u + cM));
    px = valid ? (eL >> 24) + (eM >> 24) + cO : 0u;
    tot = valid ? px + ZSB_W_NB(word) : 0u;
}
__device__ __forceinline__ Seq2One seq2_values(const uint8_t *base8, int64_t top, bool valid, uint32_t eL, uint32_t eM, uint32_t cO, uint32_t px, int &bad) {
    Seq2One r; r.ll = 0; r.ml = 0; r.ov = 1;
    if (valid) {
        const uint32_t xL = eL >> 24, xM = eM >> 24;
        int64_t a = top - px;
        if (a < 0) { bad = 1; a = 0; }                            // only after an over-read (the producer reports it too)
        const uint64_t Wx = px ? fast_win_at(base8, a, px) : 0ull;
        r.ll = (eL & 0xFFFFFFu) + ((uint32_t)Wx & ((1u << xL) - 1u));
        r.ml = (eM & 0xFFFFFFu) + ((uint32_t)(Wx >> xL) & ((1u << xM) - 1u));
        r.ov = (1u << cO) + ((uint32_t)(Wx >> (xL + xM)) & ((1u << cO) - 1u));
    }
    return r;
}
// Phase 2 for SEQ_WIN = 32 * SEQ_PER_LANE consecutive sequences of one block, SEQ_PER_LANE per lane (lane l: sequences
// i0 + K*l .. i0 + K*l + K-1).  A lane folds its sequences locally and the warp prefix operations run once per window: a
// fraction of the shuffles and compositions of a one-per-lane layout, which matters because the shuffles share the SM's
// load/store path with the producer's table loads.
#define SEQ_PER_LANE 4
template <int K>
__device__ __forceinline__ void seq2_records(const uint8_t *base8, uint32_t tab_sa, const uint32_t *wd, uint32_t i0, uint32_t nseq, uint32_t regen,
                                             Seq2Carry &C, uint32_t lane, bool (&v)[K], uint64_t (&r)[K]) {
    const uint32_t ia = i0 + K * lane;
    uint32_t eL[K], eM[K], cO[K], px[K], tot[K];
    uint32_t lane_tot = 0;
#pragma unroll
    for (int j = 0; j < K; j++) { v[j] = ia + j < nseq; seq2_codes(tab_sa, wd[j], v[j], eL[j], eM[j], cO[j], px[j], tot[j], C.bad); lane_tot += tot[j]; }
    // bit positions: exclusive prefix of the bits consumed
    uint32_t inc = lane_tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, d); if (lane >= (uint32_t)d) inc += t; }
    int64_t tp = C.top - (int64_t)(inc - lane_tot);
    Seq2One Q[K];
    uint32_t sll = 0, sout = 0;
#pragma unroll
    for (int j = 0; j < K; j++) { Q[j] = seq2_values(base8, tp, v[j], eL[j], eM[j], cO[j], px[j], C.bad); tp -= tot[j]; sll += Q[j].ll; sout += Q[j].ll + Q[j].ml; }
    C.top -= (int64_t)__shfl_sync(FULL, inc, 31);
    // literal / output positions
    uint64_t pos = (uint64_t)sll | ((uint64_t)sout << 32);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(FULL, pos, d); if (lane >= (uint32_t)d) pos += t; }
    uint32_t lit = C.lit_acc + (uint32_t)pos - sll, out = C.out_acc + (uint32_t)(pos >> 32) - sout;     // before this lane's first sequence
    C.lit_acc = __shfl_sync(FULL, C.lit_acc + (uint32_t)pos, 31); C.out_acc = __shfl_sync(FULL, C.out_acc + (uint32_t)(pos >> 32), 31);
    // repeat-offset history: the lane's transforms folded, an inclusive prefix over the lanes, then the block-level carry
    Hist P[K];                                                     // P[j]: this lane's sequences 0..j applied in order
#pragma unroll
    for (int j = 0; j < K; j++) {
        const Hist F = v[j] ? hist_of_sequence(Q[j].ov, Q[j].ll, C.bad) : hist_identity();
        P[j] = j ? hist_compose(F, P[j - 1], C.bad) : F;
    }
    Hist G = P[K - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        // a transform whose three slots are constants is not changed by what came before it: once every lane that still has a
        // partner (lane >= d) is closed, the remaining rounds are no-ops (after 8-16 sequences nearly every transform is closed)
        if (!__any_sync(FULL, lane >= (uint32_t)d && ((G.h0 | G.h1 | G.h2) & ZSB_OFF_SYM))) break;
        const Hist E = hist_shfl_up(G, d);
        if (lane >= (uint32_t)d) G = hist_compose(G, E, C.bad);
    }
    Hist E = hist_shfl_up(G, 1);                                   // history before this lane's first sequence, relative to the window start
    if (lane == 0) E = hist_identity();
#pragma unroll
    for (int j = 0; j < K; j++) {
        lit += Q[j].ll; out += Q[j].ll + Q[j].ml;
        if (v[j] && (lit > regen || out + (regen - lit) > ZSB_BLOCK_MAX)) C.bad = 1;               // decoding_context.rs:86-90, Block_Maximum_Size
        const uint32_t off = hist_pick(C.H, hist_pick(E, P[j].h0, C.bad), C.bad);                  // the offset a sequence uses is slot 0 after it
        r[j] = (uint64_t)out | ((uint64_t)lit << ZSB_REC_POS_BITS) | ((uint64_t)off << (2 * ZSB_REC_POS_BITS));
    }
    C.H = hist_compose(hist_bcast(G, 31), C.H, C.bad);
}
__device__ __forceinline__ void seq2_window(const uint8_t *base8, uint32_t tab_sa, const uint32_t *wd, uint32_t i0, uint32_t nseq, uint32_t regen,
                                            uint64_t *rec, Seq2Carry &C, uint32_t lane) {
    constexpr int K = SEQ_PER_LANE;
    const uint32_t ia = i0 + K * lane;
    bool v[K]; uint64_t r[K];
    seq2_records<K>(base8, tab_sa, wd, i0, nseq, regen, C, lane, v, r);
    // (records are 16-byte aligned: seq_buf is kept even and K is even)
#pragma unroll
    for (int j = 0; j < K; j += 2) {
        if (v[j + 1]) *reinterpret_cast<ulonglong2 *>(rec + ia + j) = make_ulonglong2(r[j], r[j + 1]);
        else if (v[j]) rec[ia + j] = r[j];
    }
}

// -DZSB_SEQ_TIMING (tools/probes/seq_timing.py): per-CTA cycle counts of the stages of k_seq
#ifdef ZSB_SEQ_TIMING
__device__ long long g_seq_timing[160][8];     // per CTA: start, tables built, states read, producer loop done, cycles the producer waited for free windows, helper 1: waited / worked / windows
extern "C" int zsb_debug_seq_timing(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_seq_timing, sizeof g_seq_timing); }
#define SEQ_T(i) do { if (lane == 0 && blockIdx.x < 160) g_seq_timing[blockIdx.x][i] = clock64(); } while (0)
#else
#define SEQ_T(i) do {} while (0)
#endif
// k_seq  = k_seq_t<16, 128, 2, false>: records to HBM, executed by k_exec2 / k_exec afterwards.
// k_seqx = k_seq_t<28, 32, 4, true>:   one consumer warp per chain does phase 2 one sequence per lane and, where the place of the
//          block's output is known beforehand (first block of a frame whose predecessors all declare their sizes), executes
//          the sequences at once (seqx_consume, further down): no record round trip through HBM, and execution fills the issue slots the
//          chain leaves empty.  Chains whose block cannot be placed write records as k_seq does.
struct SeqxArgs { const zsb_frame *frames; const zsb_block *blocks; const uint64_t *pre_off; const uint8_t *lit_pool; uint8_t *dst; };
template <int HELPERS, int WIN, int NBUF>
__device__ __forceinline__ void seqx_consume(uint8_t *smem_rings, SeqSharedT<WIN * NBUF + 4> &S, const uint8_t *src, const uint8_t *base8, ZsbBlockWork *work,
                                             ZsbCounters *cnt, uint32_t *slow_list, const SeqxArgs &A, uint32_t warp, uint32_t lane);
template <int HELPERS, int WIN, int NBUF, bool FUSED>
__global__ void __launch_bounds__(32 * (1 + HELPERS), 1) k_seq_t(const uint8_t *__restrict__ src, ZsbBlockWork *work, const uint32_t *__restrict__ seq_list,
                                                                ZsbCounters *cnt, uint64_t *seq_pool, uint32_t *slow_list, uint32_t used, SeqxArgs A) {
    constexpr int NT = 32 * (1 + HELPERS), CPH = SEQ_CHAINS / HELPERS;       // threads of the CTA; chains per phase-2 warp
    constexpr uint32_t BAR_FULL = 1u, BAR_FREE = 1u + NBUF, BAR_HELPERS = 1u + 2u * NBUF;
    typedef SeqSharedT<WIN * NBUF + 4> SeqShared;
    extern __shared__ __align__(1024) uint8_t smem[];      // the stream rings must be 512-byte aligned (SEQ_STEP)
    if (cnt->overflow) return;
    uint32_t *tbl = reinterpret_cast<uint32_t *>(smem);
    int16_t *counts = reinterpret_cast<int16_t *>(smem + SEQ_TBL_BYTES);
    SeqShared &S = *reinterpret_cast<SeqShared *>(smem + SEQ_TBL_BYTES + 256 * SEQ_CHAINS * 2);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = cnt->n_seq;
    const uint32_t mis = (uint32_t)((uintptr_t)src & 7);
    const uint8_t *base8 = src - mis;

    if (warp == 0) SEQ_T(0);
    // ---- set-up: the phase-2 warps fill the code tables and build the 3 x 32 FSE tables, one warp per table (fse_build_table_warp);
    // then the producer lanes read their initial states
    SeqTables T;
    struct { uint64_t bs_off, seq_buf; uint32_t bs_len, nseq, lit_regen; int32_t lit_status; } w = {0, 0, 0, 0, 0, 0};   // the fields of work[bi] the producer lane needs
    bool active = false;
    uint32_t bi = 0;
    FastWin F; StreamRing R;
    int32_t top = 0, startbit = 0;
    uint32_t aL = 0, aO = 0, aM = 0, tbL = 0, tbO = 0, tbM = 0;
    R.sa = 0; R.pl = base8; R.low = 0;
    if (warp == 0) {
        const uint32_t idx = blockIdx.x * used + lane;
        active = lane < used && idx < n;
        bi = active ? seq_list[idx] : 0;
        if (active) {
            const ZsbBlockWork &g = work[bi];
            w.bs_off = g.bs_off; w.seq_buf = g.seq_buf; w.bs_len = g.bs_len; w.nseq = g.nseq; w.lit_regen = g.lit_regen; w.lit_status = g.lit_status;
            active = g.status == ZSB_OK;
        }
    } else {
        for (uint32_t k = threadIdx.x - 32; k < 36 + 53; k += 32 * HELPERS) S.tab[k] = k < 36 ? zsb_ll_entry(k) : zsb_ml_entry(k - 36);
        __syncwarp();
        // the code tables are filled by all phase-2 warps together: wait for all of them
        seq_bar_sync<32 * HELPERS>(BAR_HELPERS);
        FseWarpScratch &X = reinterpret_cast<FseWarpScratch *>(counts)[warp - 1];
        // (the word ring is idle during set-up: every warp keeps the symbols of the table it is building there, see fse_build_table_warp)
        uint8_t *stage = reinterpret_cast<uint8_t *>(&S.words[0][0]) + (warp - 1) * SEQ_TBL_CELLS;
        static_assert(sizeof S.words >= HELPERS * SEQ_TBL_CELLS, "symbol staging does not fit the word ring");
        for (uint32_t q = 0; q < 3 * CPH; q++) {
            const uint32_t c = (warp - 1) * CPH + q / 3, t = q % 3, idx = blockIdx.x * used + c;
            int rc = ZSB_OK, al = 0;
            if (c < used && idx < n) {
                const ZsbBlockWork &wb = work[seq_list[idx]];
                if (wb.status == ZSB_OK) {
                    uint32_t *tb = tbl + (t == 0 ? 0 : t == 1 ? SEQ_TBL_CELLS * SEQ_CHAINS : (SEQ_TBL_CELLS + SEQ_OF_CELLS) * SEQ_CHAINS) + c;
                    rc = seq_build_table_warp(src, wb, (int)t, tb, SEQ_CHAINS, X, S.tab, t == 1 ? 8 : 9, al, lane, stage);
                }
            }
            if (lane == 0) { S.tal[c][t] = al; S.trc[c][t] = rc; }
        }
    }
    __syncthreads();                                   // tables built; the count area becomes the stream rings
    if (warp == 0) SEQ_T(1);
    if (warp == 0) {
        T.ts = SEQ_CHAINS;
        T.tbl[0] = tbl + lane; T.tbl[1] = tbl + SEQ_TBL_CELLS * SEQ_CHAINS + lane; T.tbl[2] = tbl + (SEQ_TBL_CELLS + SEQ_OF_CELLS) * SEQ_CHAINS + lane;
        int rc = ZSB_OK;
        if (active) {
            rc = S.trc[lane][0] ? S.trc[lane][0] : S.trc[lane][1] ? S.trc[lane][1] : S.trc[lane][2];     // the reference's order: LL, OF, ML
            T.al[0] = S.tal[lane][0]; T.al[1] = S.tal[lane][1]; T.al[2] = S.tal[lane][2];
        }
        if (active && !rc) {
            // == seq_fast_phase1 up to the first sequence (zsb_seqfast.h), on the stream ring
            const uint64_t start = w.bs_off;
            const uint32_t lastb = w.bs_len ? src[start + w.bs_len - 1] : 0;
            if (w.bs_len == 0) rc = ZSB_E_EMPTY_INPUT_DATA;
            else if (lastb == 0) rc = ZSB_E_NULL_BYTE;
            else if (start < 16 || w.bs_len > (1u << 24)) rc = ZSB_NEEDS_SLOW;
            else {
                R.sa = (uint32_t)__cvta_generic_to_shared(smem + SEQ_TBL_BYTES) + lane * 512u;
                R.pl = (const uint8_t *)(((uintptr_t)(src + start) - 16) & ~(uintptr_t)127);
                const uint32_t d0 = (uint32_t)((src + start) - R.pl);                     // 16 .. 143
                top = (int32_t)((d0 + w.bs_len - 1) * 8) + zsb_flog2(lastb);
                startbit = (int32_t)(d0 * 8);
                const uint32_t a0 = (uint32_t)T.al[0], a1 = (uint32_t)T.al[1], a2 = (uint32_t)T.al[2];
                if (top - startbit < (int32_t)(a0 + a1 + a2)) rc = ZSB_E_NOT_ENOUGH_BITS;
                else {
                    sr_init<7, SEQ_PF>(R, top);
                    sr_load<7>(R, F, top);
                    const uint64_t W = fast_win_get(F);
                    const uint32_t sL = (uint32_t)zsb_shr64(W, 64 - a0), sO = (uint32_t)zsb_shr64(zsb_shl64(W, a0), 64 - a1),
                                   sM = (uint32_t)zsb_shr64(zsb_shl64(W, a0 + a1), 64 - a2);
                    top -= (int32_t)(a0 + a1 + a2);
                    // states are kept as shared-memory byte addresses of their cells: next = (table + base*stride) + bits*stride
                    tbL = (uint32_t)__cvta_generic_to_shared(T.tbl[0]); tbO = (uint32_t)__cvta_generic_to_shared(T.tbl[1]);
                    tbM = (uint32_t)__cvta_generic_to_shared(T.tbl[2]);
                    aL = tbL + sL * (SEQ_CHAINS * 4); aO = tbO + sO * (SEQ_CHAINS * 4); aM = tbM + sM * (SEQ_CHAINS * 4);
                }
            }
        }
        if (active && rc) {
            if (rc == ZSB_TABLE_TOO_SMALL) rc = ZSB_NEEDS_SLOW;
            if (rc == ZSB_NEEDS_SLOW) slow_list[atomicAdd(&cnt->n_slow, 1u)] = bi;
            work[bi].status = rc;
            active = false;
        }
        S.nseq[lane] = active ? w.nseq : 0u;
        S.regen[lane] = w.lit_regen; S.bi[lane] = bi;
        S.top0[lane] = (unsigned long long)((int64_t)(R.pl - base8) * 8 + top);
        S.rec[lane] = (unsigned long long)(uintptr_t)(seq_pool + w.seq_buf);
        S.final_rc[lane] = ZSB_OK;
        if (FUSED) {
            // the block can be executed at once if its place in the output is known now: the first block of a frame the host could place
            // (pre_off: every frame before it declares its size), literals there (k_huf ran before this kernel)
            unsigned long long g = 0; uint32_t lim = 0;
            if (active && A.pre_off && w.lit_status == ZSB_OK) {
                const uint32_t f = A.blocks[bi].frame;
                const uint64_t po = A.pre_off[f];
                if (po != ~0ull && A.frames[f].first_block == bi) {
                    g = (unsigned long long)(uintptr_t)(A.dst + po);
                    const uint64_t cs = A.frames[f].content_size;
                    lim = cs > ZSB_BLOCK_MAX ? (uint32_t)ZSB_BLOCK_MAX : (uint32_t)cs;
                }
            }
            S.gdst[lane] = g; S.limit[lane] = lim;
        }
    }
    __syncthreads();

    if (warp == 0) SEQ_T(2);
#ifdef ZSB_SEQ_TIMING
    long long t_wait = 0, t_work = 0;
#endif
    if (warp == 0) {
        // ---- producer: == the loop of seq_fast_phase1
        const uint32_t nseq = active ? w.nseq : 0u;
        uint32_t maxn = nseq;
#pragma unroll
        for (int d = 16; d; d >>= 1) maxn = max(maxn, __shfl_xor_sync(FULL, maxn, d));
        uint32_t *wrow = &S.words[lane < SEQ_CHAINS ? lane : 0][0];
        // The bit window in registers.  The stream ring is never read on the chain: q2, q1, q0 are the ring words k, k-1, k-2
        // (k = the word holding the next unread bit, `o` bits of it already consumed), n0 the word below them, n1..n3 three more
        // that every step requests behind its cell loads, T2:T1:T0 the 96 bits from the cursor on, top-aligned ((q2:q1:q0:n0) << o).  A sequence consumes px <= 63
        // extra bits and then <= 27 state bits, i.e. the state bits lie inside T whatever the codes are.
        const uint32_t ring_sa = R.sa;                   // 512-byte aligned: ring byte address = ring_sa | (byte offset & 0x1FC)
        const uint32_t zr = cnt->zero;                   // 0, but not to the compiler: see SEQ_STEP
        uint32_t kb = 0, o = 0, q2 = 0, q1 = 0, q0 = 0, n0 = 0, n1 = 0, n2 = 0, n3 = 0, T2 = 0, T1 = 0, T0 = 0;
        if (active) {
            kb = (uint32_t)((top - 1) >> 5) * 4u; o = (32u - ((uint32_t)top & 31u)) & 31u;
            q2 = zsb_lds32v((kb & 0x1FCu) | ring_sa); q1 = zsb_lds32v(((kb - 4u) & 0x1FCu) | ring_sa); q0 = zsb_lds32v(((kb - 8u) & 0x1FCu) | ring_sa);
            n0 = zsb_lds32v(((kb - 12u) & 0x1FCu) | ring_sa);
            T2 = zsb_fsl(q1, q2, o); T1 = zsb_fsl(q0, q1, o); T0 = zsb_fsl(n0, q0, o);
        }
        // One step of the chain.  On the chain: the three cell
        // loads, their sum, one funnel shift that skips the extra bits (the sum is its shift amount: byte 0 = extra bits, and bit 5
        // picks the word pair), one per state that isolates its bits, the new cell address: one shared-memory load and ~6 ALU
        // operations per sequence.  Off the chain: the window moves on by (extra + state bits) -- a dot product, two levels of
        // selects over the carried words (0..3 whole words), three loads that refill the look-ahead, three shifts for the new T.
#ifndef SEQ_DEP
#define SEQ_DEP 0
#endif
#define SEQ_STEP(i_)                                                                                                                        \
        {                                                                                                                                   \
            const uint32_t eL = zsb_lds32(aL), eM = zsb_lds32(aM), eO = zsb_lds32(aO);                                                      \
            /* (behind the cell loads, which are on the chain) the look-ahead words this step's window move may need */                     \
            const uint32_t kd = SEQ_DEP ? kb + (eL & zr) : kb; /* == kb, but (SEQ_DEP) only known once the cell is there */                 \
            n1 = zsb_lds32v(((kd - 16u) & 0x1FCu) | ring_sa); n2 = zsb_lds32v(((kd - 20u) & 0x1FCu) | ring_sa);                              \
            n3 = zsb_lds32v(((kd - 24u) & 0x1FCu) | ring_sa);                                                                               \
            const uint32_t sum = eL + eO + eM;                 /* byte 0: extra bits (<= 63), byte 1: state bits (<= 27) */                 \
            const uint32_t tA = zsb_fsl(T1, T2, sum), tB = zsb_fsl(T0, T1, sum);                                                            \
            const uint32_t t = (sum & 32u) ? tB : tA;          /* the state bits, top-aligned */                                            \
            const uint32_t nL = eL >> 8, nM = eM >> 8, nO = eO >> 8, nLM = (eL + eM) >> 8;    /* shift amounts: nb in bits 0..4 */         \
            const uint32_t bL = zsb_fsl(t, 0, nL), bM = zsb_fsl(zsb_fsl(0, t, nL), 0, nM), bO = zsb_fsl(zsb_fsl(0, t, nLM), 0, nO);         \
            aL = tbL + (ZSB_CELL_BASE(eL) + bL) * (SEQ_CHAINS * 4); aM = tbM + (ZSB_CELL_BASE(eM) + bM) * (SEQ_CHAINS * 4);                 \
            aO = tbO + (ZSB_CELL_BASE(eO) + bO) * (SEQ_CHAINS * 4);                                        /* sequence.rs:80-88 */          \
            wrow[(i_) & (WIN * NBUF - 1)] = seq_fast_word(eL, eO, eM, sum);                                                                \
            /* the cursor behind the extra bits of the last sequence (no state update follows it, sequence.rs:80): below the stream = over-read */ \
            if ((i_) + 1 == nseq) top = SEQ_TOP() - (int32_t)(sum & 0xFFu);                                                                  \
            /* the window moves on */                                                                                                       \
            const uint32_t o2 = (uint32_t)__dp4a((int)sum, 0x00000101, (int)o);                           /* o + extra bits + state bits */ \
            o = o2 & 31u;                                                                                                                   \
            kb -= (o2 >> 5) * 4u;                                                                                                           \
            if (o2 & 32u) { q2 = q1; q1 = q0; q0 = n0; n0 = n1; n1 = n2; n2 = n3; }                                                         \
            if (o2 & 64u) { q2 = q0; q1 = n0; q0 = n1; n0 = n2; }                                                                           \
            T2 = zsb_fsl(q1, q2, o); T1 = zsb_fsl(q0, q1, o); T0 = zsb_fsl(n0, q0, o);                                                      \
        }
#define SEQ_TOP() ((int32_t)(kb * 8u + 32u - o))
        for (uint32_t i0 = 0; i0 < maxn; i0 += WIN) {
            const uint32_t B = i0 / WIN;
#ifdef ZSB_SEQ_TIMING
            const long long tw0 = clock64();
#endif
            if (B >= NBUF) seq_bar_sync<NT>(BAR_FREE + B % NBUF);   // the phase-2 warps are done with window B-NBUF, whose ring slots window B overwrites
#ifdef ZSB_SEQ_TIMING
            t_wait += clock64() - tw0;
#endif
            // A chain runs every step of every window it has a sequence in: past its last sequence it walks on from valid states over
            // whatever the ring holds (nothing of that is used: phase 2 stops at nseq), so that no window needs a per-step test.  The
            // stream rings are topped up every 8 steps (8 x 90 bits + the 224 bits of look-ahead < one 128-byte line), by all lanes in
            // the same pass.
            if (i0 < nseq) {
                for (uint32_t i8 = i0; i8 < i0 + WIN; i8 += 8) {
                    sr_check<7, SEQ_PF>(R, SEQ_TOP() - 160);
#pragma unroll 8
                    for (uint32_t i = i8; i < i8 + 8; i++) SEQ_STEP(i)
                }
            }
            __threadfence_block();
            seq_bar_arrive<NT>(BAR_FULL + B % NBUF);                // window B is in the ring
        }
#undef SEQ_TOP
#undef SEQ_STEP
        SEQ_T(3);
#ifdef ZSB_SEQ_TIMING
        if (lane == 0 && blockIdx.x < 160) g_seq_timing[blockIdx.x][4] = t_wait;
#endif
        // an over-read shows as a cursor below the stream start; illegal codes are caught by phase 2, which sees every code
        if (active && top < startbit) S.final_rc[lane] = ZSB_NEEDS_SLOW;
    } else if constexpr (FUSED) {
        seqx_consume<HELPERS, WIN, NBUF>(smem + SEQ_TBL_BYTES + 256 * SEQ_CHAINS * 2 + ((sizeof(SeqShared) + 15) & ~15ull), S, src, base8, work, cnt, slow_list, A, warp, lane);
        return;
    } else {
        // ---- phase 2: this warp's chains, batch by batch as the producer delivers them
        const uint32_t h = warp - 1, c0 = h * CPH;     // this warp's chains: c0 .. c0 + CPH - 1
        const uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(S.tab);
        Seq2Carry C[CPH];
        uint32_t nsq[CPH];
#pragma unroll
        for (int k = 0; k < CPH; k++) {
            nsq[k] = S.nseq[c0 + k];
            C[k].top = (int64_t)S.top0[c0 + k]; C[k].lit_acc = 0; C[k].out_acc = 0; C[k].H = hist_identity(); C[k].bad = 0;
        }
        uint32_t nbat = 0;                                            // windows of the longest chain of the CTA: every warp takes part in every hand-over
#pragma unroll
        for (int c = 0; c < SEQ_CHAINS; c++) nbat = max(nbat, (S.nseq[c] + WIN - 1) / WIN);
        for (uint32_t b = 0; b < nbat; b++) {
#ifdef ZSB_SEQ_TIMING
            const long long tw0 = clock64();
#endif
            seq_bar_sync<NT>(BAR_FULL + b % NBUF);
#ifdef ZSB_SEQ_TIMING
            const long long tw1 = clock64(); t_wait += tw1 - tw0;
#endif
#pragma unroll
            for (int k = 0; k < CPH; k++) {
                if (b * WIN < nsq[k]) {
                    const uint4 ww = *reinterpret_cast<const uint4 *>(&S.words[c0 + k][(b % NBUF) * WIN + SEQ_PER_LANE * lane]);
                    const uint32_t wd[4] = {ww.x, ww.y, ww.z, ww.w};
#ifndef SEQ_NO_P2
                    seq2_window(base8, tab_sa, wd, b * WIN, nsq[k], S.regen[c0 + k], reinterpret_cast<uint64_t *>((uintptr_t)S.rec[c0 + k]), C[k], lane);
#else
                    C[k].lit_acc += wd[0] & 1;
#endif
                }
            }
            if (b + NBUF < nbat) seq_bar_arrive<NT>(BAR_FREE + b % NBUF);
#ifdef ZSB_SEQ_TIMING
            t_work += clock64() - tw1;
#endif
        }
#ifdef ZSB_SEQ_TIMING
        if (warp == 1 && lane == 0 && blockIdx.x < 160) { g_seq_timing[blockIdx.x][5] = t_wait; g_seq_timing[blockIdx.x][6] = t_work; g_seq_timing[blockIdx.x][7] = nbat; }
#endif
        __syncthreads();                           // the producer's final verdicts
#pragma unroll
        for (int k = 0; k < CPH; k++) {
            if (nsq[k] == 0) continue;
            const bool bad = __any_sync(FULL, C[k].bad) || S.final_rc[c0 + k] != ZSB_OK;
            if (lane == 0) {
                ZsbBlockWork &g = work[S.bi[c0 + k]];
                if (bad) { g.status = ZSB_NEEDS_SLOW; slow_list[atomicAdd(&cnt->n_slow, 1u)] = S.bi[c0 + k]; }
                else {
                    g.lit_used = C[k].lit_acc; g.out_size = C[k].out_acc + (S.regen[c0 + k] - C[k].lit_acc);
                    g.rep_out[0] = C[k].H.h0; g.rep_out[1] = C[k].H.h1; g.rep_out[2] = C[k].H.h2;
                }
            }
        }
        return;
    }
    __syncthreads();
}

// k_seq_slow: the careful decoder (zsb_seq.h, the reference's exact error order) for the blocks the fast path
// handed over.  One warp per CTA, one block per lane, tables interleaved across the 32 lanes.
#define SEQ_SLOW_SMEM_BYTES (3 * SEQ_TBL_CELLS * 32 * 4 + 256 * 32 * 2 + 96 * 4)
__global__ void __launch_bounds__(32, 1) k_seq_slow(const uint8_t *__restrict__ src, uint64_t src_len, ZsbBlockWork *work,
                                                    const uint32_t *__restrict__ slow_list, const ZsbCounters *__restrict__ cnt,
                                                    uint64_t *seq_pool) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (cnt->overflow) return;
    uint32_t *tbl = reinterpret_cast<uint32_t *>(smem);
    int16_t *counts = reinterpret_cast<int16_t *>(smem + 3 * SEQ_TBL_CELLS * 32 * 4);
    uint32_t *bases = reinterpret_cast<uint32_t *>(smem + 3 * SEQ_TBL_CELLS * 32 * 4 + 256 * 32 * 2);   // [0..35] LL, [36..88] ML
    const uint32_t lane = threadIdx.x;
    const uint32_t n = cnt->n_slow, idx = blockIdx.x * 32 + lane;
    if (blockIdx.x * 32 >= n) return;
    for (uint32_t k = lane; k < 36 + 53; k += 32) bases[k] = k < 36 ? zsb_ll_base(k) : zsb_ml_base(k - 36);
    __syncwarp();
    if (idx >= n) return;
    const uint32_t bi = slow_list[idx];
    ZsbBlockWork w = work[bi];
    SeqTables T;
    T.ts = 32; T.max_al[0] = T.max_al[1] = T.max_al[2] = 0;
    T.tbl[0] = tbl + lane; T.tbl[1] = tbl + SEQ_TBL_CELLS * 32 + lane; T.tbl[2] = tbl + 2 * SEQ_TBL_CELLS * 32 + lane;
    int rc = seq_build_tables(src, w, T, counts + lane, 32);
    if (!rc) rc = seq_decode(src, src_len, w, T, bases, bases + 36, seq_pool + w.seq_buf);
    ZsbBlockWork &g = work[bi];
    g.status = rc;
    if (rc) return;
    g.out_size = w.out_size; g.lit_used = w.lit_used;
    g.rep_out[0] = w.rep_out[0]; g.rep_out[1] = w.rep_out[1]; g.rep_out[2] = w.rep_out[2];
}

// ======================================================================================= k_plan2
__global__ void __launch_bounds__(1024) k_plan2(const zsb_frame *__restrict__ frames, uint32_t nf, const zsb_block *__restrict__ blocks,
                                                ZsbBlockWork *work, ZsbFrameOut *fout, ZsbCounters *cnt, uint64_t dst_cap, uint32_t flags,
                                                const uint64_t *__restrict__ pre_off) {
    __shared__ uint64_t s_warp[33];
    __shared__ uint64_t s_base, s_max;
    __shared__ uint32_t s_last;
    const uint32_t tid = threadIdx.x;
    if (cnt->overflow) return;
    // (a) per frame, all CTAs side by side: size and status
    for (uint32_t fb = blockIdx.x * blockDim.x + (tid & ~31u); fb < nf; fb += gridDim.x * blockDim.x) {
        const uint32_t f = fb + (tid & 31u);
        uint64_t len = 0;
        int st = f < nf ? fout[f].status : ZSB_E_ARG;
        const bool z = st == ZSB_OK && frames[f].kind == 0;
        const bool big = z && frames[f].n_blocks > PLAN_WARP_BLOCKS;
        uint32_t ea = 0, eb = 0;
        if (st == ZSB_OK && frames[f].kind == 1) len = (flags & ZSB_PRINT_SKIPPABLE) ? blocks[frames[f].first_block].size : 0;   // main.rs:45-49
        if (z && !big) st = plan_frame(frames[f], blocks, work, len, ea, eb);
        for (uint32_t m = __ballot_sync(FULL, big); m; m &= m - 1) {              // frames of many blocks: by the whole warp
            const uint32_t j = __ffs(m) - 1;
            uint64_t l2 = 0; uint32_t a2 = 0, b2 = 0;
            const int rc = plan_frame_warp(frames[fb + j], blocks, work, l2, a2, b2, tid & 31u);
            if ((tid & 31u) == j) { st = rc; len = l2; ea = a2; eb = b2; }
        }
        if (z) {
            if (st != ZSB_OK) { fout[f].err_a = ea; fout[f].err_b = eb; }
            if (st == ZSB_OK && frames[f].has_content_size && len != frames[f].content_size && !(flags & ZSB_REFERENCE_QUIRKS)) st = ZSB_E_CONTENT_SIZE;
            if (st != ZSB_OK) len = 0;
        }
        if (f < nf) { fout[f].dst_len = len; fout[f].status = st; }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) { s_last = atomicAdd(&cnt->ticket2, 1u) == gridDim.x - 1 ? 1u : 0u; s_base = 0; s_max = 0; }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // (b) the last CTA: output offsets by an exclusive scan, four consecutive frames per thread and pass
    for (uint32_t f0 = 0; f0 < nf; f0 += 4 * blockDim.x) {
        uint64_t len[4]; int st[4]; uint64_t sum = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t f = f0 + 4 * tid + j;
            len[j] = 0; st[j] = ZSB_OK;
            if (f < nf) { st[j] = __ldcg(&fout[f].status); len[j] = __ldcg(&fout[f].dst_len); }
            sum += len[j];
        }
        uint64_t tot;
        uint64_t off = cta_scan_excl(sum, s_warp, tot) + s_base;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t f = f0 + 4 * tid + j;
            if (f < nf) {
                if (st[j] == ZSB_OK && off + len[j] > dst_cap) st[j] = ZSB_E_DST_TOO_SMALL;
                fout[f].dst_off = off; fout[f].dst_len = (st[j] == ZSB_OK) ? len[j] : 0; fout[f].status = st[j];
                if (st[j] == ZSB_OK && len[j]) atomicMax((unsigned long long *)&s_max, (unsigned long long)(off + len[j]));
                // a block k_seqx has executed already lies where the host expected the frame to go
                if (pre_off && st[j] == ZSB_OK && frames[f].kind == 0 && frames[f].n_blocks && __ldcg(&work[frames[f].first_block].fused) && off != pre_off[f]) cnt->refuse = 1u;
            }
            off += len[j];
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
// Raw blocks and skippable payloads are copies, RLE blocks fills (block.rs:76-79, frame.rs:81).  One CTA per block; the 16-byte aligned
// body moves by bulk copies (cp.async.bulk): an RLE block from one shared-memory tile of the byte, a raw block whose source and
// destination share their alignment modulo 16 through two shared-memory tiles in alternation (HBM -> tile on an mbarrier, tile -> HBM
// in a bulk group); threads only touch the unaligned head and tail.  A raw block whose alignments differ is copied by the threads.
#define RR_TILE 16384u
__global__ void __launch_bounds__(256) k_rawrle(const uint8_t *__restrict__ src, const zsb_block *__restrict__ blocks,
                                                const ZsbBlockWork *__restrict__ work, const ZsbFrameOut *__restrict__ fout,
                                                const uint32_t *__restrict__ list, const ZsbCounters *__restrict__ cnt, uint8_t *dst) {
    __shared__ __align__(128) uint8_t tile[2][RR_TILE];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    if (cnt->overflow) return;
    const uint32_t bi = list[blockIdx.x];
    const zsb_block b = blocks[bi];
    const ZsbFrameOut fo = fout[b.frame];
    if (fo.status != ZSB_OK || fo.dst_len == 0 || b.size == 0) return;
    uint8_t *d = dst + fo.dst_off + work[bi].out_off;
    const uint32_t tid = threadIdx.x, n = b.size;
    uint32_t head = (16u - ((uint32_t)(uintptr_t)d & 15u)) & 15u; if (head > n) head = n;
    const uint32_t body = (n - head) & ~15u;
    if (b.type == ZSB_BT_RLE) {                                              // block.rs:77-79
        const uint8_t v = src[b.src_off];
        if (body >= 1024) {
            const uint32_t w = v * 0x01010101u, fill = body < RR_TILE ? body : RR_TILE;
            for (uint32_t i = tid; i < fill / 16; i += blockDim.x) reinterpret_cast<uint4 *>(tile[0])[i] = make_uint4(w, w, w, w);
            zsb_fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(tile[0]);
                for (uint32_t at = 0; at < body; at += RR_TILE) zsb_bulk_s2g(d + head + at, sa, body - at < RR_TILE ? body - at : RR_TILE);
                zsb_bulk_commit();
            }
            for (uint32_t i = tid; i < head; i += blockDim.x) d[i] = v;
            for (uint32_t i = head + body + tid; i < n; i += blockDim.x) d[i] = v;
            if (tid == 0) zsb_bulk_wait_read();                              // the tile is read before the CTA gives its shared memory back
        } else cta_fill_g(d, v, n);
        return;
    }
    const uint8_t *sp = src + b.src_off;                                     // block.rs:76 ; skippable payload frame.rs:81
    if (body < 2048 || (((uintptr_t)d ^ (uintptr_t)sp) & 15) != 0) { cta_copy_g2g(d, sp, n); return; }
    const uint32_t sa[2] = {(uint32_t)__cvta_generic_to_shared(tile[0]), (uint32_t)__cvta_generic_to_shared(tile[1])};
    const uint32_t ma[2] = {(uint32_t)__cvta_generic_to_shared(&s_mbar[0]), (uint32_t)__cvta_generic_to_shared(&s_mbar[1])};
    if (tid == 0) {
        zsb_mbar_init(ma[0], 1); zsb_mbar_init(ma[1], 1);
        const uint32_t nt = (body + RR_TILE - 1) / RR_TILE;
        uint32_t ph[2] = {0, 0};
        auto len_of = [&](uint32_t t) { return body - t * RR_TILE < RR_TILE ? body - t * RR_TILE : RR_TILE; };
        zsb_mbar_expect(ma[0], len_of(0)); zsb_bulk_g2s(sa[0], sp + head, len_of(0), ma[0]);
        for (uint32_t t = 0; t < nt; t++) {
            const uint32_t k = t & 1;
            if (t + 1 < nt) {
                if (t >= 1) zsb_bulk_wait_read();                            // tile k^1 was handed to a store one round ago: read by now
                zsb_mbar_expect(ma[k ^ 1], len_of(t + 1)); zsb_bulk_g2s(sa[k ^ 1], sp + head + (t + 1) * RR_TILE, len_of(t + 1), ma[k ^ 1]);
            }
            zsb_mbar_wait(ma[k], ph[k]); ph[k] ^= 1u;
            zsb_bulk_s2g(d + head + t * RR_TILE, sa[k], len_of(t)); zsb_bulk_commit();
        }
        zsb_bulk_wait_read();
    } else {
        for (uint32_t i = tid - 1; i < head; i += blockDim.x - 1) d[i] = sp[i];
        for (uint32_t i = head + body + tid - 1; i < n; i += blockDim.x - 1) d[i] = sp[i];
    }
}

// ======================================================================================= XXH64 primitives
#define XP1 0x9E3779B185EBCA87ull
#define XP2 0xC2B2AE3D27D4EB4Full
#define XP3 0x165667B19E3779F9ull
#define XP4 0x85EBCA77C2B2AE63ull
#define XP5 0x27D4EB2F165667C5ull
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xround(uint64_t acc, uint64_t in) { return rotl64(acc + in * XP2, 31) * XP1; }
__device__ __forceinline__ uint64_t xmerge(uint64_t h, uint64_t v) { return (h ^ xround(0, v)) * XP1 + XP4; }
// The accumulator chain is what bounds every XXH64 kernel here (round = rotl(acc + x * P2, 31) * P1, a dependent chain per accumulator), so it is
// carried in the form that makes a round three levels deep instead of the seven nvcc makes of the expression above: the state is the ROTATED sum
// r (acc = r * P1, P1 is odd: r0 = acc0 * P1^-1), and with a = x * P2 computed off the chain
//      level 1   (sl, h0) = rl * P1.lo + a   (one IMAD.WIDE with a 64-bit addend: the carry into the high word is free)
//                t = rl * P1.hi,  u = rh * P1.lo
//      level 2   sh = h0 + t + u             (ptxas makes two levels of it: u rides on t's IMAD, then one add; a forced three-input add -- vadd --
//                                             comes with PRMTs on the chain and is no shorter)
//      level 3   rl' = (sl:sh) >> 1,  rh' = (sh:sl) >> 1   (two funnel shifts: rotl 31 = swap halves, rotr 1)
// Measured in k_xxh_one (one warp, an SM to itself): 50 -> 24.8 cycles per round.
#define XP1_INV 0x887493432BADB37ull
static_assert(XP1_INV * XP1 == 1ull, "P1^-1 mod 2^64");
struct XAcc {
    uint32_t rl, rh;
    __device__ __forceinline__ void init(uint64_t acc) { const uint64_t r = acc * XP1_INV; rl = (uint32_t)r; rh = (uint32_t)(r >> 32); }
    __device__ __forceinline__ void round_a(uint64_t a) {                       // a = x * P2
        const uint64_t w = (uint64_t)rl * (uint32_t)XP1 + a;
        const uint32_t sl = (uint32_t)w;
        const uint32_t sh = (uint32_t)(w >> 32) + rl * (uint32_t)(XP1 >> 32) + rh * (uint32_t)XP1;
        rl = __funnelshift_r(sh, sl, 1);
        rh = __funnelshift_r(sl, sh, 1);
    }
    __device__ __forceinline__ void round(uint64_t x) { round_a(x * XP2); }
    // z: zero at run time, unknown to ptxas, which otherwise moves a's own product into the chain (two dependent IMAD.WIDE per round)
    __device__ __forceinline__ void round(uint64_t x, uint32_t z) { round_a((x * XP2) ^ z); }
    __device__ __forceinline__ uint64_t acc() const { return ((uint64_t)rh << 32 | rl) * XP1; }
};
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

// XXH64 of one frame by ONE warp that trails the executor of the same CTA (k_exec<.., true>): the frame's bytes are hashed from
// HBM/L2 as soon as the executor has committed them (*done = committed bytes, ~0 = the frame failed), so that the checksum of a
// frame of many blocks -- four dependent accumulator chains over all of its stripes, 33.5 M rounds for 1 GiB -- runs beside the
// execution instead of behind it.  All 32 lanes load: one instruction fetches 8 stripes (256 consecutive bytes, lane l the word
// l & 3 of stripe l >> 2) and 8 such loads are in flight (the chains are bound by their own latency only when ~64 stripes are on
// their way: a lane that loaded its own word of every stripe, 8 at a time, took 115 cycles per round); lanes 0..3 hold the
// accumulators and collect their words by shuffle.
struct WaveCtx;
__device__ void wave_advance(const WaveCtx &V);
__device__ __forceinline__ uint64_t xxh_finish(const uint8_t *p, uint64_t len, uint64_t nstripes, uint64_t v1, uint64_t v2, uint64_t v3, uint64_t v4);
__device__ __forceinline__ void xxh_trail(const uint8_t *p, uint64_t len, volatile unsigned long long *done, uint64_t *result, const WaveCtx *V = nullptr) {
    const uint32_t lane = threadIdx.x & 31, q = lane & 3;
    XAcc va; va.init(q == 0 ? XP1 + XP2 : q == 1 ? XP2 : q == 2 ? 0ull : 0ull - XP1);
    const uint32_t zero = blockIdx.y;                                  // (one-dimensional grids: see XAcc::round)
    const uint64_t nstripes = len >> 5;
    uint64_t cur = 0;
    constexpr int NR = 8;            // loads in flight per lane; 8 stripes each
    const uint8_t *g = p + 8 * lane;                                   // this lane's word of the first group of 8 stripes
    const uint32_t sh = (uint32_t)((uintptr_t)g & 7) * 8;
    const unsigned long long *ga = reinterpret_cast<const unsigned long long *>((uintptr_t)g & ~(uintptr_t)7);
    while (cur < nstripes) {
        // (lane 0's view for all lanes: the loop must not diverge, it shuffles)
        const unsigned long long d = __shfl_sync(FULL, *done, 0);
        if (d == ~0ull) return;
        uint64_t tgt = d >> 5; if (tgt > nstripes) tgt = nstripes;
        if (tgt < nstripes && tgt - cur < 8 * NR) { __nanosleep(500); if (V && lane == 0) wave_advance(*V); __syncwarp(); continue; }      // wait for a whole round of loads (or the end)
        __threadfence();
        // whole rounds of 8 * NR stripes (requesting the next round before this one is hashed, with a second register set, changes nothing:
        // the round is bound by the 64 dependent chain steps, ~65 cycles each beside 31 executing warps, not by its loads)
        while (cur + 8 * NR <= tgt) {
            uint64_t X[NR];
            const unsigned long long *gs = ga + 4 * cur;
            if (sh == 0) {
#pragma unroll
                for (int r = 0; r < NR; r++) X[r] = __ldcg(gs + 32 * r);
            } else {
#pragma unroll
                for (int r = 0; r < NR; r++) X[r] = (__ldcg(gs + 32 * r) >> sh) | (__ldcg(gs + 32 * r + 1) << (64 - sh));
            }
#pragma unroll
            for (int r = 0; r < NR; r++) {
#pragma unroll
                for (int j = 0; j < 8; j++) va.round(__shfl_sync(FULL, X[r], 4 * j + q), zero);
            }
            cur += 8 * NR;
        }
        // the rest of the stretch (only at the end of the frame), stripe by stripe
        if (tgt == nstripes) {
            for (; cur < tgt; cur++) va.round(ld64_any(p + (cur << 5) + 8 * q), zero);
        }
    }
    // everything is committed only once *done == len (the tail bytes)
    for (;;) {
        const unsigned long long d = __shfl_sync(FULL, *done, 0);
        if (d == ~0ull) return;
        if (d >= len) break;
        __nanosleep(500);
        if (V && lane == 0) wave_advance(*V);
        __syncwarp();
    }
    __threadfence();
    const uint64_t v = va.acc();
    const uint64_t v1 = __shfl_sync(FULL, v, 0), v2 = __shfl_sync(FULL, v, 1), v3 = __shfl_sync(FULL, v, 2), v4 = __shfl_sync(FULL, v, 3);
    if (lane == 0) *result = xxh_finish(p, len, nstripes, v1, v2, v3, v4);
}

// ======================================================================================= k_exec
// k_exec<512>: shards of the pipelined host path (it shares the SMs with other shards' k_seq: 1 024 threads of 60 registers would not fit
// beside one); k_exec<1024>: frames of many blocks on their own (twice the batches in flight: 0.18 -> 0.11 ms per block)
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

// ---- several CTAs per frame (k_exec in wavefront mode).  The blocks of a frame are handed out in order by a ticket, so the CTAs that run
// always hold the lowest unfinished blocks (no CTA waits for a block nobody runs); a block's entropy stages are long done, its place and
// its repeat offsets are known (k_plan2), and what it needs from the blocks before it is bytes: a match whose source reaches below the block
// start waits until the frame is committed to HBM up to the end of that source (front_pos: bytes committed contiguously from the frame's
// start; frame.rs:232-260 decodes block after block -- the result is the same bytes, the order of the work is not the reference's).
struct ZsbWave {
    unsigned long long front_pos;    // ~0: the frame failed, nobody waits any more
    uint32_t front_blk;              // blocks 0 .. front_blk-1 of the frame are committed
    uint32_t ticket;                 // next block to hand out
    uint32_t pad[4];
};
struct WaveCtx { ZsbWave *wf; uint32_t *done; const ZsbBlockWork *work; unsigned long long *s_front; uint32_t first, n_blocks; uint64_t frame_len; };   // done, work: of the frame's first block
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
// moves the frontier over every block that is marked done; called by whoever finishes a block and by whoever waits
__device__ __noinline__ void wave_advance(const WaveCtx &V) {
    for (;;) {
        const uint32_t fb = ld_acquire_u32(&V.wf->front_blk);
        if (fb >= V.n_blocks || !ld_acquire_u32(V.done + fb)) return;
        if (atomicCAS(&V.wf->front_blk, fb, fb + 1) == fb)
            atomicMax(&V.wf->front_pos, fb + 1 < V.n_blocks ? (unsigned long long)V.work[fb + 1].out_off : (unsigned long long)V.frame_len);
    }
}
// block k of the frame is in HBM
__device__ __forceinline__ void wave_commit(const WaveCtx &V, uint32_t k) {
    __threadfence();
    atomicExch(V.done + k, 1u);
    asm volatile("fence.sc.gpu;" ::: "memory");          // two CTAs that finish neighbouring blocks at once: at least one sees the other's mark
    wave_advance(V);
}
// is the frame committed up to byte `need`?  (the CTA's copy of the frontier first: most sources lie far below it)
__device__ __forceinline__ bool wave_ready(const WaveCtx &V, uint64_t need) {
    if (need <= *(volatile unsigned long long *)V.s_front) return true;
    unsigned long long fp = ld_acquire_u64(&V.wf->front_pos);
    if (need > fp) { wave_advance(V); fp = ld_acquire_u64(&V.wf->front_pos); }
    atomicMax(V.s_front, fp);
    return need <= fp;
}

// (defined with k_exec2: copies in units of up to 8 bytes through aligned words; here the "ring" is the block image, which never wraps)
template <uint32_t RING> __device__ __forceinline__ void ex2_store8(uint8_t *ring, uint32_t d, uint32_t v0, uint32_t v1, uint32_t n);
template <uint32_t RING> __device__ __forceinline__ void ex2_load8_ring(const uint8_t *ring, uint32_t s, uint32_t n, uint32_t &v0, uint32_t &v1);
#define EXEC_IMG 262144u         // power of two above the block image

// One batch of 32 consecutive sequences, one lane per sequence.  V (wavefront mode, else nullptr): sources below the block start are awaited.
__device__ __forceinline__ void exec_batch(uint32_t batch, uint32_t nseq, const uint64_t *__restrict__ seqs, uint8_t *o, uint32_t *bm,
                                           const LitSrc &L, const uint32_t *rep_in, uint64_t P0, const uint8_t *gblk, int *s_err, const WaveCtx *V) {
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
    uint8_t *img = reinterpret_cast<uint8_t *>((uintptr_t)o & ~(uintptr_t)15);     // the image starts at the alignment of its destination: o = img + ish
    const uint32_t ish = (uint32_t)((uintptr_t)o & 15);

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
    // wavefront mode: the frame position up to which earlier blocks must be committed for this match (0: none of its source lies there)
    uint64_t need = (V && pend && src < 0) ? P0 - (uint64_t)(uint32_t)(-min(src + (int)ml, 0)) : 0ull;
    for (uint32_t spins = 0;; spins++) {
        if (spins > (1u << 20)) { *s_err = ZSB_E_CORRUPT; break; }   // watchdog: a dependency that never resolves is a bug, not a hang
        bool didm = false;
        if (pend && !longM) {
            bool ready = (e <= (int)a0) || range_ready(bm, a0, (uint32_t)e);
            if (ready && need) { ready = wave_ready(*V, need); if (ready) need = 0; }
            if (ready) {
                __threadfence_block();
                if (src + (int)ml <= 0) {
                    // the whole source lies in earlier blocks of the frame (HBM/L2; in a frame with long-distance matches that is most
                    // of them): eight bytes are requested before any is stored, so that their latencies overlap -- byte by byte, each
                    // store would wait for its own load
                    for (uint32_t k = 0; k < ml; k += 8) {
                        uint8_t v[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) v[j] = k + j < ml ? __ldcg(gblk + src + (int)(k + j)) : (uint8_t)0;
#pragma unroll
                        for (int j = 0; j < 8; j++) if (k + j < ml) o[dstm + k + j] = v[j];
                    }
                } else if (src >= 0 && off >= 8) {
                    // inside the block, no unit reads what it writes: eight bytes at a time (three aligned words in, funnel shifts, predicated
                    // byte stores) -- byte by byte every byte pays a shared-memory round trip, and the batches behind this one wait for it
                    for (uint32_t k = 0; k < ml; k += 8) {
                        const uint32_t c = min(ml - k, 8u); uint32_t v0, v1;
                        ex2_load8_ring<EXEC_IMG>(img, ish + (uint32_t)src + k, c, v0, v1);
                        ex2_store8<EXEC_IMG>(img, ish + dstm + k, v0, v1, c);
                    }
                } else {
                    for (uint32_t k = 0; k < ml; k++) {
                        const int sp = src + (int)k;
                        o[dstm + k] = sp < 0 ? __ldcg(gblk + sp) : o[sp];
                    }
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
                const uint64_t jneed = __shfl_sync(FULL, need, j);
                if (jneed) { if (lane == 0) ok = wave_ready(*V, jneed); ok = __shfl_sync(FULL, (int)ok, 0) != 0; }
            }
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
        if (V && __all_sync(FULL, !pend || need != 0)) { __nanosleep(200); if (spins > (1u << 19)) spins = 1u << 19; }   // only the frontier is missing: no watchdog for that, other CTAs are at work
    }
}

// XXH: the last warp of the CTA does not execute but hashes the frame behind the executor (xxh_trail); the others synchronise among
// themselves with a named barrier.
template <int EXEC_THREADS, bool XXH>
__global__ void __launch_bounds__(EXEC_THREADS, 1) k_exec(const uint8_t *__restrict__ src, const zsb_frame *__restrict__ frames,
                                                          const zsb_block *__restrict__ blocks, const ZsbBlockWork *__restrict__ work,
                                                          ZsbFrameOut *fout, const uint32_t *__restrict__ exec_list,
                                                          const ZsbCounters *__restrict__ cnt, const uint64_t *__restrict__ seq_pool,
                                                          const uint8_t *__restrict__ lit_pool, uint8_t *dst, uint32_t flags,
                                                          ZsbWave *wave, uint32_t *blk_done, uint32_t G) {
    // wave != nullptr: G CTAs per frame (wavefront mode, see ZsbWave); else one CTA per frame, block after block
    constexpr uint32_t ET = EXEC_THREADS - (XXH ? 32 : 0);            // executing threads
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ int s_err;
    __shared__ unsigned long long s_done;                              // bytes of the frame committed to HBM, in order; ~0: the frame failed
    if (cnt->overflow) return;
    uint8_t *out_s = smem;
    uint32_t *bm = reinterpret_cast<uint32_t *>(smem + EXEC_OUT_BYTES);
    uint8_t *lit_s = smem + EXEC_OUT_BYTES + EXEC_BM_WORDS * 4;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t li = wave ? blockIdx.x / G : blockIdx.x;
    const uint32_t f = exec_list[li];
    const ZsbFrameOut fo = fout[f];
    if (fo.status != ZSB_OK) return;
    const zsb_frame fr = frames[f];
    uint8_t *fdst = dst + fo.dst_off;
    __shared__ unsigned long long s_front;                             // wavefront mode: this CTA's copy of the frame's frontier
    __shared__ uint32_t s_k;                                           // the block the CTA works on
    WaveCtx V;
    V.wf = wave ? wave + li : nullptr; V.done = blk_done ? blk_done + fr.first_block : nullptr; V.work = work + fr.first_block; V.s_front = &s_front;
    V.first = fr.first_block; V.n_blocks = fr.n_blocks; V.frame_len = fo.dst_len;
    __shared__ __align__(8) unsigned long long s_mbar;                 // completion of the literal staging's bulk copy
    const uint32_t mbar_sa = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    uint32_t mbar_phase = 0;
    if (tid == 0) { s_err = 0; s_done = 0; s_front = 0; s_k = 0; zsb_mbar_init(mbar_sa, 1); }
    __syncthreads();
    if (XXH && warp == ET / 32) {
        // (wavefront mode: the first of the frame's CTAs hashes, behind the frame's frontier instead of its own CTA's progress)
        if ((flags & ZSB_VERIFY_CHECKSUM) && fr.has_checksum && (!wave || blockIdx.x % G == 0))
            xxh_trail(fdst, fo.dst_len, wave ? &V.wf->front_pos : &s_done, &fout[f].xxh64, wave ? &V : nullptr);
        return;
    }
    // wavefront mode: the frame's first CTA does nothing but hash -- its one warp then has an SM to itself (beside 31 executing warps a
    // round of the four accumulator chains takes 73 cycles instead of ~35, and a frame of many blocks is bound by exactly that)
    if (wave && blockIdx.x % G == 0) return;
    auto sync_exec = [&]() { if (XXH) asm volatile("bar.sync 1, %0;" ::"r"(ET) : "memory"); else __syncthreads(); };
    bool failed = false;
    for (uint32_t kk = 0;; kk++) {
        uint32_t k = kk;
        if (wave) {                                                    // the next block of the frame nobody has taken yet
            sync_exec();
            if (tid == 0) s_k = ld_acquire_u64(&V.wf->front_pos) == ~0ull ? ~0u : atomicAdd(&V.wf->ticket, 1u);
            sync_exec();
            k = s_k;
        }
        if (k >= fr.n_blocks) break;
        const uint32_t bi = fr.first_block + k;
        const ZsbBlockWork &W = work[bi];
        if (blocks[bi].type != ZSB_BT_COMPRESSED) {                    // written by k_rawrle
            if (wave) { if (tid == 0) wave_commit(V, k); }
            else if (XXH && tid == 0) s_done = W.out_off + W.out_size;
            continue;
        }
        const uint32_t out_size = W.out_size, nseq = W.nseq, regen = W.lit_regen;
        uint8_t *gblk = fdst + W.out_off;
        const uint32_t shift = (uint32_t)((uintptr_t)gblk & 15);
        uint8_t *o = out_s + shift;
        // literal source
        LitSrc L; L.rle = 0; L.s = nullptr;
        L.g = W.lit_type == ZSB_LT_RAW ? src + W.lit_src : lit_pool + W.lit_buf;
        bool lit_wait = false;
        if (W.lit_type == ZSB_LT_RLE) { L.mode = 2; L.rle = src[W.lit_src]; }
        else if (regen <= EXEC_LIT_STAGE) {
            L.mode = 0;
            const uint32_t ls = (uint32_t)((uintptr_t)L.g & 15);
            L.s = lit_s + ls;
            // stage the literals: the 16-byte aligned body by ONE bulk copy (cp.async.bulk, the copy engine moves up to 64 KiB while the
            // threads clear the bitmap), head and tail bytes by the threads
            uint32_t head = (16 - ls) & 15; if (head > regen) head = regen;
            for (uint32_t i = tid; i < head; i += ET) lit_s[ls + i] = __ldg(L.g + i);
            const uint32_t nv = (regen - head) >> 4;
            if (nv) {
                if (tid == 0) { zsb_mbar_expect(mbar_sa, nv << 4); zsb_bulk_g2s((uint32_t)__cvta_generic_to_shared(lit_s + ls + head), L.g + head, nv << 4, mbar_sa); }
                lit_wait = true;
            }
            for (uint32_t i = head + (nv << 4) + tid; i < regen; i += ET) lit_s[ls + i] = __ldg(L.g + i);
        } else L.mode = 1;
        if (nseq) for (uint32_t i = tid; i < (out_size + 31) / 32; i += ET) bm[i] = 0;
        if (lit_wait) { zsb_mbar_wait(mbar_sa, mbar_phase); mbar_phase ^= 1u; }
        sync_exec();
        if (nseq == 0) {
            for (uint32_t i = tid; i < regen; i += ET) o[i] = lit_at(L, i);        // literals-only block (RFC; reference: Q1)
        } else {
            const uint64_t *seqs = seq_pool + W.seq_buf;
            const uint64_t lastrec = __ldg(seqs + nseq - 1);
            const uint32_t oe = (uint32_t)lastrec & ZSB_REC_POS_MASK, le = (uint32_t)(lastrec >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK;
            for (uint32_t i = tid; i < regen - le; i += ET) o[oe + i] = lit_at(L, le + i);   // decoding_context.rs:101-103
            const uint32_t nbatch = (nseq + 31) / 32;
            for (uint32_t b = warp; b < nbatch; b += ET / 32)
                exec_batch(b, nseq, seqs, o, bm, L, W.rep_in, fr.kind == 0 ? W.out_off : 0, gblk, &s_err, wave ? &V : nullptr);
        }
        zsb_fence_async_smem();                                        // the image was written with ordinary stores; the copy engine reads it
        sync_exec();
        // flush the block image to HBM: the image sits at the alignment of its destination, so its 16-byte aligned body leaves by ONE
        // bulk copy (cp.async.bulk shared -> global, up to 128 KiB) issued by one thread; head and tail bytes by the threads
        {
            uint32_t head = (16 - shift) & 15; if (head > out_size) head = out_size;
            const uint32_t nv = (out_size - head) >> 4;
            if (tid == 0 && nv) { zsb_bulk_s2g(gblk + head, (uint32_t)__cvta_generic_to_shared(o + head), nv << 4); zsb_bulk_commit(); }
            for (uint32_t i = tid; i < head; i += ET) gblk[i] = o[i];
            for (uint32_t i = head + (nv << 4) + tid; i < out_size; i += ET) gblk[i] = o[i];
            if (tid == 0 && nv) zsb_bulk_wait_all();                  // complete: the image may be overwritten, the bytes are in HBM
        }
        if (XXH || wave) __threadfence();                              // the block is in HBM before the hashing warp (other CTAs) are told so
        sync_exec();
        if (s_err) {
            if (tid == 0) { fout[f].status = s_err; fout[f].dst_len = 0; if (wave) atomicMax(&V.wf->front_pos, ~0ull); }   // (nobody waits for this frame any more)
            failed = true; break;
        }
        if (wave) { if (tid == 0) wave_commit(V, k); }
        else if (XXH && tid == 0) s_done = W.out_off + out_size;
    }
    if (!wave && XXH && tid == 0) s_done = failed ? ~0ull : fo.dst_len;
}

// ======================================================================================= k_exec2
// Sequence execution, one WARP per frame, blocks and sequences strictly in order (decoding_context.rs:78-106).
//
// Why a warp and not a CTA per frame.  Text at level 3 is ~15 000 sequences of ~8.5 bytes per 128 KiB block and
// every batch of 32 of them has a few matches whose source is only a few hundred bytes back, so the batches of
// one block form a dependency chain whatever the number of warps working on the block; spreading a block over a
// CTA only adds polling (k_exec, kept for frames of many blocks).  The chain is hidden across frames instead:
// the warp keeps the last EX2_RING bytes of its frame in shared memory (near sources, resolved with __syncwarp
// only), older sources are read back from HBM/L2 (they were flushed with 16-byte stores), and a SM holds 28
// such warps (all 4 096 frames of a C2 batch are resident at once).  The ring is small on purpose: 1, 2 and 4 KiB
// run equally fast, 8 KiB is twice as slow because the literal and record loads lose their L1.
//
// Positions are 32-bit and relative to the start of the current block (negative = earlier blocks of the frame);
// the ring index of a position is its global address modulo EX2_RING, so that frames and blocks continue
// seamlessly and 16-byte units of the ring and of HBM coincide.
#define EX2_RING 2048u           // k_exec2; the fused sequence kernel uses 1 KiB rings
#define EX2_MASK (EX2_RING - 1u)
#define EX2_WARPS 4
#define EX2_LONG 16u
#define EX2_GIANT (EX2_RING / 2)

struct Ex2Lit { const uint8_t *p; uint32_t rle; bool is_rle; };
__device__ __forceinline__ uint8_t ex2_lit(const Ex2Lit &L, uint32_t i) { return L.is_rle ? (uint8_t)L.rle : __ldg(L.p + i); }

// ring -> HBM for positions [lo, hi): bytes up to the first 16-byte boundary of the global address, 16-byte units, tail bytes
template <uint32_t RING = EX2_RING>
__device__ __forceinline__ void ex2_flush(const uint8_t *ring, uint8_t *gblk, uint32_t g0, int32_t lo, int32_t hi, uint32_t lane) {
    constexpr uint32_t EX2_MASK_ = RING - 1u;
    if (hi <= lo) return;
    int32_t a = lo + (int32_t)((16u - ((g0 + (uint32_t)lo) & 15u)) & 15u); if (a > hi) a = hi;
    for (int32_t p = lo + (int32_t)lane; p < a; p += 32) gblk[p] = ring[(g0 + (uint32_t)p) & EX2_MASK_];
    int32_t b = hi - (int32_t)((g0 + (uint32_t)hi) & 15u); if (b < a) b = a;
    for (int32_t p = a + 16 * (int32_t)lane; p < b; p += 512)
        *reinterpret_cast<uint4 *>(gblk + p) = *reinterpret_cast<const uint4 *>(ring + ((g0 + (uint32_t)p) & EX2_MASK_));
    for (int32_t p = b + (int32_t)lane; p < hi; p += 32) gblk[p] = ring[(g0 + (uint32_t)p) & EX2_MASK_];
}
// one already produced byte of the frame: from the ring if it is still there, else from HBM
template <uint32_t RING = EX2_RING>
__device__ __forceinline__ uint8_t ex2_src(const uint8_t *ring, const uint8_t *gblk, uint32_t g0, int32_t ring_lo, int32_t p) {
    return p >= ring_lo ? ring[(g0 + (uint32_t)p) & (RING - 1u)] : __ldcg(gblk + p);
}

// ---- copies in units of up to 8 bytes: three aligned source words, two funnel shifts, predicated byte stores
// store the low n (1..8) bytes of v1:v0 at ring position d (unmasked)
template <uint32_t RING>
__device__ __forceinline__ void ex2_store8(uint8_t *ring, uint32_t d, uint32_t v0, uint32_t v1, uint32_t n) {
    d &= RING - 1u;
    if (d + 8 <= RING) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(ring) + d;
        asm volatile(
            "{\n\t.reg .pred p1, p2, p3, p4, p5, p6, p7;\n\t.reg .b32 t;\n\t"
            "setp.gt.u32 p1, %3, 1;\n\tsetp.gt.u32 p2, %3, 2;\n\tsetp.gt.u32 p3, %3, 3;\n\tsetp.gt.u32 p4, %3, 4;\n\t"
            "setp.gt.u32 p5, %3, 5;\n\tsetp.gt.u32 p6, %3, 6;\n\tsetp.gt.u32 p7, %3, 7;\n\t"
            "st.shared.u8 [%0], %1;\n\t"
            "shr.u32 t, %1, 8;\n\t@p1 st.shared.u8 [%0+1], t;\n\t"
            "shr.u32 t, %1, 16;\n\t@p2 st.shared.u8 [%0+2], t;\n\t"
            "shr.u32 t, %1, 24;\n\t@p3 st.shared.u8 [%0+3], t;\n\t"
            "@p4 st.shared.u8 [%0+4], %2;\n\t"
            "shr.u32 t, %2, 8;\n\t@p5 st.shared.u8 [%0+5], t;\n\t"
            "shr.u32 t, %2, 16;\n\t@p6 st.shared.u8 [%0+6], t;\n\t"
            "shr.u32 t, %2, 24;\n\t@p7 st.shared.u8 [%0+7], t;\n\t}"
            ::"r"(a), "r"(v0), "r"(v1), "r"(n) : "memory");
    } else {
        const uint64_t v = ((uint64_t)v1 << 32) | v0;
        for (uint32_t t = 0; t < n; t++) ring[(d + t) & (RING - 1u)] = (uint8_t)(v >> (8 * t));
    }
}
// n (<= 8) bytes starting at ring position s (unmasked)
template <uint32_t RING>
__device__ __forceinline__ void ex2_load8_ring(const uint8_t *ring, uint32_t s, uint32_t n, uint32_t &v0, uint32_t &v1) {
    const uint32_t a = s & (RING - 1u) & ~3u, sh = (s & 3u) * 8u;
    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(ring + a);
    const uint32_t w1 = *reinterpret_cast<const uint32_t *>(ring + ((a + 4) & (RING - 1u)));
    const uint32_t w2 = *reinterpret_cast<const uint32_t *>(ring + ((a + 8) & (RING - 1u)));
    v0 = __funnelshift_r(w0, w1, sh); v1 = __funnelshift_r(w1, w2, sh);
}
// n (<= 16) bytes starting at global address g, in two steps so that the loads of several sources are in flight together:
// ex2_issue requests the aligned words that hold a wanted byte, ex2_unit shifts unit u (bytes 8u .. 8u+7) into place
struct ExRaw { uint32_t w0, w1, w2, w3, w4, sh; };
__device__ __forceinline__ void ex2_issue(const uint8_t *g, uint32_t n, ExRaw &r, bool cg) {
    const uint32_t m = (uint32_t)(uintptr_t)g & 3u;
    const uint32_t *a = reinterpret_cast<const uint32_t *>(g - m);
    r.sh = m * 8u; r.w1 = r.w2 = r.w3 = r.w4 = 0;
    if (cg) {
        r.w0 = __ldcg(a);
        if (m + n > 4) r.w1 = __ldcg(a + 1);
        if (m + n > 8) r.w2 = __ldcg(a + 2);
        if (m + n > 12) r.w3 = __ldcg(a + 3);
        if (m + n > 16) r.w4 = __ldcg(a + 4);
    } else {
        r.w0 = __ldg(a);
        if (m + n > 4) r.w1 = __ldg(a + 1);
        if (m + n > 8) r.w2 = __ldg(a + 2);
        if (m + n > 12) r.w3 = __ldg(a + 3);
        if (m + n > 16) r.w4 = __ldg(a + 4);
    }
}
__device__ __forceinline__ void ex2_unit(const ExRaw &r, int u, uint32_t &v0, uint32_t &v1) {
    if (u == 0) { v0 = __funnelshift_r(r.w0, r.w1, r.sh); v1 = __funnelshift_r(r.w1, r.w2, r.sh); }
    else { v0 = __funnelshift_r(r.w2, r.w3, r.sh); v1 = __funnelshift_r(r.w3, r.w4, r.sh); }
}

// per-frame state of a warp that executes sequences (k_exec2, and the consumer warps of the fused sequence kernel)
struct Ex2State {
    uint8_t *ring;              // RING bytes of shared memory, index = global address & (RING - 1)
    uint8_t *gblk;              // global address of the current block's first byte; positions are relative to it
    uint32_t g0;                // low 32 bits of gblk
    int32_t flushed, ring_lo;   // [ring_lo, ..) is in the ring, [.., flushed) is in HBM; ring_lo <= flushed always
    uint32_t c_out, c_lit;      // output / literal position behind the last item executed
    ExRaw LR; bool lr_has;      // literal words of the next batch, requested during the current one
    int err;
};
// One batch of (up to) 32 consecutive items of a block, one per lane: a sequence {out_end, lit_end: positions behind it, off: its resolved
// offset}, or the literal tail of the block (is_seq false, out_end = the block's size, lit_end = its literal count); lanes without an item
// repeat the ends of the last one.  has_next / next_lit_end: the lit_end of the same lane's item in the next batch (its literals are
// requested now).  P0: frame-relative position of the block.  == DecodingContext::execute_sequences for these items (decoding_context.rs:78-106).
// GIANT: the largest batch (bytes) that goes through the ring; <= RING - 16, so that everything below the ring's floor has been flushed
template <uint32_t RING, uint32_t GIANT>
__device__ __forceinline__ void ex2_batch(Ex2State &X, const Ex2Lit &L, uint64_t P0, bool is_seq, uint32_t out_end, uint32_t lit_end, uint32_t off,
                                          bool has_next, uint32_t next_lit_end, uint32_t lane) {
                uint32_t p_out = __shfl_up_sync(FULL, out_end, 1), p_lit = __shfl_up_sync(FULL, lit_end, 1);
                if (lane == 0) { p_out = X.c_out; p_lit = X.c_lit; }
                X.c_out = __shfl_sync(FULL, out_end, 31); X.c_lit = __shfl_sync(FULL, lit_end, 31);
                const uint32_t ll = lit_end - p_lit;
                uint32_t ml = out_end - p_out - ll;
                                const uint32_t dstm = p_out + ll;
                // decoding_context.rs:86-90; in 32 bits: the frames this executor takes are far below 2 GiB, beyond that every 31-bit offset has its source
                const uint32_t P0s = P0 > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)P0;
                if (is_seq && (off == 0 || off > P0s + dstm)) { X.err = ZSB_E_IMPOSSIBLE_VALUE; ml = 0; }
                const int32_t srcp = (int32_t)dstm - (int32_t)off;
                const uint32_t B0 = __shfl_sync(FULL, p_out, 0), B1 = X.c_out;
                if (B1 - B0 > GIANT) {
                    // a batch too large for the ring (a very long literal run or match): sequence by sequence, straight to HBM
                    ex2_flush<RING>(X.ring, X.gblk, X.g0, X.flushed, (int32_t)B0, lane);
                    __syncwarp();
                    for (int j = 0; j < 32; j++) {
                        const uint32_t jo = __shfl_sync(FULL, p_out, j), jll = __shfl_sync(FULL, ll, j), jml = __shfl_sync(FULL, ml, j);
                        const uint32_t jl = __shfl_sync(FULL, p_lit, j), joff = __shfl_sync(FULL, off, j);
                        const int32_t js = __shfl_sync(FULL, srcp, j);
                        const uint32_t jd = jo + jll;
                        for (uint32_t q = lane; q < jll; q += 32) X.gblk[jo + q] = ex2_lit(L, jl + q);
                        __syncwarp();
                        if (joff >= 32) {
                            for (uint32_t k0 = 0; k0 < jml; k0 += 32) {           // 32 bytes at a time: a unit never reads what it writes
                                if (k0 + lane < jml) X.gblk[jd + k0 + lane] = __ldcg(X.gblk + js + (int32_t)(k0 + lane));
                                __syncwarp();
                            }
                        } else {
                            for (uint32_t q = lane; q < jml; q += 32) X.gblk[jd + q] = __ldcg(X.gblk + js + (int32_t)(q % joff));   // periodic with period off
                        }
                        __syncwarp();
                    }
                    X.flushed = X.ring_lo = (int32_t)B1;
                    X.lr_has = false;                 // nothing was requested for the next batch
                    return;
                }
                X.ring_lo = max(X.ring_lo, (int32_t)B1 - (int32_t)RING);
                // ---- the HBM reads of the batch are requested first, so that their latencies overlap: the sources of short matches
                // that lie entirely in HBM (<= 16 bytes: two units; they depend on nothing in this batch) ...
                const int32_t send = srcp + (int32_t)ml;
                const bool farm = ml && ml <= EX2_LONG && send <= X.ring_lo && !(off < 8 && off < ml);
                ExRaw FR;
                if (farm) ex2_issue(X.gblk + srcp, ml, FR, true);
                // ... and the literals of the NEXT batch (those of this batch were requested one batch ago)
                ExRaw NR; bool n_has = false;
                if (has_next && !L.is_rle) {
                    uint32_t n_beg = __shfl_up_sync(FULL, next_lit_end, 1);
                    if (lane == 0) n_beg = X.c_lit;
                    const uint32_t n_ll = next_lit_end - n_beg;
                    n_has = n_ll && n_ll <= EX2_LONG;
                    if (n_has) ex2_issue(L.p + n_beg, n_ll, NR, false);
                }
                // ---- literals (no dependency on earlier output, decoding_context.rs:92-93)
                if (ll && ll <= EX2_LONG) {
                    if (L.is_rle) { const uint32_t v = L.rle * 0x01010101u; for (uint32_t q = 0; q < ll; q += 8) ex2_store8<RING>(X.ring, X.g0 + p_out + q, v, v, min(ll - q, 8u)); }
                    else {
                        if (!X.lr_has) ex2_issue(L.p + p_lit, ll, X.LR, false);           // first batch of a block
                        uint32_t v0, v1;
                        ex2_unit(X.LR, 0, v0, v1); ex2_store8<RING>(X.ring, X.g0 + p_out, v0, v1, min(ll, 8u));
                        if (ll > 8) { ex2_unit(X.LR, 1, v0, v1); ex2_store8<RING>(X.ring, X.g0 + p_out + 8, v0, v1, ll - 8); }
                    }
                }
                X.LR = NR; X.lr_has = n_has;
                for (uint32_t m = __ballot_sync(FULL, ll > EX2_LONG); m; m &= m - 1) {
                    const int j = __ffs(m) - 1;
                    const uint32_t jo = __shfl_sync(FULL, p_out, j), jll = __shfl_sync(FULL, ll, j), jl = __shfl_sync(FULL, p_lit, j);
                    for (uint32_t q = lane; q < jll; q += 32) X.ring[(X.g0 + jo + q) & (RING - 1u)] = ex2_lit(L, jl + q);
                }
                __syncwarp();
                // ---- matches (decoding_context.rs:95-98): a lane goes once everything it needs from other sequences is written,
                // i.e. lies below the match start of the lowest sequence still pending
                if (farm) {
                    uint32_t v0, v1;
                    ex2_unit(FR, 0, v0, v1); ex2_store8<RING>(X.ring, X.g0 + dstm, v0, v1, min(ml, 8u));
                    if (ml > 8) { ex2_unit(FR, 1, v0, v1); ex2_store8<RING>(X.ring, X.g0 + dstm + 8, v0, v1, ml - 8); }
                }
                __syncwarp();
                bool pend = ml != 0 && !farm;
                const int32_t need = min(send, (int32_t)p_out);
                for (;;) {
                    const uint32_t pm = __ballot_sync(FULL, pend);
                    if (!pm) break;
                    const int32_t done = (int32_t)__shfl_sync(FULL, dstm, __ffs(pm) - 1);
                    const bool ready = pend && need <= done;
                    if (ready && ml <= EX2_LONG) {
                        if (off < 8 && off < ml) {                     // overlapping with a short period: byte by byte
                            for (uint32_t q = 0; q < ml; q++) X.ring[(X.g0 + dstm + q) & (RING - 1u)] = ex2_src<RING>(X.ring, X.gblk, X.g0, X.ring_lo, srcp + (int32_t)q);
                        } else if (srcp >= X.ring_lo) {                  // source still in the X.ring; a unit never reads what it writes (off >= 8)
                            for (uint32_t q = 0; q < ml; q += 8) {
                                const uint32_t c = min(ml - q, 8u); uint32_t v0, v1;
                                ex2_load8_ring<RING>(X.ring, X.g0 + (uint32_t)srcp + q, c, v0, v1);
                                ex2_store8<RING>(X.ring, X.g0 + dstm + q, v0, v1, c);
                            }
                        } else {                                       // straddles the X.ring floor (sources entirely in HBM were done above)
                            for (uint32_t q = 0; q < ml; q++) X.ring[(X.g0 + dstm + q) & (RING - 1u)] = ex2_src<RING>(X.ring, X.gblk, X.g0, X.ring_lo, srcp + (int32_t)q);
                        }
                        pend = false;
                    }
                    for (uint32_t m = __ballot_sync(FULL, ready && ml > EX2_LONG); m; m &= m - 1) {
                        const int j = __ffs(m) - 1;
                        const uint32_t jml = __shfl_sync(FULL, ml, j), jd = __shfl_sync(FULL, dstm, j), joff = __shfl_sync(FULL, off, j);
                        const int32_t js = __shfl_sync(FULL, srcp, j);
                        if (joff >= 32) {
                            for (uint32_t k0 = 0; k0 < jml; k0 += 32) {
                                if (k0 + lane < jml) X.ring[(X.g0 + jd + k0 + lane) & (RING - 1u)] = ex2_src<RING>(X.ring, X.gblk, X.g0, X.ring_lo, js + (int32_t)(k0 + lane));
                                __syncwarp();
                            }
                        } else {
                            for (uint32_t q = lane; q < jml; q += 32) X.ring[(X.g0 + jd + q) & (RING - 1u)] = ex2_src<RING>(X.ring, X.gblk, X.g0, X.ring_lo, js + (int32_t)(q % joff));
                        }
                        if ((int)lane == j) pend = false;
                    }
                    __syncwarp();
                }
                // ---- whole 16-byte units go to HBM -- not after every batch: a batch of text regenerates ~270 bytes, 17 units for 32 lanes, and the
                // bookkeeping costs the same for 17 as for 128.  Bytes may stay in the ring as long as the next batch still fits behind them
                // (it regenerates at most GIANT bytes, and everything below ring_lo = B1 - RING must be in HBM): the ring is flushed when more than
                // RING - GIANT bytes are pending, and at the end of a block.
                const int32_t hi = (int32_t)B1 - (int32_t)((X.g0 + B1) & 15u);
                if (hi > X.flushed && (!has_next || (int32_t)B1 - X.flushed > (int32_t)(RING - GIANT))) {
                    if (((X.g0 + (uint32_t)X.flushed) & 15u) == 0) {           // the usual case: whole 16-byte units only, at most a few per lane
                        for (int32_t p = X.flushed + 16 * (int32_t)lane; p < hi; p += 512)
                            *reinterpret_cast<uint4 *>(X.gblk + p) = *reinterpret_cast<const uint4 *>(X.ring + ((X.g0 + (uint32_t)p) & (RING - 1u)));
                    } else ex2_flush<RING>(X.ring, X.gblk, X.g0, X.flushed, hi, lane);
                    X.flushed = hi;
                    __syncwarp();               // the next batch may read these bytes back from HBM through other lanes
                }
}

// The consumer warps of k_seqx (k_seq_t<.., FUSED = true>): warp c + 1 takes chain c.  Per window of 32 sequences: phase 2 with one
// sequence per lane (seq2_records<1>), then -- if the block's place is known -- the batch BEFORE it is executed (ex2_batch wants the
// literal ends of the batch that follows, to request its literals early); else the records go to HBM as in k_seq.
// Nothing is executed once phase 2 has seen anything illegal (the careful decoder redoes the block into records, k_exec2 executes them);
// what was written until then lies inside the frame's own place.  An output that would leave that place (a frame that regenerates
// more than its Frame_Content_Size) stops the execution and sets cnt->refuse: the host runs the batch again without k_seqx.
template <int HELPERS, int WIN, int NBUF>
__device__ __forceinline__ void seqx_consume(uint8_t *smem_rings, SeqSharedT<WIN * NBUF + 4> &S, const uint8_t *src, const uint8_t *base8, ZsbBlockWork *work,
                                             ZsbCounters *cnt, uint32_t *slow_list, const SeqxArgs &A, uint32_t warp, uint32_t lane) {
    static_assert(WIN == 32, "one sequence per lane and window");
    constexpr int NT = 32 * (1 + HELPERS);
    constexpr uint32_t BAR_FULL = 1u, BAR_FREE = 1u + NBUF;
    const uint32_t c = warp - 1;
    const uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(S.tab);
    const uint32_t nsq = S.nseq[c], regen = S.regen[c], bi = S.bi[c], limit = S.limit[c];
    Seq2Carry C;
    C.top = (int64_t)S.top0[c]; C.lit_acc = 0; C.out_acc = 0; C.H = hist_identity(); C.bad = 0;
    uint32_t nbat = 0;                                                // windows of the longest chain of the CTA: every warp takes part in every hand-over
#pragma unroll
    for (int k = 0; k < SEQ_CHAINS; k++) nbat = max(nbat, (S.nseq[k] + WIN - 1) / WIN);
    uint8_t *gdst = reinterpret_cast<uint8_t *>((uintptr_t)S.gdst[c]);
    uint64_t *rec = reinterpret_cast<uint64_t *>((uintptr_t)S.rec[c]);
    const bool placed = gdst != nullptr && nsq != 0;
    bool exec = placed, refuse = false, xerr = false;
    Ex2State X; Ex2Lit L;
    X.ring = smem_rings + c * SEQX_RING; X.gblk = gdst; X.g0 = (uint32_t)(uintptr_t)gdst; X.flushed = 0; X.ring_lo = 0;
    X.c_out = 0; X.c_lit = 0; X.lr_has = false; X.err = 0;
    L.p = nullptr; L.rle = 0; L.is_rle = false;
    if (placed) {
        const ZsbBlockWork &W = work[bi];
        L.is_rle = W.lit_type == ZSB_LT_RLE; L.rle = L.is_rle ? src[W.lit_src] : 0u;
        L.p = W.lit_type == ZSB_LT_RAW ? src + W.lit_src : A.lit_pool + W.lit_buf;
    }
    const uint32_t rep0[3] = {1u, 4u, 8u};                            // the history a frame starts with (decoding_context.rs:40)
    const uint32_t items = nsq + 1;                                   // the last item is the literal tail
    const uint32_t nloop = placed ? max(nbat, (items + 31) / 32 + 1) : nbat;
    uint32_t c_oe = 0, c_le = 0, c_off = 1; bool c_seq = false;       // the batch waiting to be executed: out_end, lit_end, offset, sequence?
    for (uint32_t B = 0; B < nloop; B++) {
        uint32_t n_oe = 0, n_le = 0, n_off = 1; bool n_seq = false;
        if (B < nbat) {
            seq_bar_sync<NT>(BAR_FULL + B % NBUF);
            if (B * WIN < nsq) {
                const uint32_t wd[1] = {S.words[c][(B % NBUF) * WIN + lane]};
                bool v[1]; uint64_t r[1];
                seq2_records<1>(base8, tab_sa, wd, B * WIN, nsq, regen, C, lane, v, r);
                if (!placed) { if (v[0]) rec[B * WIN + lane] = r[0]; }
                else if (v[0]) {
                    n_seq = true; n_oe = (uint32_t)r[0] & ZSB_REC_POS_MASK; n_le = (uint32_t)(r[0] >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK;
                    n_off = seq_real_offset((uint32_t)(r[0] >> (2 * ZSB_REC_POS_BITS)), rep0);
                }
            }
            if (B + NBUF < nbat) seq_bar_arrive<NT>(BAR_FREE + B % NBUF);
        }
        if (!placed) continue;
        if (!n_seq) { n_le = regen; n_oe = C.out_acc + (regen - C.lit_acc); }          // behind the last sequence: the literal tail, then nothing
        if (exec && __any_sync(FULL, C.bad != 0)) exec = false;
        if (exec && __any_sync(FULL, n_oe > limit)) { exec = false; refuse = true; }
        if (exec && B >= 1 && (B - 1) * 32 < items) {
            ex2_batch<SEQX_RING, SEQX_RING - 32>(X, L, 0, c_seq, c_oe, c_le, c_off, B * 32 < items, n_le, lane);
            if (__any_sync(FULL, X.err != 0)) { exec = false; xerr = true; }
        }
        c_oe = n_oe; c_le = n_le; c_off = n_off; c_seq = n_seq;
    }
    __syncthreads();                           // the producer's final verdicts
    if (nsq == 0) return;
    const bool bad = __any_sync(FULL, C.bad != 0) || S.final_rc[c] != ZSB_OK;
    const uint32_t out_size = C.out_acc + (regen - C.lit_acc);
    if (exec && !bad) {                        // what is left in the ring
        uint8_t *gend = gdst + out_size;
        ex2_flush<SEQX_RING>(X.ring, gend, (uint32_t)(uintptr_t)gend, X.flushed - (int32_t)out_size, 0, lane);
    }
    if (lane == 0) {
        ZsbBlockWork &g = work[bi];
        if (bad) { g.status = ZSB_NEEDS_SLOW; slow_list[atomicAdd(&cnt->n_slow, 1u)] = bi; }
        else {
            g.lit_used = C.lit_acc; g.out_size = out_size;
            g.rep_out[0] = C.H.h0; g.rep_out[1] = C.H.h1; g.rep_out[2] = C.H.h2;
            // (refused: no records were written for this block and nothing valid lies in dst -- k_exec2 must not touch it; the host runs the batch again)
            if (xerr) g.fused = 2; else if (exec) g.fused = 1; else if (refuse) { g.fused = 3; cnt->refuse = 1u; }
        }
    }
}

__global__ void __launch_bounds__(32 * EX2_WARPS, 7) k_exec2(const uint8_t *__restrict__ src, const zsb_frame *__restrict__ frames,
                                                             const zsb_block *__restrict__ blocks, const ZsbBlockWork *__restrict__ work,
                                                             ZsbFrameOut *fout, const uint32_t *__restrict__ exec_list, uint32_t n,
                                                             const ZsbCounters *__restrict__ cnt, const uint64_t *__restrict__ seq_pool,
                                                             const uint8_t *__restrict__ lit_pool, uint8_t *dst) {
    __shared__ __align__(128) uint8_t rings[EX2_WARPS][EX2_RING];
    if (cnt->overflow) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * EX2_WARPS + warp;
    if (gw >= n) return;
    Ex2State X;
    X.ring = rings[warp]; X.err = 0; X.flushed = 0; X.ring_lo = 0;
    const uint32_t f = exec_list[gw];
    const ZsbFrameOut fo = fout[f];
    if (fo.status != ZSB_OK) return;
    const zsb_frame fr = frames[f];
    uint8_t *fdst = dst + fo.dst_off;
    for (uint32_t k = 0; k < fr.n_blocks && !X.err; k++) {
        const uint32_t bi = fr.first_block + k;
        const ZsbBlockWork &W = work[bi];
        const uint32_t out_size = W.out_size;
        X.gblk = fdst + W.out_off; X.g0 = (uint32_t)(uintptr_t)X.gblk;
        if (blocks[bi].type != ZSB_BT_COMPRESSED || W.fused) {
            if (W.fused == 2) { X.err = ZSB_E_IMPOSSIBLE_VALUE; break; }
            ex2_flush(X.ring, X.gblk, X.g0, X.flushed, 0, lane);       // what earlier blocks left in the ring
            X.flushed = X.ring_lo = (int32_t)out_size;                 // raw / RLE blocks were written by k_rawrle, blocks executed by k_seqx are there too
        } else {
            const uint32_t nseq = W.nseq, regen = W.lit_regen;
            Ex2Lit L;
            L.is_rle = W.lit_type == ZSB_LT_RLE; L.rle = L.is_rle ? src[W.lit_src] : 0u;
            L.p = W.lit_type == ZSB_LT_RAW ? src + W.lit_src : lit_pool + W.lit_buf;
            const uint64_t *seqs = seq_pool + W.seq_buf;
            const uint64_t P0 = W.out_off;
            const uint32_t items = nseq + 1;   // the last item is the literal tail (decoding_context.rs:101-103); all there is when nseq == 0
            X.c_out = 0; X.c_lit = 0; X.lr_has = false;
            // records are requested two batches ahead: the batch after this one is looked at early (its literals are prefetched)
            uint64_t rec_next = lane < nseq ? __ldg(seqs + lane) : 0ull, rec_next2 = lane + 32 < nseq ? __ldg(seqs + lane + 32) : 0ull;
            const uint64_t *rec_ahead = seqs + lane + 64;
            for (uint32_t b0 = 0; b0 < items; b0 += 32, rec_ahead += 32) {
                const uint32_t i = b0 + lane;
                const uint64_t rec = rec_next;
                rec_next = rec_next2;
                if (i + 64 < nseq) rec_next2 = __ldg(rec_ahead);
                const bool is_seq = i < nseq;
                const uint32_t out_end = is_seq ? (uint32_t)rec & ZSB_REC_POS_MASK : out_size;
                const uint32_t lit_end = is_seq ? (uint32_t)(rec >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK : regen;
                const uint32_t off = is_seq ? seq_real_offset((uint32_t)(rec >> (2 * ZSB_REC_POS_BITS)), W.rep_in) : 1u;
                const bool has_next = b0 + 32 < items;
                const uint32_t next_lit_end = (i + 32 < nseq) ? (uint32_t)(rec_next >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK : regen;
                ex2_batch<EX2_RING, EX2_GIANT>(X, L, P0, is_seq, out_end, lit_end, off, has_next, next_lit_end, lane);
                if (__any_sync(FULL, X.err != 0)) { X.err = ZSB_E_IMPOSSIBLE_VALUE; break; }
            }
        }
        X.flushed -= (int32_t)out_size; X.ring_lo -= (int32_t)out_size;   // positions become relative to the next block
        __syncwarp();
    }
    if (__any_sync(FULL, X.err != 0)) { if (lane == 0) { fout[f].status = ZSB_E_IMPOSSIBLE_VALUE; fout[f].dst_len = 0; } return; }
    uint8_t *gend = fdst + fo.dst_len;
    ex2_flush(X.ring, gend, (uint32_t)(uintptr_t)gend, X.flushed, 0, lane);
}

// ======================================================================================= k_link_init / k_link_resolve
// Sequence execution of frames of many blocks (decoding_context.rs:78-106) WITHOUT walking the frame in order.  execute_sequences is one
// dependent chain through the whole frame -- a match may copy bytes the sequence before it wrote -- and on text at zstd -3 half the
// sequences of a block depend, directly or through bytes that do, on the block before it (tools/probes/c3_dependencies.py), so a CTA per
// frame (k_exec) runs at 0.12 ms per block whatever the number of CTAs.  What IS parallel is the question "which literal does this output
// byte come from": every byte of the frame gets one 32-bit ENTRY in a scratch array,
//      resolved    LINK_RES | byte value          (a literal, a raw / RLE block byte, or a match byte whose source is known)
//      unresolved  index of its source byte       (frame relative, always smaller than its own index)
// k_link_init writes the entries of all blocks side by side (nothing depends on anything: the records say where every sequence starts),
// and k_link_resolve replaces entry[i] by entry[entry[i]] until it is resolved -- pointer jumping, in place and without any barrier:
// an entry only ever holds a valid ancestor or the final byte, 32-bit accesses are single copies, so a reader that sees an old value
// merely walks one hop more.  Nobody waits for anybody (a lane follows its chain itself if nobody shortened it), so there is nothing to
// deadlock on; chunks are handed out in frame order by ticket, so the sources of most matches were resolved and written back (one hop)
// by the time a chunk is taken.  The bytes go to dst in 32-bit words, 128 bytes per warp instruction.
// HBM: 4 bytes of entry written and read back per output byte plus the chain hops, instead of 0.12 ms per 128 KiB block in order.
#define LINK_RES 0x80000000u
#define LINK_THREADS 256
#define LINK_WIDE 4096u                       // k_link_init: a batch of 32 sequences that regenerates more bytes than this is shared by the CTA's warps
#define LINK_WIDE_MAX 64
#define LINK_CHUNK (LINK_THREADS * 16u)       // entries (output bytes) per ticket of k_link_resolve: every lane takes 16
struct ZsbLinkFrame { uint64_t e_off; uint32_t frame, pad; };   // entries of listed frame li start at ent + e_off; entry j <-> dst byte (dst_off & ~15) + j

__device__ __forceinline__ void link_fail(ZsbFrameOut *fo, int code) { atomicCAS(&fo->status, (int)ZSB_OK, code); }

__global__ void __launch_bounds__(LINK_THREADS) k_link_init(const uint8_t *__restrict__ src, const zsb_block *__restrict__ blocks,
                                                            const ZsbBlockWork *__restrict__ work, ZsbFrameOut *fout,
                                                            const uint2 *__restrict__ link_blocks, const ZsbLinkFrame *__restrict__ link_frames,
                                                            const ZsbCounters *__restrict__ cnt, const uint64_t *__restrict__ seq_pool,
                                                            const uint8_t *__restrict__ lit_pool, uint32_t *ent) {
    if (cnt->overflow) return;
    const uint2 it = link_blocks[blockIdx.x];
    const uint32_t bi = it.x;
    const ZsbLinkFrame lf = link_frames[it.y];
    ZsbFrameOut *fo = fout + lf.frame;
    if (fo->status != ZSB_OK) return;
    const ZsbBlockWork &W = work[bi];
    const zsb_block b = blocks[bi];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t P0 = (uint32_t)(fo->dst_off & 15) + (uint32_t)W.out_off;     // entry index of the block's first byte
    uint32_t *E = ent + lf.e_off + P0;
    if (b.type == ZSB_BT_RAW) {                                                   // block.rs:76
        const uint8_t *sp = src + b.src_off;
        for (uint32_t i = tid; i < b.size; i += LINK_THREADS) E[i] = LINK_RES | __ldg(sp + i);
        return;
    }
    if (b.type == ZSB_BT_RLE) {                                                   // block.rs:77-79
        const uint32_t v = LINK_RES | src[b.src_off];
        for (uint32_t i = tid; i < b.size; i += LINK_THREADS) E[i] = v;
        return;
    }
    if (b.type != ZSB_BT_COMPRESSED) return;
    const uint32_t nseq = W.nseq, regen = W.lit_regen;
    const bool lit_rle = W.lit_type == ZSB_LT_RLE;
    const uint32_t rle = lit_rle ? src[W.lit_src] : 0u;
    const uint8_t *lp = W.lit_type == ZSB_LT_RAW ? src + W.lit_src : lit_pool + W.lit_buf;
    auto lit = [&](uint32_t i) -> uint32_t { return lit_rle ? rle : (uint32_t)__ldg(lp + i); };
    if (nseq == 0) {                                                              // literals only (RFC 8878; the reference rejects it: Q1)
        for (uint32_t i = tid; i < regen; i += LINK_THREADS) E[i] = LINK_RES | lit(i);
        return;
    }
    const uint64_t *seqs = seq_pool + W.seq_buf;
    {   // the literals behind the last sequence (decoding_context.rs:101-103)
        const uint64_t lastrec = __ldg(seqs + nseq - 1);
        const uint32_t oe = (uint32_t)lastrec & ZSB_REC_POS_MASK, le = (uint32_t)(lastrec >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK;
        for (uint32_t i = tid; i < regen - le; i += LINK_THREADS) E[oe + i] = LINK_RES | lit(le + i);
    }
    const uint32_t r0 = W.rep_in[0], r1 = W.rep_in[1], r2 = W.rep_in[2];
    const uint32_t nbatch = (nseq + 31) / 32;
    // A batch of 32 sequences is one warp's: lane = sequence for the records, then lane = output byte over the bytes the batch regenerates.  On text
    // that is ~270 bytes per batch; on repetitive data a batch may regenerate the whole block with a handful of long matches (one warp, 4 096 rounds,
    // 1.2 ms: the C4 corpus).  Batches of more than LINK_WIDE bytes are therefore put aside and taken by all warps together afterwards, every warp
    // reading the batch's records again and walking every eighth 32-byte group.
    __shared__ uint32_t s_wide[LINK_WIDE_MAX], s_nwide;
    if (tid == 0) s_nwide = 0;
    __syncthreads();
    auto batch = [&](uint32_t bt, uint32_t first, uint32_t stride, bool may_defer) {
        // lane = sequence: where it starts and what it copies
        const uint32_t s = bt * 32 + lane;
        const bool valid = s < nseq;
        const uint64_t rec = valid ? __ldg(seqs + s) : 0ull;
        uint64_t prev = __shfl_up_sync(FULL, rec, 1);
        if (lane == 0) prev = bt ? __ldg(seqs + s - 1) : 0ull;
        const uint32_t out_start = (uint32_t)prev & ZSB_REC_POS_MASK, lit_start = (uint32_t)(prev >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK;
        const uint32_t out_end = valid ? (uint32_t)rec & ZSB_REC_POS_MASK : 0xFFFFFFFFu;
        const uint32_t ll = valid ? ((uint32_t)(rec >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK) - lit_start : 0u;
        const uint32_t rep[3] = {r0, r1, r2};
        uint32_t off = valid ? seq_real_offset((uint32_t)(rec >> (2 * ZSB_REC_POS_BITS)), rep) : 1u;
        const uint32_t dstm = out_start + ll;
        if (valid && (off == 0 || (uint64_t)off > W.out_off + dstm)) { if (first == 0) link_fail(fo, ZSB_E_IMPOSSIBLE_VALUE); off = 0; }   // decoding_context.rs:86-90
        // lane = output byte: the batch regenerates [lo, hi); a byte finds its sequence by bisection over the lanes' end positions
        const uint32_t lo = __shfl_sync(FULL, out_start, 0);
        const uint32_t nval = min(32u, nseq - bt * 32);
        const uint32_t hi = __shfl_sync(FULL, out_end, nval - 1);
        if (may_defer && hi - lo > LINK_WIDE) {
            uint32_t slot = LINK_WIDE_MAX;
            if (lane == 0) slot = atomicAdd(&s_nwide, 1u);
            slot = __shfl_sync(FULL, slot, 0);
            if (slot < LINK_WIDE_MAX) { if (lane == 0) s_wide[slot] = bt; return; }          // (more than the list holds: this warp does it alone)
        }
        for (uint32_t p0 = lo + first; p0 < hi; p0 += stride) {
            const uint32_t p = p0 + lane;
            uint32_t o = 0;
#pragma unroll
            for (uint32_t st = 16; st; st >>= 1) { const uint32_t t = __shfl_sync(FULL, out_end, o + st - 1); if (t <= p) o += st; }
            o = min(o, 31u);
            const uint32_t o_start = __shfl_sync(FULL, out_start, o), o_lit = __shfl_sync(FULL, lit_start, o);
            const uint32_t o_dstm = __shfl_sync(FULL, dstm, o), o_off = __shfl_sync(FULL, off, o);
            if (p < hi) {
                uint32_t e;
                if (p < o_dstm) e = LINK_RES | lit(o_lit + (p - o_start));                       // decoding_context.rs:92-93
                else if (o_off == 0) e = LINK_RES;                                                // the frame has failed: nothing points anywhere
                else { uint32_t k = p - o_dstm; if (k >= o_off) k %= o_off; e = P0 + o_dstm - o_off + k; }   // :95-98, periodic when the match overlaps itself
                E[p] = e;
            }
        }
    };
    for (uint32_t bt = warp; bt < nbatch; bt += LINK_THREADS / 32) batch(bt, 0, 32, true);
    __syncthreads();
    const uint32_t nwide = min(s_nwide, (uint32_t)LINK_WIDE_MAX);
    for (uint32_t i = 0; i < nwide; i++) batch(s_wide[i], 32 * warp, 32 * (LINK_THREADS / 32), false);
}

// tickets: one zeroed word per listed frame.  Grid: a few CTAs per SM, every CTA goes through the frames in order.
__global__ void __launch_bounds__(LINK_THREADS) k_link_resolve(ZsbFrameOut *fout, const ZsbLinkFrame *__restrict__ link_frames, uint32_t n,
                                                               const ZsbCounters *__restrict__ cnt, uint32_t *ent, uint32_t *tickets, uint8_t *dst) {
    __shared__ uint32_t s_c;
    if (cnt->overflow) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t li = 0; li < n; li++) {
        const ZsbLinkFrame lf = link_frames[li];
        ZsbFrameOut *fo = fout + lf.frame;
        const int st0 = *(volatile int *)&fo->status;
        const uint64_t dst_off = fo->dst_off, dst_len = *(volatile uint64_t *)&fo->dst_len;
        if (st0 != ZSB_OK) {                                                      // failed before, or in k_link_init: it regenerates nothing
            if (blockIdx.x == 0 && tid == 0) fo->dst_len = 0;
            continue;
        }
        const uint32_t shift = (uint32_t)(dst_off & 15);
        const uint32_t jlo = shift, jhi = shift + (uint32_t)dst_len;              // the frame's entries
        uint32_t *E = ent + lf.e_off;
        uint8_t *dbase = dst + (dst_off - shift);                                 // 16-byte aligned; entry j <-> dbase[j]
        for (;;) {
            __syncthreads();
            if (tid == 0) s_c = atomicAdd(&tickets[li], 1u);
            __syncthreads();
            const uint64_t c0 = (uint64_t)s_c * LINK_CHUNK;
            if (c0 >= jhi) break;
            // a warp takes 512 consecutive entries: lane l the four entries 128 r + 4 l .. + 3 of each quarter r (16-byte loads, one 32-bit
            // word of output per quarter: 128 contiguous bytes per warp instruction both ways)
            const uint32_t wbase = (uint32_t)c0 + warp * 512u;
            if (wbase >= jhi) continue;
            uint32_t e[16];
            uint32_t um = 0;                                                      // entries still unresolved
            bool corrupt = false;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t j = wbase + 128u * r + 4u * lane;
                uint4 v = make_uint4(LINK_RES, LINK_RES, LINK_RES, LINK_RES);
                if (j < jhi && j + 4 > jlo) v = __ldcg(reinterpret_cast<const uint4 *>(E + j));
                e[4 * r] = v.x; e[4 * r + 1] = v.y; e[4 * r + 2] = v.z; e[4 * r + 3] = v.w;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t jj = j + q;
                    if (jj < jlo || jj >= jhi) e[4 * r + q] = LINK_RES;           // not this frame's bytes (never written, never stored)
                    else if (!(e[4 * r + q] & LINK_RES)) {
                        if (e[4 * r + q] >= jj || e[4 * r + q] < jlo) { e[4 * r + q] = LINK_RES; corrupt = true; }
                        else um |= 1u << (4 * r + q);
                    }
                }
            }
            const uint32_t um0 = um;
            while (um) {
                uint32_t t[16];
#pragma unroll
                for (int q = 0; q < 16; q++) if (um >> q & 1u) t[q] = __ldcg(E + e[q]);
#pragma unroll
                for (int q = 0; q < 16; q++)
                    if (um >> q & 1u) {
                        if (t[q] & LINK_RES) um &= ~(1u << q);
                        else if (t[q] >= e[q] || t[q] < jlo) { t[q] = LINK_RES; um &= ~(1u << q); corrupt = true; }   // a chain only ever goes back: anything else is a bug, not a hang
                        e[q] = t[q];
                    }
            }
            if (corrupt) link_fail(fo, ZSB_E_CORRUPT);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t j = wbase + 128u * r + 4u * lane;
                if (j >= jhi || j + 4 <= jlo) continue;
                // what was resolved here shortens every chain that comes through it
                if (um0 >> (4 * r) & 15u) __stcg(reinterpret_cast<uint4 *>(E + j), make_uint4(e[4 * r], e[4 * r + 1], e[4 * r + 2], e[4 * r + 3]));
                const uint32_t w = (e[4 * r] & 255u) | (e[4 * r + 1] & 255u) << 8 | (e[4 * r + 2] & 255u) << 16 | (e[4 * r + 3] & 255u) << 24;
                if (j >= jlo && j + 4 <= jhi) *reinterpret_cast<uint32_t *>(dbase + j) = w;
                else {
#pragma unroll
                    for (int q = 0; q < 4; q++) if (j + q >= jlo && j + q < jhi) dbase[j + q] = (uint8_t)(w >> (8 * q));
                }
            }
        }
    }
}

// XXH64 of the frames k_link_resolve wrote.  A frame is four dependent accumulator chains over all of its stripes (33.5 M rounds for 1 GiB)
// whatever the number of threads, so the frame gets ONE hashing warp with an SM to itself and everything else is kept off that warp: measured,
// a single warp that also computes x * P2 is bound by its scheduler's multiplier pipe (nine IMAD-class instructions per round, 33 cycles per
// round), not by the chain.  So the CTA is a three-stage pipeline of specialised warps:
//   bulk copies (cp.async.bulk, one mbarrier per tile, issued by a helper thread a ring ahead) bring 4 KiB tiles of the frame into shared memory,
//   warps 1-3 (other schedulers) turn every 8-byte word x into a = x * P2 -- at any alignment of the frame -- into a second ring,
//   warp 0 reads a with LDS.64 one group of eight stripes ahead and runs nothing but the chain (XAcc::round_a); lanes 0-3 own the accumulators.
// Hand-over by mbarriers both ways (tile filled / products ready / products consumed).  Tiles lie wholly inside the frame (16-byte aligned
// from the frame's aligned base, 16 bytes of overhang for frames that start unaligned); the last < 2 tiles and the tail are read from HBM.
__device__ __forceinline__ uint64_t xxh_finish(const uint8_t *p, uint64_t len, uint64_t nstripes, uint64_t v1, uint64_t v2, uint64_t v3, uint64_t v4) {
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
    return h;
}
#define XO_TILE 4096u
#define XO_STAGES 6u                 // raw tiles in flight
#define XO_ASTAGES 4u                // tiles of products between the helpers and the hashing warp
#define XO_STRIDE (XO_TILE + 16u)
#define XO_HELPERS 96u
__device__ __forceinline__ void zsb_mbar_arrive(uint32_t mbar_sa) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar_sa) : "memory"); }
__device__ __forceinline__ void xo_load(uint64_t (&X)[8], const uint8_t *tp) {
#pragma unroll
    for (int j = 0; j < 8; j++) X[j] = *reinterpret_cast<const unsigned long long *>(tp + 32 * j);
}
__global__ void __launch_bounds__(32 + XO_HELPERS) k_xxh_one(const uint8_t *__restrict__ dst, ZsbFrameOut *fout, const zsb_frame *__restrict__ frames,
                                                             const ZsbLinkFrame *__restrict__ link_frames, const ZsbCounters *__restrict__ cnt) {
    __shared__ __align__(128) uint8_t s_raw[XO_STAGES * XO_STRIDE];
    __shared__ __align__(128) uint8_t s_a[XO_ASTAGES * XO_TILE];
    __shared__ __align__(8) unsigned long long s_bar[XO_STAGES + 2 * XO_ASTAGES];      // raw full | products full | products empty
    if (cnt->overflow) return;
    const uint32_t f = link_frames[blockIdx.x].frame;
    if (!frames[f].has_checksum || fout[f].status != ZSB_OK) return;
    const uint64_t len = fout[f].dst_len;
    const uint8_t *p = dst + fout[f].dst_off;
    const uint32_t tid = threadIdx.x;
    const uint32_t shift = (uint32_t)((uintptr_t)p & 15);
    const uint8_t *base = p - shift;
    const uint64_t ntiles = shift + len >= 16 + XO_TILE ? (shift + len - 16) / XO_TILE : 0;       // tile t reads base + t * XO_TILE .. + XO_TILE + 16
    const uint32_t raw_sa = (uint32_t)__cvta_generic_to_shared(s_raw), bar_sa = (uint32_t)__cvta_generic_to_shared(s_bar);
    const uint32_t afull_sa = bar_sa + 8 * XO_STAGES, aempty_sa = afull_sa + 8 * XO_ASTAGES;
    if (tid == 0) {
        for (uint32_t i = 0; i < XO_STAGES; i++) zsb_mbar_init(bar_sa + 8 * i, 1);
        for (uint32_t i = 0; i < XO_ASTAGES; i++) { zsb_mbar_init(afull_sa + 8 * i, XO_HELPERS); zsb_mbar_init(aempty_sa + 8 * i, 1); }
    }
    __syncthreads();
    if (tid >= 32) {
        // ---- helpers: bulk copies and the products
        const uint32_t hid = tid - 32;
        if (hid == 0)
            for (uint32_t i = 0; i < XO_STAGES && i < ntiles; i++) {
                zsb_mbar_expect(bar_sa + 8 * i, XO_STRIDE);
                zsb_bulk_g2s(raw_sa + i * XO_STRIDE, base + (uint64_t)i * XO_TILE, XO_STRIDE, bar_sa + 8 * i);
            }
        const uint32_t sh = (shift & 7) * 8;
        for (uint64_t t = 0; t < ntiles; t++) {
            const uint32_t st = (uint32_t)(t % XO_STAGES), sa = (uint32_t)(t % XO_ASTAGES);
            zsb_mbar_wait(bar_sa + 8 * st, (uint32_t)(t / XO_STAGES) & 1u);                       // the tile has arrived
            zsb_mbar_wait(aempty_sa + 8 * sa, ((uint32_t)(t / XO_ASTAGES) & 1u) ^ 1u);            // the products of tile t - XO_ASTAGES were consumed
            const unsigned long long *rp = reinterpret_cast<const unsigned long long *>(s_raw + st * XO_STRIDE + (shift & 8u));
            unsigned long long *ap = reinterpret_cast<unsigned long long *>(s_a + sa * XO_TILE);
            for (uint32_t w = hid; w < XO_TILE / 8; w += XO_HELPERS) {
                const uint64_t x = sh == 0 ? rp[w] : (rp[w] >> sh) | (rp[w + 1] << (64 - sh));
                ap[w] = x * XP2;
            }
            zsb_mbar_arrive(afull_sa + 8 * sa);                                                    // (release: the products are visible to who waits)
            asm volatile("bar.sync 1, %0;" ::"n"(XO_HELPERS) : "memory");                          // every helper has read the raw tile: it may be filled again
            if (hid == 0 && t + XO_STAGES < ntiles) {
                zsb_mbar_expect(bar_sa + 8 * st, XO_STRIDE);
                zsb_bulk_g2s(raw_sa + st * XO_STRIDE, base + (t + XO_STAGES) * XO_TILE, XO_STRIDE, bar_sa + 8 * st);
            }
        }
        return;
    }
    // ---- warp 0: the chain
    const uint32_t lane = tid, q = lane & 3;
    XAcc va; va.init(q == 0 ? XP1 + XP2 : q == 1 ? XP2 : q == 2 ? 0ull : 0ull - XP1);
    const uint64_t nstripes = len >> 5;
    for (uint64_t t = 0; t < ntiles; t++) {
        const uint32_t sa = (uint32_t)(t % XO_ASTAGES);
        zsb_mbar_wait(afull_sa + 8 * sa, (uint32_t)(t / XO_ASTAGES) & 1u);
        const uint8_t *tp = s_a + sa * XO_TILE + 8 * q;
        uint64_t X[8], Y[8];
        xo_load(X, tp);
#pragma unroll 1
        for (uint32_t g = 0; g < XO_TILE / 256; g += 2) {
            xo_load(Y, tp + 256 * (g + 1));
#pragma unroll
            for (int j = 0; j < 8; j++) va.round_a(X[j]);
            if (g + 2 < XO_TILE / 256) xo_load(X, tp + 256 * (g + 2));
#pragma unroll
            for (int j = 0; j < 8; j++) va.round_a(Y[j]);
        }
        __syncwarp();                                                     // every lane has read the products
        if (lane == 0) zsb_mbar_arrive(aempty_sa + 8 * sa);
    }
    for (uint64_t cur = ntiles * (XO_TILE / 32); cur < nstripes; cur++) va.round(ld64_any(p + (cur << 5) + 8 * q));
    const uint64_t v = va.acc();
    const uint64_t v1 = __shfl_sync(FULL, v, 0), v2 = __shfl_sync(FULL, v, 1), v3 = __shfl_sync(FULL, v, 2), v4 = __shfl_sync(FULL, v, 3);
    if (lane == 0) fout[f].xxh64 = xxh_finish(p, len, nstripes, v1, v2, v3, v4);
}

// ======================================================================================= k_xxh
// XXH64 has no combine step: a frame is four dependent accumulator chains over its 32-byte stripes
// (round = rotl(acc + x*P2, 31) * P1, ~25 cycles), so a frame is four lanes and the kernel is bound by that
// chain as long as the bytes arrive in time.  One warp hashes 8 frames (4 lanes each).  The frames' bytes are
// staged through shared memory with cp.async, 512 bytes per frame and step, double buffered: coalesced 16-byte
// requests, no registers held across the latency, and the lanes read their 8 bytes per stripe from shared memory
// at any alignment (two aligned 64-bit loads and a funnel).
#define XXH_FRAMES 8                 // frames per warp
#define XXH_CHUNK 512u               // bytes of one frame staged per step (16 stripes)
#define XXH_BUF 544u                 // chunk + 16 bytes of overhang for unaligned frames, padded to spread the banks
#define XXH_WARPS 4
__global__ void __launch_bounds__(32 * XXH_WARPS) k_xxh(const uint8_t *__restrict__ dst, ZsbFrameOut *fout, const uint32_t *__restrict__ list,
                                                        uint32_t n, const ZsbCounters *__restrict__ cnt) {
    __shared__ __align__(16) uint8_t s_buf[XXH_WARPS][2][XXH_FRAMES][XXH_BUF];
    __shared__ unsigned long long s_pal[XXH_WARPS][XXH_FRAMES];     // 16-byte aligned base of each frame
    __shared__ uint32_t s_tot[XXH_WARPS][XXH_FRAMES];               // bytes to stage from that base (stripes only)
    if (cnt->overflow) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = lane & 3, fj = lane >> 2;
    const uint32_t gq = (blockIdx.x * XXH_WARPS + warp) * XXH_FRAMES + fj;
    const bool active = gq < n;
    const uint32_t f = active ? list[gq] : 0;
    const bool ok = active && fout[f].status == ZSB_OK;
    const uint8_t *p = dst + (ok ? fout[f].dst_off : 0);
    const uint64_t len = ok ? fout[f].dst_len : 0;
    const uint64_t nstripes = len >> 5;
    const uint32_t m = (uint32_t)(uintptr_t)p & 15u;
    // a frame longer than 4 GiB would overflow the 32-bit staging offsets: hash it in 2 GiB sections
    XAcc va; va.init(q == 0 ? XP1 + XP2 : q == 1 ? XP2 : q == 2 ? 0ull : 0ull - XP1);
    const uint32_t zero = blockIdx.y;                               // (the grid is one-dimensional: see XAcc::round)
    const uint64_t SECT = 1ull << 26;                               // stripes per section (2 GiB)
    // Frames that start 8-byte aligned (all of them when the sizes are multiples of 8) are read straight from HBM/L2: the four
    // lanes of a frame take the four words of a stripe (one full 32-byte sector per request), 32 stripes are requested while the
    // previous 32 are hashed (the chain of 32 rounds outlasts the DRAM latency), and nothing but the load, x * P2 and the round
    // itself is executed per stripe: 12 instructions instead of the 23 of the staged path below, which remains for the rest.
    const bool direct = __all_sync(FULL, !ok || ((uintptr_t)p & 7u) == 0);
    if (direct) {
        constexpr int NS = 32;
        const unsigned long long *g = reinterpret_cast<const unsigned long long *>(p) + q;
        uint64_t maxst = nstripes;
#pragma unroll
        for (int d = 16; d; d >>= 1) { const uint64_t o = __shfl_xor_sync(FULL, maxst, d); maxst = o > maxst ? o : maxst; }
        // three buffers of 32 words in rotation: two steps (64 requests per lane) are in flight while one is hashed
        uint64_t A[NS], B[NS], C[NS];
        auto fetch = [&](uint64_t (&X)[NS], uint64_t st) {
            const unsigned long long *gs = g + 4 * st;
            if (st + NS <= nstripes) {
#pragma unroll
                for (int s = 0; s < NS; s++) X[s] = __ldcg(gs + 4 * s);
            } else {
#pragma unroll
                for (int s = 0; s < NS; s++) X[s] = st + s < nstripes ? __ldcg(gs + 4 * s) : 0ull;
            }
        };
        auto hash = [&](const uint64_t (&X)[NS], uint64_t st) {
            if (st + NS <= nstripes) {
#pragma unroll
                for (int s = 0; s < NS; s++) va.round(X[s], zero);
            } else {
#pragma unroll
                for (int s = 0; s < NS; s++) if (st + s < nstripes) va.round(X[s], zero);
            }
        };
        fetch(A, 0); fetch(B, NS);
        for (uint64_t s0 = 0; s0 < maxst; s0 += 3 * NS) {
            fetch(C, s0 + 2 * NS); hash(A, s0);
            fetch(A, s0 + 3 * NS); hash(B, s0 + NS);
            fetch(B, s0 + 4 * NS); hash(C, s0 + 2 * NS);
        }
    }
    for (uint64_t s0 = 0; !direct && __any_sync(FULL, s0 < nstripes); s0 += SECT) {
        const uint64_t left = s0 < nstripes ? nstripes - s0 : 0;
        const uint32_t ns = (uint32_t)(left < SECT ? left : SECT);  // stripes of this section for this frame
        const uint8_t *ps = p + (s0 << 5);
        __syncwarp();
        if (q == 0) { s_pal[warp][fj] = (unsigned long long)(uintptr_t)(ps - m); s_tot[warp][fj] = ns ? m + (ns << 5) : 0u; }
        __syncwarp();
        uint32_t nch = (ns + 15) >> 4;                              // steps this frame needs
#pragma unroll
        for (int d = 16; d; d >>= 1) nch = max(nch, __shfl_xor_sync(FULL, nch, d));
        const uint32_t sh = (m & 7u) * 8u;
        // stage step c of all 8 frames into buffer b: lane l copies 16-byte unit l of every frame, lane 0 also the overhang unit
        auto stage = [&](uint32_t c, uint32_t b) {
#pragma unroll
            for (int j = 0; j < XXH_FRAMES; j++) {
                const uint8_t *g = reinterpret_cast<const uint8_t *>((uintptr_t)s_pal[warp][j]);
                const uint32_t tot = s_tot[warp][j];
                const uint32_t A = c * XXH_CHUNK + 16u * lane;
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&s_buf[warp][b][j][16u * lane]);
                const uint32_t sz = A < tot ? 16u : 0u;             // 0: nothing is read from global, the unit is zero filled
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(g + (sz ? A : 0u)), "r"(sz) : "memory");
                if (lane == 0) {
                    const uint32_t A2 = c * XXH_CHUNK + XXH_CHUNK, sz2 = A2 < tot ? 16u : 0u;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sa + XXH_CHUNK), "l"(g + (sz2 ? A2 : 0u)), "r"(sz2) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (nch) stage(0, 0);
        for (uint32_t c = 0; c < nch; c++) {
            __syncwarp();                                           // buffer (c+1)&1 was fully read in step c-1
            if (c + 1 < nch) { stage(c + 1, (c + 1) & 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const uint8_t *bp = &s_buf[warp][c & 1][fj][0];
            const uint32_t s_lo = c << 4;
            const unsigned long long *w = reinterpret_cast<const unsigned long long *>(bp + ((m + 8u * q) & ~7u));
            if (s_lo + 16 <= ns) {
                // a full step: the 16 (or 17) words are loaded first, x * P2 is off the chain, the chain is rotl(acc + y, 31) * P1
                uint64_t y[16];
                if (sh == 0) {
#pragma unroll
                    for (int s = 0; s < 16; s++) y[s] = (w[4 * s] * XP2) ^ zero;
                } else {
#pragma unroll
                    for (int s = 0; s < 16; s++) y[s] = ((zsb_shr64(w[4 * s], sh) | zsb_shl64(w[4 * s + 1], 64 - sh)) * XP2) ^ zero;
                }
#pragma unroll
                for (int s = 0; s < 16; s++) va.round_a(y[s]);
            } else {
                for (uint32_t s = 0; s < 16 && s_lo + s < ns; s++) {
                    va.round(zsb_shr64(w[4 * s], sh) | zsb_shl64(w[4 * s + 1], 64 - sh));
                }
            }
        }
    }
    __syncwarp();
    const uint32_t qb = lane & ~3u;
    const uint64_t v = va.acc();
    const uint64_t v1 = __shfl_sync(FULL, v, qb), v2 = __shfl_sync(FULL, v, qb + 1), v3 = __shfl_sync(FULL, v, qb + 2), v4 = __shfl_sync(FULL, v, qb + 3);
    if (ok && q == 0) fout[f].xxh64 = xxh_finish(p, len, nstripes, v1, v2, v3, v4);
}

// ======================================================================================= stage kernels (one lane)
__global__ void __launch_bounds__(32) k_stage_fse(const uint8_t *desc, uint32_t n, int max_sym, const int16_t *dist_in, int ndist_in, int al_in,
                                                  int *res /* rc, al, nsym, consumed */, uint32_t *cells, int16_t *dist_out) {
    __shared__ FseWarpScratch X;
    const uint32_t lane = threadIdx.x;
    int al = al_in, nsym = ndist_in, rc = 0; uint32_t consumed = 0;
    if (desc) {
        if (lane == 0) {
            FwdBits f; fwd_init(f, desc, n);
            rc = fse_read_ncount(f, X.cnt, 1, max_sym < 64 ? max_sym : 64, al, nsym);
            consumed = fwd_bytes_read(f);
        }
        rc = __shfl_sync(FULL, rc, 0); al = __shfl_sync(FULL, al, 0); nsym = __shfl_sync(FULL, nsym, 0); consumed = __shfl_sync(FULL, consumed, 0);
    } else for (int i = (int)lane; i < nsym; i += 32) X.cnt[i] = dist_in[i];
    __syncwarp();
    if (!rc && dist_out) for (int i = (int)lane; i < nsym; i += 32) dist_out[i] = X.cnt[i];
    if (!rc) { if (al > ZSB_MAX_AL || al < 5) rc = al > ZSB_MAX_AL ? ZSB_E_LARGE_ACCURACY_LOG : ZSB_E_ARG; else rc = fse_build_table_warp(X, nsym, al, cells, 1, 3, nullptr, lane); }
    if (lane == 0) { res[0] = rc; res[1] = al; res[2] = nsym; res[3] = (int)consumed; }
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
// Small results straight into page-locked host memory (the kernel writes over the bus itself): a cudaMemcpyAsync of a few KB would
// queue behind the megabytes the copy engines are moving for other shards, and the host would learn a shard's sizes only then.
__global__ void __launch_bounds__(256) k_publish(unsigned long long *host, const unsigned long long *a, uint32_t na, const unsigned long long *b, uint32_t nb, uint32_t b_at) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += gridDim.x * blockDim.x) {
        if (i < na) host[i] = a[i]; else host[b_at + (i - na)] = b[i - na];
    }
}
void zsbk_publish(cudaStream_t st, void *host_dev_ptr, const ZsbCounters *cnt, const ZsbFrameOut *fout, uint32_t nf) {
    static_assert(sizeof(ZsbCounters) % 8 == 0 && sizeof(ZsbCounters) <= 64 && sizeof(ZsbFrameOut) % 8 == 0, "published as 8-byte words");
    const uint32_t na = sizeof(ZsbCounters) / 8, nb = nf * (uint32_t)(sizeof(ZsbFrameOut) / 8);
    k_publish<<<(na + nb + 255) / 256 < 64 ? (na + nb + 255) / 256 : 64, 256, 0, st>>>((unsigned long long *)host_dev_ptr, (const unsigned long long *)cnt, na,
                                                                                   (const unsigned long long *)fout, nb, 8);
}

cudaError_t zsbk_init() {
    cudaError_t e = set_smem((const void *)k_seq_slow, SEQ_SLOW_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = set_smem((const void *)k_huf, HUF_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = set_smem((const void *)k_seq_t<SEQ_HELPERS, SEQ_WIN, 2, false>, SEQ_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = set_smem((const void *)k_seq_t<SEQX_WARPS, SEQX_WIN, SEQX_NBUF, true>, SEQX_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = set_smem((const void *)k_exec<512, false>, EXEC_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return set_smem((const void *)k_exec<1024, true>, EXEC_SMEM_BYTES);
}
void zsbk_parse(cudaStream_t st, const uint8_t *src, const zsb_block *blocks, ZsbBlockWork *work, uint32_t nb, uint32_t flags) {
    if (nb) k_parse<<<(nb + 127) / 128, 128, 0, st>>>(src, blocks, work, nb, flags);
}
void zsbk_plan1(cudaStream_t st, const zsb_frame *frames, uint32_t nf, const zsb_block *blocks, uint32_t nb, ZsbBlockWork *work,
                ZsbFrameOut *fout, uint32_t *huf_list, uint32_t *seq_list, ZsbCounters *cnt, uint64_t lit_cap, uint64_t seq_cap, uint32_t flags) {
    k_plan1<<<nf > 1024 ? (nf + 1023) / 1024 : 1, 1024, 0, st>>>(frames, nf, blocks, nb, work, fout, huf_list, seq_list, cnt, lit_cap, seq_cap, flags);
}
void zsbk_huf(cudaStream_t st, uint32_t ncomp, const uint8_t *src, uint64_t src_len, ZsbBlockWork *work, const uint32_t *huf_list,
              ZsbCounters *cnt, uint8_t *lit_pool, uint64_t lit_cap, uint64_t over_cap, uint32_t flags) {
    if (ncomp) k_huf<<<(ncomp + HUF_SLOTS - 1) / HUF_SLOTS, HUF_THREADS, HUF_SMEM_BYTES, st>>>(src, src_len, work, huf_list, cnt, lit_pool, lit_cap, over_cap, flags);
}
void zsbk_seq(cudaStream_t st, uint32_t ncomp, const uint8_t *src, ZsbBlockWork *work, const uint32_t *seq_list, ZsbCounters *cnt,
              uint64_t *seq_pool, uint32_t *slow_list, bool shared_device, uint32_t chains_hint, const zsb_frame *frames, const zsb_block *blocks,
              const uint64_t *pre_off, const uint8_t *lit_pool, uint8_t *dst) {
    if (!ncomp) return;
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148; }
    // one CTA per SM (shared memory): as few waves as possible, and the blocks spread evenly over the CTAs of those waves
    const uint32_t waves = (ncomp + (uint32_t)n_sm * SEQ_CHAINS - 1) / ((uint32_t)n_sm * SEQ_CHAINS);
    uint32_t used = (ncomp + waves * (uint32_t)n_sm - 1) / (waves * (uint32_t)n_sm);
    static const uint32_t sub_chains = getenv("ZSB_SUB_CHAINS") ? (uint32_t)atoi(getenv("ZSB_SUB_CHAINS")) : SEQ_CHAINS;     // experiment knob
    if (used > SEQ_CHAINS) used = SEQ_CHAINS;
    if (shared_device) used = sub_chains;                          // (batches of other streams run beside this one: few SMs each)
    if (chains_hint && chains_hint <= SEQ_CHAINS) used = chains_hint;      // low-latency shards of the pipelined host path
    if (used < 1) used = 1;
    SeqxArgs A; A.frames = frames; A.blocks = blocks; A.pre_off = pre_off; A.lit_pool = lit_pool; A.dst = dst;
    if (pre_off) {
        // k_seqx: one consumer warp per chain, at most SEQX_WARPS chains per CTA
        const uint32_t wx = (ncomp + (uint32_t)n_sm * SEQX_WARPS - 1) / ((uint32_t)n_sm * SEQX_WARPS);
        uint32_t ux = (ncomp + wx * (uint32_t)n_sm - 1) / (wx * (uint32_t)n_sm);
        if (ux > SEQX_WARPS) ux = SEQX_WARPS;
        if (ux < 1) ux = 1;
        k_seq_t<SEQX_WARPS, SEQX_WIN, SEQX_NBUF, true><<<(ncomp + ux - 1) / ux, 32 * (1 + SEQX_WARPS), SEQX_SMEM_BYTES, st>>>(src, work, seq_list, cnt, seq_pool, slow_list, ux, A);
        return;
    }
    k_seq_t<SEQ_HELPERS, SEQ_WIN, 2, false><<<(ncomp + used - 1) / used, 32 * (1 + SEQ_HELPERS), SEQ_SMEM_BYTES, st>>>(src, work, seq_list, cnt, seq_pool, slow_list, used, A);
}
void zsbk_seq_slow(cudaStream_t st, uint32_t ncomp, const uint8_t *src, uint64_t src_len, ZsbBlockWork *work, const uint32_t *slow_list,
                   const ZsbCounters *cnt, uint64_t *seq_pool) {
    if (ncomp) k_seq_slow<<<(ncomp + 31) / 32, 32, SEQ_SLOW_SMEM_BYTES, st>>>(src, src_len, work, slow_list, cnt, seq_pool);
}
void zsbk_plan2(cudaStream_t st, const zsb_frame *frames, uint32_t nf, const zsb_block *blocks, ZsbBlockWork *work, ZsbFrameOut *fout,
                ZsbCounters *cnt, uint64_t dst_cap, uint32_t flags, const uint64_t *pre_off) {
    k_plan2<<<nf > 1024 ? (nf + 1023) / 1024 : 1, 1024, 0, st>>>(frames, nf, blocks, work, fout, cnt, dst_cap, flags, pre_off);
}
void zsbk_rawrle(cudaStream_t st, uint32_t n, const uint8_t *src, const zsb_block *blocks, const ZsbBlockWork *work, const ZsbFrameOut *fout,
                 const uint32_t *list, const ZsbCounters *cnt, uint8_t *dst) {
    if (n) k_rawrle<<<n, 256, 0, st>>>(src, blocks, work, fout, list, cnt, dst);
}
void zsbk_exec(cudaStream_t st, uint32_t n, const uint8_t *src, const zsb_frame *frames, const zsb_block *blocks, const ZsbBlockWork *work,
               ZsbFrameOut *fout, const uint32_t *exec_list, const ZsbCounters *cnt, const uint64_t *seq_pool, const uint8_t *lit_pool, uint8_t *dst, bool shared_device,
               uint32_t flags, void *wave, uint32_t *blk_done, uint32_t ctas_per_frame) {
    if (!n) return;
    if (wave && !shared_device && ctas_per_frame > 1) {
        k_exec<1024, true><<<n * ctas_per_frame, 1024, EXEC_SMEM_BYTES, st>>>(src, frames, blocks, work, fout, exec_list, cnt, seq_pool, lit_pool, dst, flags,
                                                                              (ZsbWave *)wave, blk_done, ctas_per_frame);
        return;
    }
    // (on its own: 31 executing warps and one that hashes the frame behind them, so a frame of many blocks is not hashed by k_xxh afterwards)
    if (shared_device) k_exec<512, false><<<n, 512, EXEC_SMEM_BYTES, st>>>(src, frames, blocks, work, fout, exec_list, cnt, seq_pool, lit_pool, dst, flags, nullptr, nullptr, 1);
    else k_exec<1024, true><<<n, 1024, EXEC_SMEM_BYTES, st>>>(src, frames, blocks, work, fout, exec_list, cnt, seq_pool, lit_pool, dst, flags, nullptr, nullptr, 1);
}
void zsbk_exec2(cudaStream_t st, uint32_t n, const uint8_t *src, const zsb_frame *frames, const zsb_block *blocks, const ZsbBlockWork *work,
                ZsbFrameOut *fout, const uint32_t *exec_list, const ZsbCounters *cnt, const uint64_t *seq_pool, const uint8_t *lit_pool, uint8_t *dst) {
    if (n) k_exec2<<<(n + EX2_WARPS - 1) / EX2_WARPS, 32 * EX2_WARPS, 0, st>>>(src, frames, blocks, work, fout, exec_list, n, cnt, seq_pool, lit_pool, dst);
}
void zsbk_link(cudaStream_t st, uint32_t n_frames, uint32_t n_blocks, const uint8_t *src, const zsb_block *blocks, const ZsbBlockWork *work, ZsbFrameOut *fout,
               const void *link_blocks, const void *link_frames, const ZsbCounters *cnt, const uint64_t *seq_pool, const uint8_t *lit_pool,
               uint32_t *ent, uint32_t *tickets, uint8_t *dst, int n_sm, int which) {
    if (!n_frames) return;
    if (n_blocks && (which & 1)) k_link_init<<<n_blocks, LINK_THREADS, 0, st>>>(src, blocks, work, fout, (const uint2 *)link_blocks, (const ZsbLinkFrame *)link_frames, cnt, seq_pool, lit_pool, ent);
    if (which & 2) k_link_resolve<<<(n_sm > 0 ? n_sm : 148) * 8, LINK_THREADS, 0, st>>>(fout, (const ZsbLinkFrame *)link_frames, n_frames, cnt, ent, tickets, dst);
}
void zsbk_xxh_one(cudaStream_t st, uint32_t n_frames, const uint8_t *dst, ZsbFrameOut *fout, const zsb_frame *frames, const void *link_frames, const ZsbCounters *cnt) {
    if (n_frames) k_xxh_one<<<n_frames, 32 + XO_HELPERS, 0, st>>>(dst, fout, frames, (const ZsbLinkFrame *)link_frames, cnt);
}
void zsbk_xxh(cudaStream_t st, uint32_t n, const uint8_t *dst, ZsbFrameOut *fout, const uint32_t *list, const ZsbCounters *cnt) {
    if (n) k_xxh<<<(n + XXH_WARPS * XXH_FRAMES - 1) / (XXH_WARPS * XXH_FRAMES), 32 * XXH_WARPS, 0, st>>>(dst, fout, list, n, cnt);
}
void zsbk_stage_fse(cudaStream_t st, const uint8_t *desc, uint32_t n, int max_sym, const int16_t *dist_in, int ndist_in, int al_in,
                    int *res, uint32_t *cells, int16_t *dist_out) {
    k_stage_fse<<<1, 32, 0, st>>>(desc, n, max_sym, dist_in, ndist_in, al_in, res, cells, dist_out);
}
void zsbk_stage_huf(cudaStream_t st, const uint8_t *desc, uint32_t n, int *res, uint8_t *lens, uint16_t *lut) {
    k_stage_huf<<<1, 1, 0, st>>>(desc, n, res, lens, lut);
}
void zsbk_stage_records(cudaStream_t st, const uint32_t *tri, uint32_t nseq, uint32_t nlit, uint64_t *rec, ZsbBlockWork *w) {
    k_stage_records<<<1, 1, 0, st>>>(tri, nseq, nlit, rec, w);
}
