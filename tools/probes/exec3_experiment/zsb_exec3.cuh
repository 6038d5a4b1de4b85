// zsb_exec3.cuh -- sequence execution organised by OUTPUT BYTES (included by zsb_kernels.cu).
//
// == DecodingContext::execute_sequences (decoding_context.rs:78-106): per sequence, `ll` literals are appended, then `ml`
//    bytes are copied one at a time from `offset` bytes back (overlap allowed), and the literals left over follow the last
//    sequence (:101-103).
//
// Round 1's executor (k_exec2) gave every lane one sequence: ~590 warp instructions per 32 sequences, 18 of 32 lanes
// active on average, most of them in byte-wise predicated stores.  Text at level 3 is ~8.5 output bytes per sequence, so
// the unit of work here is the output byte instead: a warp produces 32 consecutive output bytes per step ("chunk"),
//
//   lane j  ->  byte p = P + j  ->  which sequence covers p (popc over a bitmap of the sequence starts that fall into the
//   chunk, one REDUX.OR)  ->  that sequence's record {S start, M match start, dlt, off} (one 16-byte shared-memory load)
//   ->  p < M ? literal[p - dlt] : output[p - off]  ->  one byte store,
//
// every lane active, ~30 instructions per 32 bytes whatever the mix of literal runs, short and long matches (a 100 KiB
// match is simply 3 200 chunks).  A match byte whose source lies inside its own chunk (offset < 32) is resolved by pointer
// jumping over the lanes (<= 5 shuffle rounds, only when such a lane exists).  Every other match byte is read back from
// HBM/L2, where the chunks are stored as they are produced (32 consecutive bytes per store instruction; L2 merges them into
// whole sectors): in text ~90 % of the match sources are further back than any shared-memory window of a few KiB would reach,
// so there is no such window -- only the 128 bytes of the step under construction are kept in shared memory.
//
// The records of a window of sequences are staged in shared memory as 16-byte {S, M, dlt, off}; the stand-alone kernel
// k_exec3 converts them from the packed 64-bit records of the sequence stage, the fused sequence kernel writes them directly.
#pragma once

#define EX3_WIN 128u                  // records staged per window (stand-alone kernel)
#define EX3_WARPS 4
#define EX3_U 4                       // chunks (of 32 bytes) per step: their HBM / literal loads are all in flight together
#define EX3_FAR_BIAS 0x10000000u      // offsets are < 2^27 (ZSB_OFF_MAX) and positions < 2^17: sp + bias is a non-negative 32-bit index

struct __align__(16) Ex3Rec { uint32_t S, M, dlt, off; };   // positions relative to the start of the block; literal index = p - dlt

__device__ __forceinline__ uint32_t ex3_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t ex3_lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void ex3_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// The (at most one) byte a lane requests in part A -- a literal, or an earlier byte of the frame -- as ONE predicated load whose
// register is left undefined when the lane requests nothing.  (Written in C++ as "v = dflt; if (p) v = load", or as two predicated
// loads into one register, the compiler puts a move behind each load; the move waits for the load, and the loads of a step no
// longer overlap: measured 2.6 ms instead of the figures in DESIGN.md.)  A plain ld.global: earlier bytes of the frame were stored
// by this warp, and L1 is coherent for the stores of its own SM.
__device__ __forceinline__ uint32_t ex3_request(const uint8_t *a, bool p) {
    uint32_t v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u8 %0, [%1];\n\t}" : "=r"(v) : "l"(a), "r"((uint32_t)p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ex3_lds128(uint32_t a) {
    uint4 r; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory"); return r;
}

// ---- the fast form: shared-memory ring for near sources, start bitmap instead of a search, loads one step ahead --------------
//
// A first form of this executor stored every chunk straight to HBM and read every match source back from there (no ring): 2.6 ms on
// C2 (4 096 x 128 KiB text frames) against 1.48 ms for k_exec2, because every step waited for its own HBM loads (~1 000 cycles, and
// nothing of the next step could be requested before the step was stored: a source may lie in it) and cost ~60 instructions per
// chunk.  Here
//   * the last EX3_RING bytes of the frame live in a shared-memory ring (index = global address & mask, so that 16-byte units of
//     the ring and of HBM coincide); sources inside it are read there at the moment the chunk is assembled, so everything OLDER can
//     be requested a whole step ahead: part A of step s+1 runs before part B of step s;
//   * which record covers a byte comes from a bitmap of record starts built when the records are staged (bit d of word u: a
//     record starts at byte Q0 + 32 u + d + 1): one broadcast 16-byte load per step instead of a REDUX per chunk;
//   * whole 16-byte units leave the ring for HBM every 512 bytes.
#define EX3_RING 1024u                // 1 KiB aligned: a ring address is ring | (global address & mask), one LOP3
#define EX3_MASK (EX3_RING - 1u)
#define EX3_STEP (32 * EX3_U)         // 128 bytes
#define EX3_SUB 2048                  // bytes one start bitmap covers (EX3_SUB / 32 words + 4 of slack, 16-byte aligned)
#define EX3_SBITS_WORDS (EX3_SUB / 32 + 4)

struct Ex3Ring {
    uint32_t ring_sa;       // shared-memory address of the frame's ring
    uint8_t *gblk;          // global address of the block's first byte; positions are relative to it
    uint32_t g0;            // low 32 bits of gblk
    int32_t flushed;        // [.., flushed) is in HBM; everything from there up to the position being produced is in the ring
    int32_t ring_floor;     // the ring holds nothing below this position (block start after a raw / RLE block, frame start)
};
__device__ __forceinline__ void ex3_flush_units(Ex3Ring &X, int32_t upto, uint32_t lane) {
    const int32_t hi = upto - (int32_t)((X.g0 + (uint32_t)upto) & 15u);
    if (hi <= X.flushed) return;
    int32_t a = X.flushed + (int32_t)((16u - ((X.g0 + (uint32_t)X.flushed) & 15u)) & 15u);      // first 16-byte boundary
    if (a > hi) a = hi;
    for (int32_t p = X.flushed + (int32_t)lane; p < a; p += 32) X.gblk[p] = (uint8_t)ex3_lds8(X.ring_sa | ((X.g0 + (uint32_t)p) & EX3_MASK));   // only at a frame start / behind a raw or RLE block
    for (int32_t p = a + 16 * (int32_t)lane; p < hi; p += 512)
        *reinterpret_cast<uint4 *>(X.gblk + p) = ex3_lds128(X.ring_sa | ((X.g0 + (uint32_t)p) & EX3_MASK));
    X.flushed = hi;
}
// everything up to `upto`, byte-wise tail included (end of a block that a raw / RLE block or the end of the frame follows)
__device__ __forceinline__ void ex3_flush_all(Ex3Ring &X, int32_t upto, uint32_t lane) {
    ex3_flush_units(X, upto, lane);
    for (int32_t p = X.flushed + (int32_t)lane; p < upto; p += 32) X.gblk[p] = (uint8_t)ex3_lds8(X.ring_sa | ((X.g0 + (uint32_t)p) & EX3_MASK));
    if (upto > X.flushed) X.flushed = upto;
}

// Executes [Q0, limit) with limit <= Q0 + EX3_SUB.  stag_sa: the staged records (record `cur` covers byte Q0; behind the last
// record sits {end, 0x7FFFFFFF, ..}); sbits_sa: the start bitmap relative to Q0 (16-byte aligned, the end of the staged range
// counts as a start); cur is advanced to the record covering `limit`.  Every staged offset is valid.  All lanes take part.
struct Ex3Step { uint32_t v[EX3_U], ix[EX3_U], hz, rlo; };    // requested bytes; sources (position + bias; 0xFFFFFFFF: a literal); chunks holding offsets < 32; ring floor + bias

// part A of the step at Pn (its bitmap words at `wa`): which record covers each byte; literal and older-than-the-ring bytes are requested
template <bool RLE>
__device__ __forceinline__ void ex3_part_a(Ex3Step &T, const Ex3Ring &X, uint32_t stag_sa, uint32_t wa, uint32_t &cur, int32_t Pn, int32_t limit,
                                           const uint8_t *__restrict__ lit, uint32_t lane, uint32_t lt) {
    const uint4 mm = ex3_lds128(wa);
    const uint32_t m[4] = {mm.x, mm.y, mm.z, mm.w};
    T.rlo = (uint32_t)max(X.ring_floor, Pn + EX3_STEP - (int32_t)EX3_RING) + EX3_FAR_BIAS;        // (<= flushed: see the flush rule)
    T.hz = 0;
#pragma unroll
    for (int u = 0; u < EX3_U; u++) {
        const uint4 r = ex3_lds128(stag_sa + 16u * (cur + __popc(m[u] & lt)));      // {S, M, dlt, off} of the record covering this byte
        cur += __popc(m[u]);
        const int32_t p = Pn + 32 * u + (int32_t)lane;
        const bool is_m = (uint32_t)p >= r.y;
        const uint32_t idx = (uint32_t)p + EX3_FAR_BIAS - r.w;
        T.v[u] = ex3_request(is_m ? X.gblk + (p - (int32_t)r.w) : lit + ((uint32_t)p - r.z), p < limit && (is_m ? idx < T.rlo : !RLE));
        if (__any_sync(0xFFFFFFFFu, is_m && r.w < 32u)) T.hz |= 1u << u;
        T.ix[u] = is_m ? idx : 0xFFFFFFFFu;
    }
}
// part B of the step at P: chunk by chunk; bytes from the ring, bytes from inside the chunk, store into the ring
template <bool RLE>
__device__ __forceinline__ void ex3_part_b(const Ex3Step &T, const Ex3Ring &X, int32_t P, int32_t limit, uint32_t rle, uint32_t lane) {
#pragma unroll
    for (int u = 0; u < EX3_U; u++) {
        const int32_t Pu = P + 32 * u;
        if (Pu >= limit) break;
        uint32_t val = T.v[u];
        if (RLE && T.ix[u] == 0xFFFFFFFFu) val = rle;
        if (T.ix[u] - T.rlo < (uint32_t)Pu + EX3_FAR_BIAS - T.rlo) val = ex3_lds8(X.ring_sa | ((X.g0 - EX3_FAR_BIAS + T.ix[u]) & EX3_MASK));      // ring floor <= source < Pu
        if (T.hz & (1u << u)) {
            // sources inside this chunk: follow the parent links to a lane whose byte is a literal or older than the chunk (<= 5 rounds)
            const uint32_t rel = T.ix[u] - ((uint32_t)Pu + EX3_FAR_BIAS);
            uint32_t pr = rel < 32u ? rel : lane;
#pragma unroll 1
            for (int rd = 0; rd < 5; rd++) {
                const uint32_t pp = __shfl_sync(0xFFFFFFFFu, pr, pr);
                if (__all_sync(0xFFFFFFFFu, pp == pr)) break;
                pr = pp;
            }
            val = __shfl_sync(0xFFFFFFFFu, val, pr);
        }
        // (lanes past the limit write too: those ring slots are ahead of the output and are written again before anything reads them)
        ex3_sts8(X.ring_sa | ((X.g0 + (uint32_t)Pu + lane) & EX3_MASK), val);
        __syncwarp();
    }
}

// Executes [Q0, limit) with limit <= Q0 + EX3_SUB.  stag_sa: the staged records (record `cur` covers byte Q0; behind the last
// record sits {end, 0x7FFFFFFF, ..}); sbits_sa: the start bitmap relative to Q0 (16-byte aligned, the end of the staged range
// counts as a start); cur is advanced to the record covering `limit`.  Every staged offset is valid.  All lanes take part.
// Two register sets in alternation (no copies between them: a copy of a requested byte would wait for its load).
template <bool RLE>
__device__ __forceinline__ void ex3_fast(Ex3Ring &X, uint32_t stag_sa, uint32_t sbits_sa, uint32_t &cur, int32_t Q0, int32_t limit,
                                         const uint8_t *__restrict__ lit, uint32_t rle, uint32_t lane) {
    const uint32_t lt = (1u << lane) - 1u;
    Ex3Step A, B;
    int32_t P = Q0;
    uint32_t wa = sbits_sa;
    ex3_part_a<RLE>(A, X, stag_sa, wa, cur, P, limit, lit, lane, lt);
    while (P < limit) {
        // flush rule: before the next step's loads are requested, everything below its ring floor (P + 2 * EX3_STEP - EX3_RING) must be in HBM
        if (P - X.flushed >= 512) { ex3_flush_units(X, P, lane); __syncwarp(); }
        if (P + EX3_STEP < limit) { wa += 16u; ex3_part_a<RLE>(B, X, stag_sa, wa, cur, P + EX3_STEP, limit, lit, lane, lt); }
        ex3_part_b<RLE>(A, X, P, limit, rle, lane);
        P += EX3_STEP;
        if (P >= limit) break;
        if (P - X.flushed >= 512) { ex3_flush_units(X, P, lane); __syncwarp(); }
        if (P + EX3_STEP < limit) { wa += 16u; ex3_part_a<RLE>(A, X, stag_sa, wa, cur, P + EX3_STEP, limit, lit, lane, lt); }
        ex3_part_b<RLE>(B, X, P, limit, rle, lane);
        P += EX3_STEP;
    }
}

// Stages the start bitmap of [Q0, Q0 + EX3_SUB) for the records stag[0..n] (n = the end record) held by this warp: lane l owns
// records l, l + 32, ...
__device__ __forceinline__ void ex3_build_sbits(uint32_t *sbits, const Ex3Rec *stag, uint32_t n, int32_t Q0, uint32_t lane) {
    for (uint32_t w = lane; w < EX3_SBITS_WORDS; w += 32) sbits[w] = 0;
    __syncwarp();
    for (uint32_t t = lane; t <= n; t += 32) {
        const uint32_t d = stag[t].S - (uint32_t)Q0 - 1u;
        if (d < (uint32_t)EX3_SUB) atomicOr(&sbits[d >> 5], 1u << (d & 31u));
    }
    __syncwarp();
}

// the stand-alone kernel: one warp per frame, blocks and sequences in order; records from the sequence stage's pool
#ifndef EX3_MINB
#define EX3_MINB 7
#endif
__global__ void __launch_bounds__(32 * EX3_WARPS, EX3_MINB) k_exec3(const uint8_t *__restrict__ src, const zsb_frame *__restrict__ frames,
                                                             const zsb_block *__restrict__ blocks, const ZsbBlockWork *__restrict__ work,
                                                             ZsbFrameOut *fout, const uint32_t *__restrict__ exec_list, uint32_t n,
                                                             const ZsbCounters *__restrict__ cnt, const uint64_t *__restrict__ seq_pool,
                                                             const uint8_t *__restrict__ lit_pool, uint8_t *dst) {
    __shared__ __align__(1024) uint8_t rings[EX3_WARPS][EX3_RING];
    __shared__ Ex3Rec stags[EX3_WARPS][EX3_WIN + 1];
    __shared__ __align__(16) uint32_t sbitss[EX3_WARPS][EX3_SBITS_WORDS];
    if (cnt->overflow) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * EX3_WARPS + warp;
    if (gw >= n) return;
    const uint32_t f = exec_list[gw];
    const ZsbFrameOut fo = fout[f];
    if (fo.status != ZSB_OK) return;
    const zsb_frame fr = frames[f];
    uint8_t *fdst = dst + fo.dst_off;
    Ex3Rec *stag = stags[warp];
    uint32_t *sbits = sbitss[warp];
    const uint32_t stag_sa = (uint32_t)__cvta_generic_to_shared(stag), sbits_sa = (uint32_t)__cvta_generic_to_shared(sbits);
    Ex3Ring X;
    X.ring_sa = (uint32_t)__cvta_generic_to_shared(rings[warp]); X.flushed = 0; X.ring_floor = 0; X.gblk = fdst; X.g0 = (uint32_t)(uintptr_t)fdst;
    bool err = false;
    for (uint32_t kb = 0; kb < fr.n_blocks && !err; kb++) {
        const uint32_t bi = fr.first_block + kb;
        const ZsbBlockWork &W = work[bi];
        const uint32_t out_size = W.out_size;
        X.gblk = fdst + W.out_off; X.g0 = (uint32_t)(uintptr_t)X.gblk;
        if (blocks[bi].type != ZSB_BT_COMPRESSED) {
            ex3_flush_all(X, 0, lane);                                   // what earlier blocks left in the ring
            X.flushed = X.ring_floor = (int32_t)out_size;              // raw / RLE blocks were written by k_rawrle
        } else {
            const uint32_t nseq = W.nseq, regen = W.lit_regen;
            const bool is_rle = W.lit_type == ZSB_LT_RLE;
            const uint8_t *lit = W.lit_type >= ZSB_LT_COMPRESSED ? lit_pool + W.lit_buf : src + W.lit_src;      // raw: the bytes, RLE: the byte
            const uint32_t rle = is_rle ? lit[0] : 0u;
            const uint64_t *seqs = seq_pool + W.seq_buf;
            const uint32_t rep0 = W.rep_in[0], rep1 = W.rep_in[1], rep2 = W.rep_in[2];
            const uint64_t blk_off = W.out_off;
            uint32_t c_out = 0, c_lit = 0;
            // items: the sequences, then the literal tail (decoding_context.rs:101-103) as one more all-literal item
            for (uint32_t w0 = 0; w0 <= nseq && !err; w0 += EX3_WIN) {
                const uint32_t n_items = min(EX3_WIN, nseq + 1 - w0);
                const int32_t P0 = (int32_t)c_out;
                bool bad = false;
#pragma unroll
                for (uint32_t h = 0; h < EX3_WIN; h += 32) {
                    const uint32_t t = h + lane, i = w0 + t;
                    const bool is_seq = i < nseq;
                    const uint64_t rec = is_seq ? __ldg(seqs + i) : 0ull;
                    const uint32_t out_end = is_seq ? (uint32_t)rec & ZSB_REC_POS_MASK : out_size;
                    const uint32_t lit_end = is_seq ? (uint32_t)(rec >> ZSB_REC_POS_BITS) & ZSB_REC_POS_MASK : regen;
                    uint32_t p_out = __shfl_up_sync(0xFFFFFFFFu, out_end, 1), p_lit = __shfl_up_sync(0xFFFFFFFFu, lit_end, 1);
                    if (lane == 0) { p_out = c_out; p_lit = c_lit; }
                    if (h < n_items) {                                   // (warp uniform) this half holds items
                        const uint32_t last = min(31u, n_items - 1 - h);
                        c_out = __shfl_sync(0xFFFFFFFFu, out_end, last); c_lit = __shfl_sync(0xFFFFFFFFu, lit_end, last);
                    }
                    if (t < n_items) {
                        uint32_t off = 1u;
                        Ex3Rec R; R.S = p_out; R.M = p_out + (lit_end - p_lit); R.dlt = p_out - p_lit;
                        if (is_seq) {
                            const uint32_t coded = (uint32_t)(rec >> (2 * ZSB_REC_POS_BITS));
                            if (!(coded & ZSB_OFF_SYM)) off = coded;
                            else { const uint32_t sl = ZSB_OFF_SLOT(coded), b = sl == 0 ? rep0 : sl == 1 ? rep1 : rep2, dd = ZSB_OFF_DEC(coded); off = b > dd ? b - dd : 0u; }
                            // decoding_context.rs:86-90: offset 0 (a repeat offset that reached zero) or beyond what the frame has produced so far
                            if (off == 0 || (uint64_t)off > blk_off + R.M) bad = true;
                        } else R.M = out_size;                              // the tail: literals only
                        R.off = off;
                        stag[t] = R;
                    }
                }
                if (lane == 0) { Ex3Rec T; T.S = c_out; T.M = 0x7FFFFFFFu; T.dlt = 0; T.off = 1; stag[n_items] = T; }      // the end record
                err = __any_sync(0xFFFFFFFFu, bad);
                __syncwarp();
                uint32_t cur = 0;
                for (int32_t Q0 = P0; Q0 < (int32_t)c_out && !err; Q0 += EX3_SUB) {
                    ex3_build_sbits(sbits, stag, n_items, Q0, lane);
                    const int32_t lim = min((int32_t)c_out, Q0 + EX3_SUB);
                    if (is_rle) ex3_fast<true>(X, stag_sa, sbits_sa, cur, Q0, lim, lit, rle, lane);
                    else ex3_fast<false>(X, stag_sa, sbits_sa, cur, Q0, lim, lit, 0u, lane);
                }
            }
        }
        X.flushed -= (int32_t)out_size; X.ring_floor -= (int32_t)out_size;      // positions become relative to the next block
        __syncwarp();
    }
    if (err) { if (lane == 0) { fout[f].status = ZSB_E_IMPOSSIBLE_VALUE; fout[f].dst_len = 0; } return; }
    X.gblk = fdst + fo.dst_len; X.g0 = (uint32_t)(uintptr_t)X.gblk;
    ex3_flush_all(X, 0, lane);
}
