// zsb_stream.h -- reading a backward bitstream at a moving cursor: the 64-bit window ending at a bit position, on the
// host from memory in place, on the GPU from a per-lane shared-memory ring that cp.async keeps two lines ahead of
// the cursor (BackwardBitParser, parsing.rs:191-259, read MSB-first from the end; see zsb_bits.h for the semantics).
// Used by the fast sequence path (zsb_seqfast.h) and the fast Huffman stream decode (zsb_huf.h).
#pragma once
#include "zsb_common.h"

ZSB_HD uint32_t zsb_prmt(uint32_t a, uint32_t b, uint32_t sel) {      // PTX prmt.b32, default mode
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a; uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
#endif
}
// high word of (hi:lo) << (n & 31)
ZSB_HD uint32_t zsb_fsl(uint32_t lo, uint32_t hi, uint32_t n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, n);
#else
    n &= 31; return n ? (hi << n) | (lo >> (32 - n)) : hi;
#endif
}

// `need` (<= 63) stream bits starting at absolute bit a >= 0, in the low bits of the result
ZSB_HD uint64_t fast_win_at(const uint8_t *base8, int64_t a, uint32_t need) {
    const int64_t wi = a >> 6; const uint32_t sh = (uint32_t)(a & 63);
    const uint64_t lo = zsb_ld64(base8, wi);
    const uint64_t hi = (sh + need > 64) ? zsb_ld64(base8, wi + 1) : 0ull;
    return zsb_shr64(lo, sh) | zsb_shl64(hi, 64 - sh);
}


struct FastWin { uint64_t lo, hi; uint32_t sh; };
ZSB_HD void fast_win_load(FastWin &f, const uint8_t *pw, int32_t top) {
    int32_t a = top - 64; a = a < 0 ? 0 : a;        // below the stream only after an over-read (reported at the end)
    const uint32_t wi = (uint32_t)a >> 6;
    f.sh = (uint32_t)a & 63u;
    f.lo = zsb_ld64(pw, wi);
    f.hi = f.sh ? zsb_ld64(pw, wi + 1) : 0ull;      // word wi+1 holds bit top-1, a stream bit, exactly when sh != 0
}
ZSB_HD uint64_t fast_win_get(const FastWin &f) { return zsb_shr64(f.lo, f.sh) | zsb_shl64(f.hi, 64 - f.sh); }

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t zsb_lds32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t zsb_lds32v(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }   // ordered after the cp.async waits
__device__ __forceinline__ uint64_t zsb_lds64v(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }

// The bitstream of one lane staged through a shared-memory ring of four lines (LB = log2 of the line size in
// bytes: 128-byte lines for the sequence bitstream, 64-byte lines for the Huffman streams), filled with cp.async
// two lines ahead of the (backward moving) cursor: the window loads of the chain become shared-memory loads with
// a fixed latency instead of global loads whose misses would stall every chain of the warp.
// Bit positions are relative to `pl`, a line-aligned address at least 16 bytes below the stream.
struct StreamRing { uint32_t sa; const uint8_t *pl; int32_t low; };   // ring address, line base, lowest line requested
// PF > 0: also asks L2 for the line PF lines further down, so that by the time the ring requests it the bytes come from L2 and not from
// HBM (the ring is only one line ahead of the cursor: an HBM round trip under load outlasts the ~8 chain steps between two top-ups,
// and the warp-wide wait for the previous request then stalls all chains of the warp)
template <int LB, int PF = 0> __device__ __forceinline__ void sr_fetch(const StreamRing &r, int32_t line) {
    const uint8_t *g = r.pl + ((size_t)(uint32_t)line << LB);
    if (PF > 0 && line >= PF) {
#pragma unroll
        for (int k = 0; k < (1 << LB) / 32; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(g - ((size_t)PF << LB) + 32 * k));
    }
    const uint32_t d = r.sa + (((uint32_t)line & 3u) << LB);
#pragma unroll
    for (int k = 0; k < (1 << LB) / 16; k++) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16 * k), "l"(g + 16 * k) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int LB, int PF = 0> __device__ __forceinline__ void sr_init(StreamRing &r, int32_t top) {
    const int32_t l0 = (top - 1) >> (LB + 3);
    r.low = l0 - 2 < 0 ? 0 : l0 - 2;
    for (int32_t l = l0; l >= r.low; l--) sr_fetch<LB>(r, l);
    if (PF > 0) {       // the lines between the ring and the first line a top-up will ask L2 for
        for (int32_t l = r.low - 1; l >= 0 && l >= r.low - PF; l--)
            for (int k = 0; k < (1 << LB) / 32; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(r.pl + ((size_t)(uint32_t)l << LB) + 32 * k));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// Keeps the ring ahead of the cursor: when the window ending at `top` has entered line X (<= low+1), line X-2 is
// requested and line X-1 awaited.  Must run at least once per line of progress: every step (sr_load), or every k steps
// when k steps cannot consume a whole line (sr_check + sr_load_nocheck; all lanes of a warp then refill in the same
// pass instead of diverging step by step).
template <int LB, int PF = 0> __device__ __forceinline__ void sr_check(StreamRing &r, int32_t top) {
    int32_t a = top - 64; a = a < 0 ? 0 : a;
    if ((a >> (LB + 3)) <= r.low + 1 && r.low > 0) { r.low--; sr_fetch<LB, PF>(r, r.low); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
}
// window words for the 64 bits ending at `top`
template <int LB> __device__ __forceinline__ void sr_load_nocheck(const StreamRing &r, FastWin &f, int32_t top) {
    int32_t a = top - 64; a = a < 0 ? 0 : a;          // below the stream only after an over-read (reported at the end)
    const uint32_t wi = (uint32_t)a >> 6;
    f.sh = (uint32_t)a & 63u;
    const uint32_t wm = (4u << (LB - 3)) - 1u;         // words in the ring - 1
    f.lo = zsb_lds64v(r.sa + (wi & wm) * 8u);
    f.hi = f.sh ? zsb_lds64v(r.sa + ((wi + 1) & wm) * 8u) : 0ull;
}
template <int LB> __device__ __forceinline__ void sr_load(StreamRing &r, FastWin &f, int32_t top) {
    sr_check<LB>(r, top);
    sr_load_nocheck<LB>(r, f, top);
}
#endif

