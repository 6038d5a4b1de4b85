// zsb_bits.h -- bit readers with the exact semantics of the reference's parsers (parsing.rs).
//
//  FwdBits   == ForwardBitParser  (parsing.rs:114-189): LSB-first little-endian fields, used for the
//               FSE table descriptions (<= 16 bits per read here).
//  BackWin   == BackwardBitParser (parsing.rs:191-259): the stream is read from its END; the last
//               byte carries the end marker (highest set bit); below it fields are read MSB-first
//               walking towards the stream start.  Instead of reversing a copy of the stream like
//               the reference (parsing.rs:208), BackWin keeps a 128-bit register window (hi:lo,
//               top-aligned) refilled with ALIGNED 64-bit loads, so every field of one sequence
//               (<= 64 bits in the fast path) is extracted from registers with independent shifts.
//               There is no zero-extension past the stream start: over-reading is an error
//               (NotEnoughBits), tracked by `rem`.
#pragma once
#include "zsb_common.h"

struct FwdBits {
    const uint8_t *p;   // first byte
    uint32_t nbits;     // total readable bits
    uint32_t pos;       // bits consumed
};
ZSB_HD void fwd_init(FwdBits &f, const uint8_t *p, uint64_t nbytes) {
    f.p = p; f.nbits = nbytes > 0x1FFFFFFFu ? 0xFFFFFFF8u : (uint32_t)(nbytes * 8); f.pos = 0;
}
// up to 16 bits, LSB first, without consuming.  Returns false on NotEnoughBits (parsing.rs:174-179).
ZSB_HD bool fwd_peek(const FwdBits &f, uint32_t n, uint32_t &v) {
    if (f.nbits - f.pos < n) return false;
    uint32_t byte = f.pos >> 3, sh = f.pos & 7, last = (f.nbits >> 3);
    uint32_t w = f.p[byte];
    if (byte + 1 < last) w |= (uint32_t)f.p[byte + 1] << 8;
    if (byte + 2 < last) w |= (uint32_t)f.p[byte + 2] << 16;
    v = (w >> sh) & ((1u << n) - 1u);
    return true;
}
ZSB_HD bool fwd_take(FwdBits &f, uint32_t n, uint32_t &v) {
    if (!fwd_peek(f, n, v)) return false;
    f.pos += n; return true;
}
ZSB_HD uint32_t fwd_bytes_read(const FwdBits &f) { return (f.pos >> 3) + ((f.pos & 7) ? 1u : 0u); }  // parsing.rs:122-126

// ---- backward window ------------------------------------------------------------------------------
struct BackWin {
    uint64_t hi, lo;      // unread bits, top-aligned in hi:lo
    uint64_t pref;        // prefetched next aligned word (valid if next >= minw)
    const uint8_t *base8; // 8-byte aligned base the word indices refer to
    int64_t next;         // index of the word held in `pref`
    int64_t minw;         // lowest word that may be loaded
    int32_t avail;        // valid bits in hi:lo
    int64_t rem;          // bits of the stream not yet consumed (negative = over-read)
};

// stream = src[start, end).  Returns ZSB_OK, ZSB_E_EMPTY_INPUT_DATA or ZSB_E_NULL_BYTE
// (BackwardBitParser::new parsing.rs:200-220).  `src_end` = one past the last loadable byte of the buffer.
ZSB_HD int back_init(BackWin &b, const uint8_t *src, uint64_t start, uint64_t end, uint64_t src_end) {
    if (end <= start) return ZSB_E_EMPTY_INPUT_DATA;
    uint32_t lastb = src[end - 1];
    if (lastb == 0) return ZSB_E_NULL_BYTE;
    uintptr_t a = (uintptr_t)src;
    uint32_t mis = (uint32_t)(a & 7);
    b.base8 = src - mis;
    int64_t top = (int64_t)(end - 1 + mis) * 8 + zsb_flog2(lastb);   // absolute bit index of the marker
    b.rem = top - (int64_t)(start + mis) * 8;
    b.minw = (int64_t)((start + mis) >> 3);
    int64_t W = (top - 1) >> 6;            // word holding the first unread bit (top >= 1 always)
    if (top == 0) W = 0;
    int32_t cnt = (int32_t)(top - W * 64); // 0..64 unread bits inside that word
    uint64_t w0;
    if ((uint64_t)(W * 8 + 8) > src_end + mis) {   // top word pokes past the buffer: assemble from bytes
        w0 = 0;
        for (uint64_t i = (uint64_t)W * 8; i < src_end + mis; i++) w0 |= (uint64_t)b.base8[i] << (8 * (i - (uint64_t)W * 8));
    } else w0 = zsb_ld64(b.base8, W);
    b.hi = zsb_shl64(w0, (uint32_t)(64 - cnt));
    b.lo = 0; b.avail = cnt;
    b.next = W - 1;
    b.pref = (b.next >= b.minw) ? zsb_ld64(b.base8, b.next) : 0;
    return ZSB_OK;
}
// bring the window to more than 64 valid bits (or to the end of the stream data)
ZSB_HD void back_refill(BackWin &b) {
    if (b.avail <= 64) {
        uint64_t w = b.pref;
        b.hi |= zsb_shr64(w, (uint32_t)b.avail);
        b.lo = zsb_shl64(w, (uint32_t)(64 - b.avail));
        b.avail += 64;
        b.next -= 1;
        b.pref = (b.next >= b.minw) ? zsb_ld64(b.base8, b.next) : 0;
#if defined(__CUDA_ARCH__)
        if ((b.next & 15) == 0 && b.next - 48 >= b.minw) zsb_prefetch(b.base8 + 8 * (b.next - 48));
#endif
    }
}
// n (<= 32) bits starting `off` bits below the window top; off + n <= 64
ZSB_HD uint32_t back_peek(const BackWin &b, uint32_t off, uint32_t n) {
    return (uint32_t)zsb_shr64(zsb_shl64(b.hi, off), 64 - n);
}
// drop t (<= 64) bits
ZSB_HD void back_consume(BackWin &b, uint32_t t) {
    b.hi = zsb_shl64(b.hi, t) | zsb_shr64(b.lo, 64 - t);
    b.lo = zsb_shl64(b.lo, t);
    b.avail -= (int32_t)t;
    b.rem -= t;
}
// BackwardBitParser::take for n <= 32 (refill first).  Over-read is detected by rem < 0 afterwards.
ZSB_HD uint32_t back_take(BackWin &b, uint32_t n) {
    back_refill(b);
    uint32_t v = back_peek(b, 0, n);
    back_consume(b, n);
    return v;
}
