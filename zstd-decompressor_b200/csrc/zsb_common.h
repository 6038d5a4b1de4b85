// zsb_common.h -- types and helpers shared by the CUDA kernels and the host side.
//
// Everything marked ZSB_HD is lane-serial logic (one lane = one block or one stream) written so that
// the very same source also compiles with g++ for the CPU-only differential tests in tests/emul/.
// That build is test infrastructure; the product path is the sm_100a kernels in zsb_kernels.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include "../../include/zsb.h"

#if defined(__CUDACC__)
#define ZSB_HD __host__ __device__ __forceinline__
#define ZSB_HDN __host__ __device__
#else
#define ZSB_HD inline
#define ZSB_HDN inline
#endif

// ---- format constants (RFC 8878; reference: fse.rs:13-14, sequence.rs:95-97) -------------------
#define ZSB_MAX_AL 9
#define ZSB_MAX_LL_CODE 35
#define ZSB_MAX_OF_CODE 31
#define ZSB_MAX_ML_CODE 52
#define ZSB_HUF_MAX_BITS 11
#define ZSB_MAX_NSEQ_PER_BLOCK 43691u   // ml >= 3 and a block regenerates <= 128 KiB

// internal status: a fast path met something unusual and the careful (reference-order) path must decide
#define ZSB_NEEDS_SLOW (-1)

// block types in zsb_block.type
#define ZSB_BT_RAW 0
#define ZSB_BT_RLE 1
#define ZSB_BT_COMPRESSED 2
#define ZSB_BT_SKIPPABLE 4

// literals types (literals.rs:38-43)
#define ZSB_LT_RAW 0
#define ZSB_LT_RLE 1
#define ZSB_LT_COMPRESSED 2
#define ZSB_LT_TREELESS 3
#define ZSB_LT_NONE 0xFF

// sequence table modes (sequences.rs:240-256)
#define ZSB_M_PREDEFINED 0
#define ZSB_M_RLE 1
#define ZSB_M_FSE 2
#define ZSB_M_REPEAT 3

// Packed sequence record written by the sequence decoder and consumed by the executor:
//   bits  0..17  out_end : bytes of this block regenerated once this sequence is executed
//   bits 18..35  lit_end : literals consumed once this sequence is executed
//   bits 36..63  off     : match offset, or a symbolic reference to the repeat-offset history
//                          the block started with (bit 27 set): slot in bits 25..26, bits 0..24 =
//                          how many times the "offset - 1" rule was applied to it.
#define ZSB_REC_POS_BITS 18
#define ZSB_REC_POS_MASK 0x3FFFFu
#define ZSB_OFF_SYM 0x8000000u
#define ZSB_OFF_SLOT(o) (((o) >> 25) & 3u)
#define ZSB_OFF_DEC(o) ((o) & 0x1FFFFFFu)
#define ZSB_OFF_MAX 0x7FFFFFFu   // largest real offset representable (128 MiB - 1)

// FSE decoding-table cell, one 32-bit word (sequence tables and Huffman-weight table):
//   bits 0..5 xb (extra bits of the symbol's code)   8..12 nb (state bits to read)
//   bits 16..21 code (symbol, 63 = not a legal code)   22..31 base (next-state baseline)
// The layout serves the fast sequence path (zsb_seqfast.h, k_seq): the sum of three cells has the total extra bits in byte 0
// and the total state bits in byte 1 (no carries: <= 63 and <= 27), so the sum itself is the funnel-shift amount that skips the
// extra bits (the shift takes the low 5 bits, bit 5 selects the word pair), one dot product with the byte weights 1,1 gives
// the bits a sequence consumes, and base comes out with one shift.
#define ZSB_CELL(nb, xb, base, code) ((uint32_t)(xb) | ((uint32_t)(nb) << 8) | ((uint32_t)(code) << 16) | ((uint32_t)(base) << 22))
#define ZSB_CELL_NB(e) (((e) >> 8) & 0x1Fu)
#define ZSB_CELL_XB(e) ((e) & 0x3Fu)
#define ZSB_CELL_CODE(e) (((e) >> 16) & 0x3Fu)
#define ZSB_CELL_BASE(e) ((e) >> 22)

// Per-block working record in HBM (one per zsb_block).  Filled by the section-header parse, the
// per-frame chain pass, the entropy kernels and the output planner, in that order.
struct ZsbBlockWork {
    // literals section
    uint64_t lit_src;        // raw: first literal byte; RLE: the byte; huffman: first stream byte (after jump table)
    uint64_t huf_desc;       // Huffman tree description (own, or inherited by a treeless block)
    uint64_t huf_desc_end;   // end of the literals payload that holds the description
    uint32_t lit_regen;      // Regenerated_Size
    uint32_t stream_size[4]; // jump table; [3] computed
    // sequences section
    uint64_t tbl_desc[3];    // FSE description per LL/OF/ML when mode == ZSB_M_FSE
    uint64_t tbl_end;        // block end (limit for the descriptions)
    uint64_t bs_off;         // sequence bitstream
    uint32_t bs_len;
    uint32_t nseq;
    uint8_t  lit_type, n_streams, raw_modes;
    uint8_t  parse_stage;    // how far parse_block got: 0 in the literals section, 1 literals section read, 2 modes byte read, 3 / 4 / 5 LL / OF / ML table read
    uint8_t  mode[3];        // effective mode after repeat resolution (never ZSB_M_REPEAT once chained)
    uint8_t  rle_sym[3];
    uint8_t  lit_inexact;    // ZSB_REFERENCE_QUIRKS: the streams do not have the shape Regenerated_Size promises (a zero or truncated jump-table entry ...):
                             // the literals are decoded the reference's way, every stream until its bits run out (huf_decode_block_ref)
    uint8_t  fused;          // 1: the fused sequence kernel (k_seqx) executed this block into its predicted place already; 2: it did, and a match offset was impossible; 3: it gave up (the block
                             // outgrew the frame's declared size): no records exist, the batch is run again
    // scratch placement
    uint64_t lit_buf;        // byte offset in the literal scratch (Huffman literals)
    uint64_t seq_buf;        // record index in the sequence scratch
    // results
    uint64_t out_off;        // frame-relative output offset
    uint32_t out_size;       // bytes this block regenerates
    uint32_t lit_used;       // sum of literal lengths over the sequences
    uint32_t rep_out[3];     // repeat offsets after the block (coded, may be symbolic)
    uint32_t rep_in[3];      // actual repeat offsets at block start
    int32_t  status;
    uint32_t err_a, err_b;   // payload of the reference's error variant where it has one
    uint32_t seq_rem0;       // fast sequence path: unread bits of the bitstream once the three initial states are read
    int32_t  lit_status;     // status of the literals stage (runs concurrently with the sequence stage; wins over `status`, literals.rs:49 runs first)
};

// Per-frame result record in HBM.
struct ZsbFrameOut {
    uint64_t dst_off, dst_len;
    uint64_t xxh64;
    int32_t  status;
    uint32_t err_a, err_b, pad;
};

// ---- small helpers ------------------------------------------------------------------------------
ZSB_HD int zsb_flog2(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
// 64-bit shifts whose amount may be 64 (result 0), as PTX shl/shr.b64 define them
ZSB_HD uint64_t zsb_shl64(uint64_t x, uint32_t n) {
#if defined(__CUDA_ARCH__)
    uint64_t r; asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r;
#else
    return n >= 64 ? 0 : x << n;
#endif
}
ZSB_HD uint64_t zsb_shr64(uint64_t x, uint32_t n) {
#if defined(__CUDA_ARCH__)
    uint64_t r; asm("shr.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r;
#else
    return n >= 64 ? 0 : x >> n;
#endif
}
ZSB_HD uint32_t zsb_shl32(uint32_t x, uint32_t n) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r;
#else
    return n >= 32 ? 0 : x << n;
#endif
}
ZSB_HD uint32_t zsb_shr32(uint32_t x, uint32_t n) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r;
#else
    return n >= 32 ? 0 : x >> n;
#endif
}
// aligned 64-bit little-endian load of word `w` (8-byte units) relative to an 8-byte aligned base
ZSB_HD uint64_t zsb_ld64(const uint8_t *base8, int64_t w) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const unsigned long long *>(base8) + w);
#else
    uint64_t v; memcpy(&v, base8 + 8 * w, 8); return v;
#endif
}
ZSB_HD void zsb_prefetch(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// Predefined distributions (sequences.rs:29-39) and code -> (baseline, extra bits) tables
// (sequence.rs:98-191), RFC 8878 3.1.1.3.2.1.1.  Kept as functions so that host and device share them.
ZSB_HD int zsb_predef_count(int type, int s) {
    // LL: 36 symbols, OF: 29, ML: 53
    if (type == 0) {
        static constexpr int8_t v[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
        return v[s];
    } else if (type == 1) {
        static constexpr int8_t v[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
        return v[s];
    }
    static constexpr int8_t v[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    return v[s];
}
ZSB_HD int zsb_predef_nsym(int type) { return type == 0 ? 36 : type == 1 ? 29 : 53; }
ZSB_HD int zsb_predef_al(int type) { return type == 1 ? 5 : 6; }
ZSB_HD int zsb_max_code(int type) { return type == 0 ? ZSB_MAX_LL_CODE : type == 1 ? ZSB_MAX_OF_CODE : ZSB_MAX_ML_CODE; }

ZSB_HD uint32_t zsb_ll_bits(uint32_t c) {
    static constexpr uint8_t v[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    return v[c];
}
ZSB_HD uint32_t zsb_ll_base(uint32_t c) {
    static constexpr uint32_t v[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64,
                            128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
    return v[c];
}
ZSB_HD uint32_t zsb_ml_bits(uint32_t c) {
    static constexpr uint8_t v[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                           1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    return v[c];
}
ZSB_HD uint32_t zsb_ml_base(uint32_t c) {
    static constexpr uint32_t v[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                            30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051,
                            4099, 8195, 16387, 32771, 65539};
    return v[c];
}
// extra bits carried by code `c` of table `type` (0 LL, 1 OF, 2 ML); 0 for illegal codes
ZSB_HD uint32_t zsb_code_xbits(int type, uint32_t c) {
    if (c > (uint32_t)zsb_max_code(type)) return 0;
    return type == 0 ? zsb_ll_bits(c) : type == 1 ? c : zsb_ml_bits(c);
}

static_assert(sizeof(ZsbBlockWork) <= 256, "ZsbBlockWork grew past the stage-buffer slot");
