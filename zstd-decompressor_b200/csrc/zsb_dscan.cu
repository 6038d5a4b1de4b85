// zsb_dscan.cu -- zsb_scan_device: the walk over frames and blocks for a buffer that is resident in HBM.
//
// == ForwardByteParser::iter + Frame::parse (parsing.rs:29-112, frame.rs:61-230, block.rs:43-72), like zsb_scan -- and with the same walk over
//    one frame (zsb_walk.h).  What differs is the order: "each header says where the next one is" is one dependent chain through the whole
//    buffer, a DRAM round trip per header for a lane that follows it.  Here it is taken apart:
//      k_dscan_find    every 4-byte window of the buffer is tested for a frame magic (0xFD2FB528, 0x184D2A5x): the CANDIDATE frame starts
//                      (a true frame start is always one; a magic inside compressed data is a rare false candidate)
//      k_dscan_parse   a lane per candidate walks its frame: header, block headers, checksum -> where the next frame would start
//      k_dscan_link    the candidate (if any) at that position becomes its successor (hash table position -> candidate)
//      k_dscan_jump    pointer doubling over the successor lists: 2^k-th successors and the number of frames to the end of the chain
//      k_dscan_order   frame f is the f-th node of the chain that starts at offset 0 (binary decomposition of f over the jump tables);
//                      false candidates are never on that chain
//      k_dscan_emit    a lane per frame walks it once more and writes its block descriptors behind those of the frames before it
//    The chain ends at the end of the buffer, at a frame that fails, or at a position that holds no magic; the host finishes those last two
//    cases with the bytes concerned (the host walker on the failing frame's bytes: the error, its payload and the reference's eager section
//    parsing are zsb_scan's by construction).
#include <cuda_runtime.h>
#include "zsb_common.h"
#include "zsb_walk.h"
#include "zsb_kernels.h"

#define DSCAN_THREADS 256

__device__ __forceinline__ uint32_t dscan_hash(uint64_t pos, uint32_t mask) {
    uint64_t h = (pos + 1) * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(h >> 32) & mask;
}
__device__ __forceinline__ uint32_t dscan_lookup(const unsigned long long *keys, const uint32_t *vals, uint32_t mask, uint64_t pos) {
    for (uint32_t h = dscan_hash(pos, mask);; h = (h + 1) & mask) {
        const unsigned long long k = keys[h];
        if (k == pos + 1) return vals[h];
        if (k == 0) return 0xFFFFFFFFu;
    }
}

// pos_out == nullptr: count only.  Every thread looks at the 16 positions of one aligned 16-byte unit: one coalesced 128-bit load and the first
// word of the next unit (scalar loads at a stride of 16 bytes cost 16 sectors per warp instruction: measured 0.9 TB/s).  The unit that holds
// the first byte may start up to 15 bytes before the buffer and the last one may end behind it: inside the 128-byte line the ABI asks for.
__global__ void __launch_bounds__(DSCAN_THREADS) k_dscan_find(const uint8_t *__restrict__ src, uint64_t n, unsigned long long *count, uint64_t *pos_out, uint64_t cap) {
    const uint32_t a = (uint32_t)((uintptr_t)src & 15);
    const uint4 *A = reinterpret_cast<const uint4 *>(src - a);                  // byte p of the buffer is byte p + a from A
    const uint64_t units = (n + a + 15) / 16;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < units; c += stride) {
        const uint4 q = __ldg(A + c);
        const uint32_t v[5] = {q.x, q.y, q.z, q.w, c + 1 < units ? __ldg(reinterpret_cast<const uint32_t *>(A + c + 1)) : 0u};
        const int64_t p0 = (int64_t)(16 * c) - (int64_t)a;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const uint32_t win = __funnelshift_r(v[j >> 2], v[(j >> 2) + 1], 8 * (j & 3));
            const int64_t p = p0 + j;
            if (zsb_is_frame_magic(win) && p >= 0 && (uint64_t)p + 4 <= n) {
                const unsigned long long i = atomicAdd(count, 1ull);
                if (pos_out && i < cap) pos_out[i] = (uint64_t)p;
            }
        }
    }
}

struct DscanEmitNone { __host__ __device__ void operator()(const zsb_block &) const {} };
struct DscanEmitAt { zsb_block *out; __host__ __device__ void operator()(const zsb_block &b) { *out++ = b; } };

__global__ void __launch_bounds__(DSCAN_THREADS) k_dscan_parse(const uint8_t *__restrict__ src, uint64_t n, uint32_t flags, uint64_t max_window,
                                                               const uint64_t *__restrict__ pos_list, uint32_t ncand, ZsbDscanCand *cand,
                                                               unsigned long long *keys, uint32_t *vals, uint32_t mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncand) return;
    const uint64_t pos = pos_list[i];
    ZsbDscanCand c;
    memset(&c, 0, sizeof c);
    ZsbWalkErr e = {ZSB_OK, 0, 0};
    uint64_t end = pos; uint32_t nb = 0;
    const bool ok = zsb_walk_frame(src, n, pos, flags, max_window, 0u, c.f, end, e, nb, DscanEmitNone());
    c.f.src_off = pos; c.f.src_len = ok ? end - pos : n - pos; c.f.n_blocks = ok ? nb : 0u; c.f.status = ok ? ZSB_OK : e.code;     // (ZsbScanner::next: a failed frame runs to the end of the buffer and has no blocks)
    c.end = end; c.ok = ok ? 1 : 0; c.next = ncand; c.ea = ok ? 0 : e.a; c.eb = ok ? 0 : e.b; c.n_emitted = nb;
    cand[i] = c;
    for (uint32_t h = dscan_hash(pos, mask);; h = (h + 1) & mask) {
        const unsigned long long old = atomicCAS(keys + h, 0ull, (unsigned long long)(pos + 1));
        if (old == 0ull) { vals[h] = i; break; }
    }
}

// successor of every candidate; jump[0]; dist[0]; the node at offset 0 -> head[0]
__global__ void __launch_bounds__(DSCAN_THREADS) k_dscan_link(ZsbDscanCand *cand, uint32_t ncand, uint64_t n, const unsigned long long *__restrict__ keys,
                                                              const uint32_t *__restrict__ vals, uint32_t mask, uint32_t *jump0, uint32_t *dist0, uint32_t *head) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > ncand) return;
    if (i == ncand) { jump0[i] = ncand; dist0[i] = 0; head[0] = dscan_lookup(keys, vals, mask, 0); return; }     // the sink
    uint32_t nx = ncand;
    if (cand[i].ok && cand[i].end < n) { const uint32_t j = dscan_lookup(keys, vals, mask, cand[i].end); if (j != 0xFFFFFFFFu) nx = j; }
    cand[i].next = nx;
    jump0[i] = nx; dist0[i] = 1;
}

__global__ void __launch_bounds__(DSCAN_THREADS) k_dscan_jump(const uint32_t *__restrict__ jk, uint32_t *jk1, const uint32_t *__restrict__ dk, uint32_t *dk1, uint32_t ncand1) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncand1) return;
    const uint32_t j = jk[i];
    jk1[i] = jk[j]; dk1[i] = dk[i] + dk[j];
}

// frame f = the f-th node of the chain from `start`; tail[0..5] = {ok, end, node, ea, eb, n_emitted} of the last one
__global__ void __launch_bounds__(DSCAN_THREADS) k_dscan_order(const uint32_t *__restrict__ jump, uint32_t levels, uint32_t ncand1, uint32_t start, uint32_t nfr,
                                                               const ZsbDscanCand *__restrict__ cand, uint32_t *order, zsb_frame *frames_out, uint64_t *tail) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfr) return;
    uint32_t node = start;
    for (uint32_t k = 0; k < levels; k++) if ((f >> k) & 1u) node = jump[(size_t)k * ncand1 + node];
    order[f] = node;
    frames_out[f] = cand[node].f;
    if (f == nfr - 1) { const ZsbDscanCand &c = cand[node]; tail[0] = (uint64_t)c.ok; tail[1] = c.end; tail[2] = node; tail[3] = c.ea; tail[4] = c.eb; tail[5] = c.n_emitted; }
}

__global__ void __launch_bounds__(DSCAN_THREADS) k_dscan_emit(const uint8_t *__restrict__ src, uint64_t n, uint32_t flags, uint64_t max_window,
                                                              const zsb_frame *__restrict__ frames, const uint32_t *__restrict__ first_block, uint32_t nfr, zsb_block *blocks_out) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfr || frames[f].status != ZSB_OK) return;
    zsb_frame tmp; ZsbWalkErr e = {ZSB_OK, 0, 0}; uint64_t end = 0; uint32_t nb = 0;
    DscanEmitAt em = {blocks_out + first_block[f]};
    zsb_walk_frame(src, n, frames[f].src_off, flags, max_window, f, tmp, end, e, nb, em);
}

static inline uint32_t dscan_grid(uint64_t items) { return (uint32_t)((items + DSCAN_THREADS - 1) / DSCAN_THREADS); }

void zsbk_dscan_find(cudaStream_t st, const uint8_t *src, uint64_t n, unsigned long long *count, uint64_t *pos_out, uint64_t cap, int n_sm) {
    uint64_t g = ((n + 30) / 16 + DSCAN_THREADS - 1) / DSCAN_THREADS;
    const uint64_t gmax = (uint64_t)(n_sm > 0 ? n_sm : 148) * 16;
    if (g > gmax) g = gmax;
    if (g == 0) g = 1;
    k_dscan_find<<<(uint32_t)g, DSCAN_THREADS, 0, st>>>(src, n, count, pos_out, cap);
}
void zsbk_dscan_parse(cudaStream_t st, const uint8_t *src, uint64_t n, uint32_t flags, uint64_t max_window, const uint64_t *pos_list, uint32_t ncand,
                      ZsbDscanCand *cand, unsigned long long *keys, uint32_t *vals, uint32_t mask) {
    if (ncand) k_dscan_parse<<<dscan_grid(ncand), DSCAN_THREADS, 0, st>>>(src, n, flags, max_window, pos_list, ncand, cand, keys, vals, mask);
}
void zsbk_dscan_link(cudaStream_t st, ZsbDscanCand *cand, uint32_t ncand, uint64_t n, const unsigned long long *keys, const uint32_t *vals, uint32_t mask,
                     uint32_t *jump, uint32_t levels, uint32_t *dist_a, uint32_t *dist_b, uint32_t *head) {
    const uint32_t n1 = ncand + 1;
    k_dscan_link<<<dscan_grid(n1), DSCAN_THREADS, 0, st>>>(cand, ncand, n, keys, vals, mask, jump, dist_a, head);
    uint32_t *da = dist_a, *db = dist_b;
    for (uint32_t k = 0; k + 1 <= levels; k++) {                      // jump[k + 1] and the distances over 2^(k + 1) hops; the last round only completes the distances
        k_dscan_jump<<<dscan_grid(n1), DSCAN_THREADS, 0, st>>>(jump + (size_t)k * n1, jump + (size_t)(k + 1) * n1, da, db, n1);
        uint32_t *t = da; da = db; db = t;
    }
    // (levels rounds: the final distances are in dist_a when `levels` is even, else in dist_b)
}
void zsbk_dscan_order(cudaStream_t st, const uint32_t *jump, uint32_t levels, uint32_t ncand, uint32_t start, uint32_t nfr, const ZsbDscanCand *cand,
                      uint32_t *order, zsb_frame *frames_out, uint64_t *tail) {
    if (nfr) k_dscan_order<<<dscan_grid(nfr), DSCAN_THREADS, 0, st>>>(jump, levels, ncand + 1, start, nfr, cand, order, frames_out, tail);
}
void zsbk_dscan_emit(cudaStream_t st, const uint8_t *src, uint64_t n, uint32_t flags, uint64_t max_window, const zsb_frame *frames, const uint32_t *first_block,
                     uint32_t nfr, zsb_block *blocks_out) {
    if (nfr) k_dscan_emit<<<dscan_grid(nfr), DSCAN_THREADS, 0, st>>>(src, n, flags, max_window, frames, first_block, nfr, blocks_out);
}
