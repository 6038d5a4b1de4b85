// zsb_host.cu -- C ABI (include/zsb.h): container walk on the host, device context, batch decode.
//
// Host work is limited to what the reference does before any entropy decoding: finding frame and
// block boundaries (frame.rs:61-230, block.rs:43-72 minus section parsing).  Everything below a
// block header runs in the kernels of zsb_kernels.cu.  There is no CPU decode path in this file.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <new>
#include <string>
#include <vector>
#include "zsb_kernels.h"
#include "zsb_scan.h"
#include "zsb_walk.h"

// ======================================================================================= context
namespace {
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
const int kMaxKernels = 16;
const int kProfRing = 32;     // launches whose per-kernel events are kept
const uint32_t kBigFrameBlocks = 64;   // frames with more blocks than this are executed by a whole CTA (k_exec), not a warp (k_exec2)
const uint64_t kLitOverflow = 4u << 20; // ZSB_REFERENCE_QUIRKS: room behind the literal scratch for blocks whose (corrupted) streams regenerate more than announced
const size_t kTinyBatchFrames = 64;    // ... and in batches of at most this many frames every frame: a warp on its own takes 1.2 ms for one 128 KiB block (k_exec2), k_link 0.2 ms
const size_t kSmallBatchFrames = 296;  // batches of at most this many frames (two per SM) give every multi-block frame a CTA
}  // namespace

struct zsb_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, aux_stream = nullptr;   // aux: the literals stage runs beside the sequence stage
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_huf[kProfRing][2] = {};
    DevBuf src, frames, blocks, work, fout, huf_list, seq_list, rawrle_list, exec_list, exec2_list, xxh_list, lit_pool, seq_pool, slow_list, counters, dst, stage, pre_off, wave;
    DevBuf dscan, dscan_blocks, dscan_pos;                                  // zsb_scan_device: candidates, hash table, jump tables / block descriptors
    DevBuf link_ent, link_meta;                                    // k_link: one 32-bit entry per output byte of the listed frames; frame list, block list, tickets
    std::vector<uint64_t> h_link_meta;                             // host image of link_meta (frames: 2 words each, then blocks: 1 word each, then tickets)
    uint32_t n_link = 0, n_link_blocks = 0;                        // frames / blocks executed by k_link_init + k_link_resolve
    int link_mode = 1;                                             // ZSB_LINK: 0 never (k_exec), 1 frames k_exec would take (default), 2 every frame
    int n_sm = 148;
    std::string last_err;
    // prepared batch
    const uint8_t *d_src = nullptr; uint8_t *d_dst = nullptr; uint8_t *h_dst = nullptr;
    size_t src_len = 0, dst_cap = 0;
    uint32_t nf = 0, nb = 0, ncomp = 0, n_rawrle = 0, n_exec = 0, n_exec2 = 0, n_xxh = 0, flags = 0;
    uint64_t lit_cap = 0, seq_cap = 0;
    std::vector<zsb_frame> h_frames;
    std::vector<zsb_block> h_blocks;
    std::vector<uint32_t> h_xxh_list, h_rawrle, h_exec, h_exec2;   // host copies stay alive while their uploads are in flight
    std::vector<uint64_t> h_pre_off;                               // where every frame is expected in dst (k_seqx), ~0: not known before decoding
    int seqx_state = 0;                                            // zsb_last_seqx_state
    uint32_t wave_max = 3;                                         // ZSB_WAVE: most CTAs per frame k_exec may use (1: one CTA per frame, block after block, hashing beside them)
    bool wave_forced = false;                                      // ... set by the environment: used even when there is nothing to hash
    uint32_t wave_ctas = 1;                                        // CTAs per frame of k_exec (wavefront mode when > 1)
    bool use_seqx = false;                                         // this batch runs k_seqx
    std::vector<uint32_t> h_err_a, h_err_b;                        // per-frame error payloads of the last finished batch (zsb_decode_errors)
    std::vector<zsb_ctx *> subs;          // child contexts of the pipelined host path (own stream + scratch each)
    bool is_sub = false;
    bool low_latency = false;             // pipelined path, first shards: CTA-per-frame execution (k_exec: 0.15 ms per block instead of 1.2 ms per frame, at a third of the throughput)
    // pipelined path: all shards upload on one stream and download on another, in shard order (copies issued from
    // several streams share the copy engines in no particular order, which delays the first shards)
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaEvent_t ev_up = nullptr, ev_kdone = nullptr, ev_down = nullptr;
    // pipelined path: counters and per-frame results are read back right behind the kernels, ahead of the shard's big download
    // (page-locked staging: pin = ZsbCounters, then nf x ZsbFrameOut)
    uint8_t *pin = nullptr; size_t pin_cap = 0; bool pin_valid = false;
    uint64_t eager_d2h = 0;               // pipelined path: bytes of output to send to the host right behind the kernels (size known from the headers)
    bool prepared = false, launched = false;
    bool debug_sync = false;   // ZSB_DEBUG=1
    const char *dbg_prev = "the uploads";
    bool seqx = false;         // ZSB_SEQX=1: sequence decoding and execution in one kernel (k_seqx) where a frame's place is known beforehand; measured
                               // slower than k_seq + k_exec2 (5.6 ms against 2.9 ms on C2: 64 registers per thread and 28 KiB of L1 for 29 warps), kept as an experiment
    bool overlap = false;      // ZSB_OVERLAP=1: literals stage on the auxiliary stream beside k_seq (measured slower: both are latency bound and share schedulers)
    // profiling
    bool profile = false;
    cudaEvent_t ev[kProfRing][kMaxKernels + 1] = {};
    const char *kname[kMaxKernels] = {};
    float kms[kMaxKernels] = {};
    int nk = 0, launches = 0;
    int prof_slot = 0, prof_count = 0;   // ring position / launches recorded since profiling was switched on
    // ZSB_PIPE_TRACE=1: device timeline of the pipelined host path (base, uploads done, kernels done, download done), printed to stderr
    bool trace = false;
    cudaEvent_t ev_tr[4] = {};
};

#define CK(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { (ctx)->last_err = std::string(#call) + ": " + cudaGetErrorString(e__); (void)cudaGetLastError(); return ZSB_E_CUDA; } } while (0)

extern "C" int zsb_ctx_create(zsb_ctx **out, int device) {
    if (!out) return ZSB_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { (void)cudaGetLastError(); return ZSB_E_CUDA; }
    zsb_ctx *c = new (std::nothrow) zsb_ctx();
    if (!c) return ZSB_E_NOMEM;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        zsbk_init() != cudaSuccess) { (void)cudaGetLastError(); delete c; return ZSB_E_CUDA; }
    c->stream = c->own_stream;
    { const char *e = getenv("ZSB_OVERLAP"); c->overlap = e && *e && *e != '0'; }
    { const char *e = getenv("ZSB_SEQX"); c->seqx = e && *e && *e != '0'; }
    { const char *e = getenv("ZSB_DEBUG"); c->debug_sync = e && *e && *e != '0'; }
    { const char *e = getenv("ZSB_WAVE"); const int v = e ? atoi(e) : 3; c->wave_forced = e != nullptr; c->wave_max = v < 1 ? 1u : v > 64 ? 64u : (uint32_t)v; }
    { const char *e = getenv("ZSB_PIPE_TRACE"); c->trace = e && *e && *e != '0'; }
    { const char *e = getenv("ZSB_LINK"); c->link_mode = e ? atoi(e) : 1; }
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) c->n_sm = v; else (void)cudaGetLastError(); }
    if (c->trace) for (int i = 0; i < 4; i++) cudaEventCreate(&c->ev_tr[i]);
    for (int r = 0; r < kProfRing; r++) for (int i = 0; i <= kMaxKernels; i++) cudaEventCreate(&c->ev[r][i]);
    for (int r = 0; r < kProfRing; r++) { cudaEventCreate(&c->ev_huf[r][0]); cudaEventCreate(&c->ev_huf[r][1]); }
    if (cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_kdone, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_down, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); zsb_ctx_destroy(c); return ZSB_E_CUDA; }
    *out = c;
    return ZSB_OK;
}
extern "C" void zsb_ctx_destroy(zsb_ctx *c) {
    if (!c) return;
    for (zsb_ctx *sub : c->subs) zsb_ctx_destroy(sub);
    c->subs.clear();
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf *all[] = {&c->src, &c->frames, &c->blocks, &c->work, &c->fout, &c->huf_list, &c->seq_list, &c->rawrle_list, &c->exec_list, &c->exec2_list, &c->pre_off, &c->wave, &c->link_ent, &c->link_meta, &c->dscan, &c->dscan_blocks, &c->dscan_pos,
                     &c->xxh_list, &c->lit_pool, &c->seq_pool, &c->slow_list, &c->counters, &c->dst, &c->stage};
    for (DevBuf *b : all) b->release();
    for (int r = 0; r < kProfRing; r++) for (int i = 0; i <= kMaxKernels; i++) if (c->ev[r][i]) cudaEventDestroy(c->ev[r][i]);
    for (int r = 0; r < kProfRing; r++) for (int i = 0; i < 2; i++) if (c->ev_huf[r][i]) cudaEventDestroy(c->ev_huf[r][i]);
    for (int i = 0; i < 4; i++) if (c->ev_tr[i]) cudaEventDestroy(c->ev_tr[i]);
    if (c->pin) cudaFreeHost(c->pin);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_up) cudaEventDestroy(c->ev_up);
    if (c->ev_kdone) cudaEventDestroy(c->ev_kdone);
    if (c->ev_down) cudaEventDestroy(c->ev_down);
    if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}
extern "C" int zsb_ctx_set_stream(zsb_ctx *c, void *s) { if (!c) return ZSB_E_ARG; c->stream = s ? (cudaStream_t)s : c->own_stream; return ZSB_OK; }
extern "C" const char *zsb_last_cuda_error(const zsb_ctx *c) { return c ? c->last_err.c_str() : "no context"; }
extern "C" int zsb_ctx_set_profile(zsb_ctx *c, int en) { if (!c) return ZSB_E_ARG; c->profile = en != 0; c->prof_slot = 0; c->prof_count = 0; return ZSB_OK; }
extern "C" int zsb_last_launch_count(const zsb_ctx *c) { return c ? c->launches : 0; }
extern "C" int zsb_last_seqx_state(const zsb_ctx *c) { return c ? c->seqx_state : 0; }
extern "C" int zsb_last_kernel_times(const zsb_ctx *c, const char **names, float *ms, int cap) {
    if (!c) return 0;
    int n = c->nk < cap ? c->nk : cap;
    for (int i = 0; i < n; i++) { if (names) names[i] = c->kname[i]; if (ms) ms[i] = c->kms[i]; }
    return n;
}
extern "C" int zsb_kernel_times_avg(zsb_ctx *c, const char **names, float *ms, int cap, int *n_launches) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    const int L = c->prof_count < kProfRing ? c->prof_count : kProfRing;
    int n = c->nk < cap ? c->nk : cap;
    for (int i = 0; i < n; i++) {
        double sum = 0;
        for (int r = 0; r < L; r++) { float t = 0; if (cudaEventElapsedTime(&t, c->ev[r][i], c->ev[r][i + 1]) == cudaSuccess) sum += t; else (void)cudaGetLastError(); }
        if (names) names[i] = c->kname[i];
        if (ms) ms[i] = L ? (float)(sum / L) : 0.f;
    }
    if (n < cap && c->nk && c->overlap) {       // the literals stage ran on the auxiliary stream, beside k_seq
        double sum = 0;
        for (int r = 0; r < L; r++) { float t = 0; if (cudaEventElapsedTime(&t, c->ev_huf[r][0], c->ev_huf[r][1]) == cudaSuccess) sum += t; else (void)cudaGetLastError(); }
        if (names) names[n] = "k_huf(aux stream)";
        if (ms) ms[n] = L ? (float)(sum / L) : 0.f;
        n++;
    }
    if (n_launches) *n_launches = L;
    return n;
}

extern "C" int zsb_decode_errors(const zsb_ctx *c, uint32_t *err_a, uint32_t *err_b, size_t n_frames) {
    if (!c) return ZSB_E_ARG;
    for (size_t f = 0; f < n_frames; f++) {
        if (err_a) err_a[f] = f < c->h_err_a.size() ? c->h_err_a[f] : 0;
        if (err_b) err_b[f] = f < c->h_err_b.size() ? c->h_err_b[f] : 0;
    }
    return ZSB_OK;
}

extern "C" void *zsb_host_alloc(size_t n) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void zsb_host_free(void *p) { if (p && cudaFreeHost(p) != cudaSuccess) (void)cudaGetLastError(); }

// ======================================================================================= batch decode
extern "C" int zsb_decode_prepare(zsb_ctx *c, const uint8_t *src, size_t n, const zsb_frame *frames, size_t nf,
                                  const zsb_block *blocks, size_t nb, uint8_t *dst, size_t dst_cap, uint32_t flags) {
    if (!c || (!src && n) || (!frames && nf) || (!blocks && nb) || nf > 0x7FFFFFFFu || nb > 0x7FFFFFFFu) return ZSB_E_ARG;
    CK(c, cudaSetDevice(c->device));
    c->prepared = false; c->launched = false; c->seqx_state = 0;
    cudaStream_t st = c->up_stream ? c->up_stream : c->stream;      // every copy of this function goes to `st`
    // compressed bytes: resident already, or uploaded once (padded so that aligned 8-byte loads near the end stay inside)
    if (flags & ZSB_SRC_ON_DEVICE) c->d_src = src;
    else {
        CK(c, c->src.ensure(n + 128));                // the stream rings of k_seq / k_huf read whole aligned 128-byte lines
        if (n) CK(c, cudaMemcpyAsync(c->src.p, src, n, cudaMemcpyHostToDevice, st));
        CK(c, cudaMemsetAsync((uint8_t *)c->src.p + n, 0, 128, st));
        c->d_src = (const uint8_t *)c->src.p;
    }
    if (flags & ZSB_DST_ON_DEVICE) { c->d_dst = dst; c->h_dst = nullptr; }
    else { CK(c, c->dst.ensure(dst_cap + 64)); c->d_dst = (uint8_t *)c->dst.p; c->h_dst = dst; }
    c->src_len = n; c->dst_cap = dst_cap; c->nf = (uint32_t)nf; c->nb = (uint32_t)nb; c->flags = flags;
    c->h_frames.assign(frames, frames + nf);
    c->h_blocks.assign(blocks, blocks + nb);
    frames = c->h_frames.data(); blocks = c->h_blocks.data();      // the caller's arrays are not touched after this call returns
    // host-side work lists (block types and frame kinds are known from the scan)
    std::vector<uint32_t> &rawrle = c->h_rawrle, &execl = c->h_exec, &exec2l = c->h_exec2;
    rawrle.clear(); execl.clear(); exec2l.clear();
    c->h_xxh_list.clear();
    uint64_t ncomp = 0, lit_cap = 0, seq_cap = 0;
    for (size_t i = 0; i < nb; i++) {
        const zsb_block &b = blocks[i];
        if (b.type == ZSB_BT_COMPRESSED) {
            ncomp++;
            uint64_t lc = (uint64_t)b.size * 4; if (lc > ZSB_BLOCK_MAX) lc = ZSB_BLOCK_MAX;
            lit_cap += lc + 16;
            seq_cap += (b.size < ZSB_MAX_NSEQ_PER_BLOCK ? b.size : ZSB_MAX_NSEQ_PER_BLOCK) + 1;
        } else if (b.type == ZSB_BT_SKIPPABLE) { if (flags & ZSB_PRINT_SKIPPABLE) rawrle.push_back((uint32_t)i); }
        else rawrle.push_back((uint32_t)i);
    }
    for (size_t f = 0; f < nf; f++) {
        if (frames[f].kind != 0 || frames[f].status != ZSB_OK) continue;
        bool has_c = false;
        for (uint32_t k = 0; k < frames[f].n_blocks && !has_c; k++) has_c = blocks[frames[f].first_block + k].type == ZSB_BT_COMPRESSED;
        // frames of many blocks: one CTA per frame (k_exec); all others: one warp per frame (k_exec2)
        // ... and, in a batch too small to fill the GPU with warps (a single file of a few frames, like BASELINE config C1), every frame of
        // more than one block: a CTA takes 0.11 ms per block, a warp on its own 0.8 ms
        static const int big_env = getenv("ZSB_BIG_FRAME_BLOCKS") ? atoi(getenv("ZSB_BIG_FRAME_BLOCKS")) : -1;     // experiment knob
        const uint32_t big = c->low_latency ? 0u : big_env >= 0 ? (uint32_t)big_env : (nf <= kTinyBatchFrames && c->link_mode && !c->seqx) ? 0u : nf <= kSmallBatchFrames ? 1u : kBigFrameBlocks;
        const bool cta = has_c && frames[f].n_blocks > big;
        if (has_c) (cta ? execl : exec2l).push_back((uint32_t)f);
        // frames executed by k_exec<1024> are hashed by its trailing warp, all others by k_xxh
        if ((flags & ZSB_VERIFY_CHECKSUM) && frames[f].has_checksum && !(cta && !c->is_sub)) c->h_xxh_list.push_back((uint32_t)f);
    }
    // k_seqx: a frame whose predecessors all declare their size has a known place before anything is decoded, and its first block can
    // be executed by the sequence kernel itself.  Only frames the warp-per-frame executor would take (k_exec2: it skips what is done),
    // only when the literals are there before the sequence stage starts (no k_huf beside k_seq).  A frame that does not regenerate what
    // it declares (an error, or accepted under ZSB_REFERENCE_QUIRKS) moves the frames behind it: k_plan2 notices, the batch runs again.
    {
        c->h_pre_off.assign(nf, ~0ull);
        c->use_seqx = false;
        if (c->seqx && !c->overlap && !c->low_latency && !exec2l.empty()) {
            uint64_t off = 0; size_t e2 = 0; uint32_t placed = 0;
            for (size_t f = 0; f < nf; f++) {
                const zsb_frame &fr = frames[f];
                uint64_t len = 0;
                if (fr.status != ZSB_OK) len = 0;
                else if (fr.kind == 1) len = (flags & ZSB_PRINT_SKIPPABLE) && fr.n_blocks ? blocks[fr.first_block].size : 0;
                else if (!fr.has_content_size) break;                      // from here on nothing has a known place
                else len = fr.content_size;
                while (e2 < exec2l.size() && exec2l[e2] < f) e2++;
                const bool warp_frame = e2 < exec2l.size() && exec2l[e2] == f;
                if (warp_frame && fr.n_blocks && blocks[fr.first_block].type == ZSB_BT_COMPRESSED && len && off + len <= dst_cap) { c->h_pre_off[f] = off; placed++; }
                off += len;
            }
            c->use_seqx = placed != 0;
            if (getenv("ZSB_DEBUG")) fprintf(stderr, "zsb: k_seqx places %u of %zu frames\n", placed, nf);
        }
    }
    // Frames of many blocks (and every multi-block frame of a small batch): k_link_init + k_link_resolve -- one entry per output byte, pointer
    // jumping -- instead of a CTA that walks the frame block after block (k_exec), as long as the entries fit (4 bytes per output byte; positions
    // are 31 bits).  ZSB_LINK=0 keeps k_exec, ZSB_LINK=2 sends every frame this way (experiment).
    c->n_link = 0; c->n_link_blocks = 0; c->h_link_meta.clear();
    if (c->link_mode && !c->is_sub && !c->low_latency) {
        std::vector<uint32_t> cand, keep;
        if (c->link_mode >= 2) { cand = execl; cand.insert(cand.end(), exec2l.begin(), exec2l.end()); std::sort(cand.begin(), cand.end()); }
        else cand = execl;
        std::vector<uint64_t> fmeta, bmeta;
        std::vector<uint32_t> taken;
        uint64_t e_total = 0;
        for (uint32_t f : cand) {
            const zsb_frame &fr = frames[f];
            uint64_t bound = 0;
            for (uint32_t k = 0; k < fr.n_blocks; k++) { const zsb_block &b = blocks[fr.first_block + k]; bound += b.type == ZSB_BT_COMPRESSED ? (uint64_t)ZSB_BLOCK_MAX : b.size; }
            if (bound > dst_cap) bound = dst_cap;
            if (bound + 64 >= (1ull << 31)) { keep.push_back(f); continue; }
            const uint32_t li = (uint32_t)taken.size();
            fmeta.push_back(e_total); fmeta.push_back((uint64_t)f);
            for (uint32_t k = 0; k < fr.n_blocks; k++) bmeta.push_back((uint64_t)(fr.first_block + k) | (uint64_t)li << 32);
            e_total += (bound + 16 + 63) & ~63ull;
            taken.push_back(f);
        }
        if (!taken.empty() && c->link_ent.ensure(4 * e_total + 256) == cudaSuccess) {
            c->n_link = (uint32_t)taken.size(); c->n_link_blocks = (uint32_t)bmeta.size();
            c->h_link_meta = fmeta;
            c->h_link_meta.insert(c->h_link_meta.end(), bmeta.begin(), bmeta.end());
            c->h_link_meta.resize(c->h_link_meta.size() + (taken.size() + 1) / 2, 0ull);     // tickets
            if (c->link_mode >= 2) {
                auto drop = [&](std::vector<uint32_t> &v) { std::vector<uint32_t> r; for (uint32_t f : v) if (!std::binary_search(taken.begin(), taken.end(), f)) r.push_back(f); v.swap(r); };
                drop(execl); drop(exec2l);
                // (their checksums: k_xxh_one)
                std::vector<uint32_t> r; for (uint32_t f : c->h_xxh_list) if (!std::binary_search(taken.begin(), taken.end(), f)) r.push_back(f); c->h_xxh_list.swap(r);
            } else execl = keep;
        } else (void)cudaGetLastError();
    }
    // k_exec in wavefront mode (ZSB_WAVE=n CTAs per frame, default 3 when checksums are verified): the CTA-per-frame list is short (a single
    // large frame, a file of a few frames), so every frame gets one CTA that only hashes behind the frame's frontier and n-1 that take its
    // blocks in order (ZsbWave).  The hashing CTA pays: its warp has an SM to itself (C3, 256 MiB, verified: 314 -> 247 ms, what the frame
    // takes without its checksum).  A second executing CTA hides the staging and flushing of one block behind the execution of the next;
    // more gain nothing on text at zstd -3: half the sequences of a block wait, directly or through bytes that do, for the block before it
    // (tools/probes/c3_dependencies.py), and a batch of 32 sequences proceeds only when all of its lanes can (4 / 6 CTAs: 246 / 245 ms).
    {
        uint32_t maxb = 0;
        for (uint32_t f : execl) maxb = std::max(maxb, frames[f].n_blocks);
        uint32_t g = execl.empty() ? 1u : (uint32_t)(148 / execl.size());
        if (g > c->wave_max) g = c->wave_max;
        if (g > maxb + 1) g = maxb + 1;
        if (!c->wave_forced && !(flags & ZSB_VERIFY_CHECKSUM)) g = 1;   // nothing to hash: one CTA per frame, block after block, is 3 % faster than two that take turns
        c->wave_ctas = (c->is_sub || g < 2) ? 1u : g;
    }
    c->ncomp = (uint32_t)ncomp; c->n_rawrle = (uint32_t)rawrle.size(); c->n_exec = (uint32_t)execl.size(); c->n_exec2 = (uint32_t)exec2l.size(); c->n_xxh = (uint32_t)c->h_xxh_list.size();
    if (lit_cap > c->lit_cap) c->lit_cap = lit_cap;
    if (seq_cap > c->seq_cap) c->seq_cap = seq_cap;
    CK(c, c->frames.ensure(sizeof(zsb_frame) * (nf + 1)));
    CK(c, c->blocks.ensure(sizeof(zsb_block) * (nb + 1)));
    CK(c, c->work.ensure(sizeof(ZsbBlockWork) * (nb + 1)));
    CK(c, c->fout.ensure(sizeof(ZsbFrameOut) * (nf + 1)));
    CK(c, c->huf_list.ensure(4 * (ncomp + 1)));
    CK(c, c->seq_list.ensure(4 * (ncomp + 1)));
    CK(c, c->slow_list.ensure(4 * (ncomp + 1)));
    CK(c, c->rawrle_list.ensure(4 * (rawrle.size() + 1)));
    CK(c, c->exec_list.ensure(4 * (execl.size() + 1)));
    CK(c, c->exec2_list.ensure(4 * (exec2l.size() + 1)));
    CK(c, c->pre_off.ensure(8 * (nf + 1)));
    if (c->wave_ctas > 1) CK(c, c->wave.ensure(32 * execl.size() + 4 * (nb + 1)));
    CK(c, c->xxh_list.ensure(4 * (c->h_xxh_list.size() + 1)));
    if (c->n_link) CK(c, c->link_meta.ensure(8 * c->h_link_meta.size()));
    CK(c, c->counters.ensure(sizeof(ZsbCounters)));
    c->lit_cap = (c->lit_cap + 15) & ~(uint64_t)15;
    CK(c, c->lit_pool.ensure(c->lit_cap + kLitOverflow + 64));
    CK(c, c->seq_pool.ensure(8 * (c->seq_cap + 8)));
    if (nf) CK(c, cudaMemcpyAsync(c->frames.p, frames, sizeof(zsb_frame) * nf, cudaMemcpyHostToDevice, st));
    if (nb) CK(c, cudaMemcpyAsync(c->blocks.p, blocks, sizeof(zsb_block) * nb, cudaMemcpyHostToDevice, st));
    if (!rawrle.empty()) CK(c, cudaMemcpyAsync(c->rawrle_list.p, rawrle.data(), 4 * rawrle.size(), cudaMemcpyHostToDevice, st));
    if (!execl.empty()) CK(c, cudaMemcpyAsync(c->exec_list.p, execl.data(), 4 * execl.size(), cudaMemcpyHostToDevice, st));
    if (!exec2l.empty()) CK(c, cudaMemcpyAsync(c->exec2_list.p, exec2l.data(), 4 * exec2l.size(), cudaMemcpyHostToDevice, st));
    if (c->use_seqx) CK(c, cudaMemcpyAsync(c->pre_off.p, c->h_pre_off.data(), 8 * nf, cudaMemcpyHostToDevice, st));
    if (c->n_link) CK(c, cudaMemcpyAsync(c->link_meta.p, c->h_link_meta.data(), 8 * c->h_link_meta.size(), cudaMemcpyHostToDevice, st));
    if (!c->h_xxh_list.empty()) CK(c, cudaMemcpyAsync(c->xxh_list.p, c->h_xxh_list.data(), 4 * c->h_xxh_list.size(), cudaMemcpyHostToDevice, st));
    if (c->trace) cudaEventRecord(c->ev_tr[1], st);
    if (c->up_stream) { CK(c, cudaEventRecord(c->ev_up, st)); CK(c, cudaStreamWaitEvent(c->stream, c->ev_up, 0)); }
    c->prepared = true;               // nothing waited for: the uploads read the context's own host copies, and `src` if it is host memory
    return ZSB_OK;
}

// ZSB_DEBUG=1: every stage is waited for before the next one is enqueued, so that a fault is reported with the name of the kernel that caused it
#define MARK(ctx, name) do { \
    if ((ctx)->debug_sync) { \
        cudaError_t e__ = cudaStreamSynchronize(st); \
        if (e__ != cudaSuccess) { (ctx)->last_err = std::string("before ") + (name) + " (after " + (ctx)->dbg_prev + "): " + cudaGetErrorString(e__); fprintf(stderr, "zsb: %s\n", (ctx)->last_err.c_str()); return ZSB_E_CUDA; } \
        (ctx)->dbg_prev = (name); \
    } \
    if ((ctx)->profile && (ctx)->nk < kMaxKernels) { (ctx)->kname[(ctx)->nk] = name; cudaEventRecord((ctx)->ev[(ctx)->prof_slot][(ctx)->nk], st); (ctx)->nk++; } } while (0)

extern "C" int zsb_decode_launch(zsb_ctx *c) {
    if (!c || !c->prepared) return ZSB_E_ARG;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint8_t *src = c->d_src;
    zsb_frame *frames = (zsb_frame *)c->frames.p; zsb_block *blocks = (zsb_block *)c->blocks.p;
    ZsbBlockWork *work = (ZsbBlockWork *)c->work.p; ZsbFrameOut *fout = (ZsbFrameOut *)c->fout.p;
    ZsbCounters *cnt = (ZsbCounters *)c->counters.p;
    c->nk = 0; c->launches = 0;
    if (c->use_seqx) c->seqx_state = 1; else if (c->seqx_state != 2) c->seqx_state = 0;
    CK(c, cudaMemsetAsync(cnt, 0, sizeof(ZsbCounters), st));
    MARK(c, "k_parse");  zsbk_parse(st, src, blocks, work, c->nb, c->flags); c->launches += c->nb ? 1 : 0;
    MARK(c, "k_plan1");  zsbk_plan1(st, frames, c->nf, blocks, c->nb, work, fout, (uint32_t *)c->huf_list.p, (uint32_t *)c->seq_list.p, cnt,
                                    c->lit_cap, c->seq_cap, c->flags); c->launches++;
    // the literals stage (k_huf) and the sequence stage (k_seq) read the same blocks and write disjoint results; with ZSB_OVERLAP=1
    // k_huf runs on the auxiliary stream beside k_seq (enqueued first: its CTAs need the larger shared-memory slice).
    // tiny low-latency shards: 8 chains per k_seq CTA instead of 32 (fewer ring bank conflicts, less phase-2 work beside the
    // producer: the first shard's output is ready 0.25 ms earlier); ZSB_LL_CHAINS overrides
    static const uint32_t ll_env = getenv("ZSB_LL_CHAINS") ? (uint32_t)atoi(getenv("ZSB_LL_CHAINS")) : 8u;
    const uint32_t ll_chains = (c->low_latency && c->ncomp <= 320) ? ll_env : 0u;
    const bool ov = c->overlap || c->low_latency;
    const uint64_t *pre_off = c->use_seqx ? (const uint64_t *)c->pre_off.p : nullptr;
    if (ov) {
        CK(c, cudaEventRecord(c->ev_fork, st));
        CK(c, cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        MARK(c, "k_seq");  zsbk_seq(st, c->ncomp, src, work, (const uint32_t *)c->seq_list.p, cnt, (uint64_t *)c->seq_pool.p, (uint32_t *)c->slow_list.p, c->is_sub, ll_chains,
                                    frames, blocks, nullptr, (const uint8_t *)c->lit_pool.p, c->d_dst);
        if (c->profile) cudaEventRecord(c->ev_huf[c->prof_slot][0], c->aux_stream);
        zsbk_huf(c->aux_stream, c->ncomp, src, c->src_len, work, (const uint32_t *)c->huf_list.p, cnt, (uint8_t *)c->lit_pool.p, c->lit_cap, 0, c->flags & ~ZSB_REFERENCE_QUIRKS);
        if (c->profile) cudaEventRecord(c->ev_huf[c->prof_slot][1], c->aux_stream);
        CK(c, cudaEventRecord(c->ev_join, c->aux_stream));
    } else {
        MARK(c, "k_huf");  zsbk_huf(st, c->ncomp, src, c->src_len, work, (const uint32_t *)c->huf_list.p, cnt, (uint8_t *)c->lit_pool.p, c->lit_cap, kLitOverflow, c->flags);
        MARK(c, pre_off ? "k_seqx" : "k_seq");
        zsbk_seq(st, c->ncomp, src, work, (const uint32_t *)c->seq_list.p, cnt, (uint64_t *)c->seq_pool.p, (uint32_t *)c->slow_list.p, c->is_sub, ll_chains,
                 frames, blocks, pre_off, (const uint8_t *)c->lit_pool.p, c->d_dst);
    }
    c->launches += c->ncomp ? 1 : 0;
    MARK(c, "k_seq_slow"); zsbk_seq_slow(st, c->ncomp, src, c->src_len, work, (const uint32_t *)c->slow_list.p, cnt, (uint64_t *)c->seq_pool.p);
    c->launches += c->ncomp ? 2 : 0;
    if (ov) { MARK(c, "wait k_huf"); CK(c, cudaStreamWaitEvent(st, c->ev_join, 0)); }
    MARK(c, "k_plan2");  zsbk_plan2(st, frames, c->nf, blocks, work, fout, cnt, c->dst_cap, c->flags, pre_off); c->launches++;
    MARK(c, "k_rawrle"); zsbk_rawrle(st, c->n_rawrle, src, blocks, work, fout, (const uint32_t *)c->rawrle_list.p, cnt, c->d_dst); c->launches += c->n_rawrle ? 1 : 0;
    MARK(c, "k_exec2");  zsbk_exec2(st, c->n_exec2, src, frames, blocks, work, fout, (const uint32_t *)c->exec2_list.p, cnt, (const uint64_t *)c->seq_pool.p,
                                    (const uint8_t *)c->lit_pool.p, c->d_dst); c->launches += c->n_exec2 ? 1 : 0;
    if (c->n_exec && c->wave_ctas > 1) CK(c, cudaMemsetAsync(c->wave.p, 0, 32 * (size_t)c->n_exec + 4 * ((size_t)c->nb + 1), st));
    MARK(c, "k_exec");   zsbk_exec(st, c->n_exec, src, frames, blocks, work, fout, (const uint32_t *)c->exec_list.p, cnt, (const uint64_t *)c->seq_pool.p,
                                   (const uint8_t *)c->lit_pool.p, c->d_dst, c->is_sub, c->flags, c->wave_ctas > 1 ? c->wave.p : nullptr,
                                   c->wave_ctas > 1 ? (uint32_t *)((uint8_t *)c->wave.p + 32 * (size_t)c->n_exec) : nullptr, c->wave_ctas); c->launches += c->n_exec ? 1 : 0;
    if (c->n_link) {
        const uint8_t *lm = (const uint8_t *)c->link_meta.p;
        const void *lfr = lm, *lbl = lm + 16 * (size_t)c->n_link;
        uint32_t *tickets = (uint32_t *)(lm + 16 * (size_t)c->n_link + 8 * (size_t)c->n_link_blocks);
        CK(c, cudaMemsetAsync(tickets, 0, 4 * (size_t)c->n_link, st));         // (a relaunch of the same prepared batch starts from zero again)
        for (int which = 1; which <= 2; which++) {
            MARK(c, which == 1 ? "k_link_init" : "k_link_resolve");
            zsbk_link(st, c->n_link, c->n_link_blocks, src, blocks, work, fout, lbl, lfr, cnt, (const uint64_t *)c->seq_pool.p, (const uint8_t *)c->lit_pool.p,
                      (uint32_t *)c->link_ent.p, tickets, c->d_dst, c->n_sm, which); c->launches++;
        }
    }
    // pipelined path: the shard's output may leave as soon as it is written -- the checksums are computed from HBM while the
    // download runs (both only read the output)
    const bool early_down = c->down_stream && c->eager_d2h && c->h_dst;
    if (early_down) {
        if (c->trace) cudaEventRecord(c->ev_tr[2], st);
        CK(c, cudaEventRecord(c->ev_kdone, st)); CK(c, cudaStreamWaitEvent(c->down_stream, c->ev_kdone, 0));
        CK(c, cudaMemcpyAsync(c->h_dst, c->d_dst, c->eager_d2h, cudaMemcpyDeviceToHost, c->down_stream));
        CK(c, cudaEventRecord(c->ev_down, c->down_stream));
        if (c->trace) cudaEventRecord(c->ev_tr[3], c->down_stream);
    }
    MARK(c, "k_xxh");    zsbk_xxh(st, c->n_xxh, c->d_dst, fout, (const uint32_t *)c->xxh_list.p, cnt); c->launches += c->n_xxh ? 1 : 0;
    if (c->n_link && (c->flags & ZSB_VERIFY_CHECKSUM)) {
        MARK(c, "k_xxh_one"); zsbk_xxh_one(st, c->n_link, c->d_dst, fout, frames, c->link_meta.p, cnt); c->launches++;
    }
    if (c->debug_sync) { const bool pr = c->profile; c->profile = false; MARK(c, "the end of the batch"); c->profile = pr; c->dbg_prev = "the uploads"; }
    if (c->profile) { cudaEventRecord(c->ev[c->prof_slot][c->nk], st); c->prof_slot = (c->prof_slot + 1) % kProfRing; c->prof_count++; }
    if (c->trace && !early_down) cudaEventRecord(c->ev_tr[2], st);
    c->pin_valid = false;
    if (c->down_stream) {
        const size_t need = 64 + sizeof(ZsbFrameOut) * (size_t)c->nf;
        if (need > c->pin_cap) {
            if (c->pin) cudaFreeHost(c->pin);
            c->pin = nullptr; c->pin_cap = 0;
            if (cudaHostAlloc((void **)&c->pin, need + need / 4, cudaHostAllocDefault) == cudaSuccess) c->pin_cap = need + need / 4; else (void)cudaGetLastError();
        }
        void *pin_dev = nullptr;
        if (c->pin && cudaHostGetDevicePointer(&pin_dev, c->pin, 0) == cudaSuccess && pin_dev) {
            zsbk_publish(st, pin_dev, cnt, fout, c->nf);        // written by the kernel itself, not by a copy engine (see k_publish)
            c->pin_valid = true;
        } else (void)cudaGetLastError();
    }
    if (!early_down) {
        if (c->eager_d2h && c->h_dst) CK(c, cudaMemcpyAsync(c->h_dst, c->d_dst, c->eager_d2h, cudaMemcpyDeviceToHost, st));
        if (c->trace) cudaEventRecord(c->ev_tr[3], st);
    }
    CK(c, cudaGetLastError());
    c->launched = true;
    return ZSB_OK;
}

extern "C" int zsb_decode_finish(zsb_ctx *c, uint64_t *dst_off, uint64_t *dst_len, int32_t *status, uint32_t *xxh32,
                                 uint8_t *checksum_ok, uint64_t *dst_total) {
    if (!c || !c->launched) return ZSB_E_ARG;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    ZsbCounters hc;
    bool from_pin = false;
    for (int attempt = 0;; attempt++) {
        if (c->pin_valid) { CK(c, cudaStreamSynchronize(st)); memcpy(&hc, c->pin, sizeof hc); from_pin = !hc.overflow && !hc.refuse; }
        else {
            CK(c, cudaMemcpyAsync(&hc, c->counters.p, sizeof hc, cudaMemcpyDeviceToHost, st));
            CK(c, cudaStreamSynchronize(st));
        }
        if (!hc.overflow && !hc.refuse) break;
        if (attempt >= 3) { c->last_err = "scratch overflow persists"; return ZSB_E_NOMEM; }
        if (!hc.overflow) {
            // k_seqx executed a block where its frame did not end up (an earlier frame failed, or regenerates another size than it
            // declares): the whole batch again, with the sequence stage writing records (k_seq)
            c->use_seqx = false; c->seqx_state = 2;
            if (getenv("ZSB_DEBUG")) fprintf(stderr, "zsb: k_seqx refused, batch runs again\n");
            int rc = zsb_decode_launch(c);
            if (rc) return rc;
            continue;
        }
        // the entropy scratch was too small for this input: size it exactly and run the batch again
        c->lit_cap = (hc.lit_total + 64 + 15) & ~(uint64_t)15; c->seq_cap = hc.seq_total + 8;
        CK(c, c->lit_pool.ensure(c->lit_cap + kLitOverflow + 64));
        CK(c, c->seq_pool.ensure(8 * (c->seq_cap + 8)));
            int rc = zsb_decode_launch(c);
        if (rc) return rc;
    }
    if (c->profile) { const int r = (c->prof_slot + kProfRing - 1) % kProfRing; for (int i = 0; i < c->nk; i++) cudaEventElapsedTime(&c->kms[i], c->ev[r][i], c->ev[r][i + 1]); }
    std::vector<ZsbFrameOut> fo(c->nf);
    if (from_pin) { if (c->nf) memcpy(fo.data(), c->pin + 64, sizeof(ZsbFrameOut) * (size_t)c->nf); }
    else if (c->nf) CK(c, cudaMemcpyAsync(fo.data(), c->fout.p, sizeof(ZsbFrameOut) * c->nf, cudaMemcpyDeviceToHost, st));
    if (c->h_dst && hc.dst_total && hc.dst_total != c->eager_d2h) CK(c, cudaMemcpyAsync(c->h_dst, c->d_dst, hc.dst_total, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    if (c->down_stream && c->eager_d2h && c->h_dst) CK(c, cudaEventSynchronize(c->ev_down));
    c->h_err_a.assign(c->nf, 0); c->h_err_b.assign(c->nf, 0);
    for (uint32_t f = 0; f < c->nf; f++) {
        const bool ok = fo[f].status == ZSB_OK;
        if (!ok) { c->h_err_a[f] = fo[f].err_a; c->h_err_b[f] = fo[f].err_b; }
        if (dst_off) dst_off[f] = fo[f].dst_off;
        if (dst_len) dst_len[f] = ok ? fo[f].dst_len : 0;
        if (status) status[f] = fo[f].status;
        const bool hashed = ok && (c->flags & ZSB_VERIFY_CHECKSUM) && c->h_frames[f].kind == 0 && c->h_frames[f].has_checksum;
        if (xxh32) xxh32[f] = hashed ? (uint32_t)fo[f].xxh64 : 0;
        if (checksum_ok) checksum_ok[f] = hashed ? ((uint32_t)fo[f].xxh64 == c->h_frames[f].stored_checksum) : 0;
    }
    if (dst_total) *dst_total = hc.dst_total;
    return ZSB_OK;
}

// Host buffers, many frames: the batch is cut into shards by frame and every shard runs on its own stream and scratch --
// upload of shard k+1, kernels of shard k and download of shard k-1 overlap (PCIe is ~3/4 of the host-to-host time of a
// batch).  Early output placement needs every frame's size before it is decoded: shards whose frames all declare
// Frame_Content_Size are placed at dispatch and kept only if every frame decoded to exactly that size (otherwise: return 1,
// the plain path runs and reports as usual); shards from the first undeclared size on are placed late (struct Pipe).
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const size_t kPipeMinFrames = 512;
static const int kPipeShardsMax = 32;
// Page-locked (cudaHostAlloc / cudaHostRegister) host memory?  Copies from and to pageable memory block the calling thread until
// they are done, which would run the shards of the pipelined path one after the other (measured: 32 ms instead of 6 ms for a
// 40 MB batch): pageable buffers take the plain path.
static bool host_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// bytes frame f contributes to the output
static inline uint64_t frame_out_bytes(const zsb_frame &f, const zsb_block *blocks, uint32_t flags) {
    return f.kind == 1 ? ((flags & ZSB_PRINT_SKIPPABLE) ? blocks[f.first_block].size : 0) : f.content_size;
}
// Shard plan.  The download of the whole output (PCIe, ~55 GB/s) is the longest leg, so the goal is to start it as early as
// possible and never let it wait: a shard's output can leave only after its upload plus ~2-3 ms of kernels (every stage is a
// per-frame dependent chain, so the latency does not shrink with the shard).  Hence a small first shard and sizes that grow
// about as fast as the download falls behind the upload (x2); the first shards also run in low-latency mode (ctx.low_latency)
// when the frames are large enough for a CTA each.  Measured on C2: 8 equal shards 14.1 ms, this plan 12.3 ms.
// ZSB_PIPE_WEIGHTS="w0,w1,..." and ZSB_PIPE_FAST_SHARDS=k override the plan (experiments).
static const double kPlanFast[] = {1, 2, 4, 8, 8, 10, 14, 17}, kPlanPlain[] = {1, 2, 4, 8, 16, 33};
static const int kPlanFastShards = 5;
static int pipe_plan(bool big_frames, double *wts, int &n_fast) {
    int ns = 0;
    n_fast = big_frames ? kPlanFastShards : 0;
    if (const char *e = getenv("ZSB_PIPE_WEIGHTS")) { while (*e && ns < kPipeShardsMax) { char *q; double v = strtod(e, &q); if (q == e) break; wts[ns++] = v > 0 ? v : 1; e = *q == ',' ? q + 1 : q; } }
    if (!ns) {
        const double *pl = big_frames ? kPlanFast : kPlanPlain;
        ns = big_frames ? (int)(sizeof kPlanFast / sizeof *kPlanFast) : (int)(sizeof kPlanPlain / sizeof *kPlanPlain);
        for (int k = 0; k < ns; k++) wts[k] = pl[k];
    }
    if (const char *e = getenv("ZSB_PIPE_FAST_SHARDS")) n_fast = atoi(e);
    return ns;
}

// The shards of one pipelined call: dispatch() enqueues upload + kernels + download of a frame range on the next child context,
// collect() waits for all of them in order and merges the per-frame results.
struct Pipe {
    zsb_ctx *c; const uint8_t *src; uint8_t *dst; size_t dst_cap; uint32_t flags;
    struct Sh { size_t f0 = 0, f1 = 0; zsb_frame *fr = nullptr; zsb_block *bl = nullptr; size_t nb = 0; uint64_t so = 0, sl = 0, doff = 0, dexp = 0; bool on = false, late = false; double host_ms = 0; } sh[kPipeShardsMax];
    int n = 0;
    uint64_t doff = 0;
    // Shards whose frames all declare their size are placed when they are dispatched and downloaded right behind their kernels.
    // From the first shard with a frame of unknown size on, placement is late: the shard decodes into its own device buffer
    // (capacity: 128 KiB per compressed block) and collect() sends it to the host once the sizes before it are known.
    bool late = false;
    double t0 = 0, t_enq = 0;
    Pipe(zsb_ctx *ctx, const uint8_t *s, uint8_t *d, size_t cap, uint32_t fl) : c(ctx), src(s), dst(d), dst_cap(cap), flags(fl) { t0 = now_ms(); }
    ~Pipe() { for (int k = 0; k < n; k++) { zsb_free(sh[k].fr); zsb_free(sh[k].bl); } }
    // frames [f0, f1) of the (possibly still growing) descriptor arrays become the next shard; false = give the pipeline up
    bool dispatch(const zsb_frame *frames, size_t nf, const zsb_block *blocks, size_t nb, size_t f0, size_t f1, bool low_latency) {
        if (n >= kPipeShardsMax) return false;
        const int k = n++;
        Sh &S = sh[k];
        S.f0 = f0; S.f1 = f1; S.doff = doff;
        if (f0 == f1) return true;
        for (size_t f = f0; f < f1 && !late; f++) late = frames[f].kind == 0 && !frames[f].has_content_size;
        S.late = late;
        while (c->subs.size() <= (size_t)k) {
            zsb_ctx *sub = nullptr;
            if (zsb_ctx_create(&sub, c->device) != ZSB_OK) return false;
            sub->is_sub = true;
            c->subs.push_back(sub);
        }
        if (zsb_shard_extract(frames, nf, blocks, nb, f0, f1, &S.fr, &S.bl, &S.nb, &S.so, &S.sl) != ZSB_OK) return false;
        zsb_ctx *sub = c->subs[k];
        uint8_t *target = nullptr; uint32_t fl = flags;
        if (!S.late) {
            for (size_t f = f0; f < f1; f++) S.dexp += frame_out_bytes(frames[f], blocks, flags);
            if (S.doff + S.dexp > dst_cap) return false;
            doff += S.dexp;
            sub->eager_d2h = S.dexp;
            target = dst + S.doff;
        } else {
            for (size_t f = f0; f < f1; f++) {                      // capacity, not a promise: what the blocks can regenerate at most
                if (frames[f].kind == 1) { S.dexp += frame_out_bytes(frames[f], blocks, flags); continue; }
                for (uint32_t b = 0; b < frames[f].n_blocks; b++) {
                    const zsb_block &B = blocks[frames[f].first_block + b];
                    S.dexp += B.type == ZSB_BT_COMPRESSED ? ZSB_BLOCK_MAX : B.size;
                }
            }
            cudaSetDevice(c->device);
            if (sub->dst.ensure(S.dexp + 64) != cudaSuccess) { (void)cudaGetLastError(); return false; }
            sub->eager_d2h = 0;
            target = (uint8_t *)sub->dst.p; fl |= ZSB_DST_ON_DEVICE;
        }
        sub->up_stream = c->own_stream; sub->down_stream = c->aux_stream;
        sub->low_latency = low_latency;
        bool first_on = true; for (int j = 0; j < k; j++) first_on = first_on && !sh[j].on;
        if (c->trace && first_on) cudaEventRecord(c->ev_tr[0], sub->up_stream);
        if (c->trace) S.host_ms = now_ms() - t0;
        if (zsb_decode_prepare(sub, src + S.so, S.sl, S.fr, f1 - f0, S.bl, S.nb, target, S.dexp, fl) != ZSB_OK ||
            zsb_decode_launch(sub) != ZSB_OK) return false;
        S.on = true;
        return true;
    }
    // returns ZSB_OK, ZSB_E_CUDA, or 1 when the result must not be used (a frame failed or disagreed with its declared size)
    int collect(uint64_t *dst_off, uint64_t *dst_len, int32_t *status, uint32_t *xxh32, uint8_t *checksum_ok, uint64_t *dst_total) {
        int rc = ZSB_OK; bool bad = false, any_late = false;
        uint64_t total = 0;                                  // == where the next late shard goes
        size_t nf_all = 0; for (int k = 0; k < n; k++) if (sh[k].f1 > nf_all) nf_all = sh[k].f1;
        c->h_err_a.assign(nf_all, 0); c->h_err_b.assign(nf_all, 0);
        t_enq = now_ms();
        for (int k = 0; k < n; k++) {
            Sh &S = sh[k];
            if (!S.on) continue;
            uint64_t t = 0;
            const int r = zsb_decode_finish(c->subs[k], dst_off ? dst_off + S.f0 : nullptr, dst_len ? dst_len + S.f0 : nullptr, status ? status + S.f0 : nullptr,
                                            xxh32 ? xxh32 + S.f0 : nullptr, checksum_ok ? checksum_ok + S.f0 : nullptr, &t);
            if (r != ZSB_OK) { rc = r; c->last_err = c->subs[k]->last_err; bad = true; continue; }
            if (c->h_err_a.size() >= S.f1) zsb_decode_errors(c->subs[k], c->h_err_a.data() + S.f0, c->h_err_b.data() + S.f0, S.f1 - S.f0);
            if (!S.late) {
                if (t != S.dexp) bad = true;
                if (dst_off) for (size_t f = S.f0; f < S.f1; f++) dst_off[f] += S.doff;
                total = S.doff + t;
            } else if (!bad) {
                if (total + t > dst_cap) { bad = true; continue; }         // the plain path reports which frames do not fit
                if (t && cudaMemcpyAsync(dst + total, c->subs[k]->d_dst, t, cudaMemcpyDeviceToHost, c->aux_stream) != cudaSuccess) {
                    c->last_err = "cudaMemcpyAsync (late shard)"; (void)cudaGetLastError(); rc = ZSB_E_CUDA; bad = true; continue;
                }
                any_late = true;
                if (c->trace) cudaEventRecord(c->subs[k]->ev_tr[3], c->aux_stream);
                if (dst_off) for (size_t f = S.f0; f < S.f1; f++) dst_off[f] += total;
                total += t;
            }
        }
        if (any_late && cudaStreamSynchronize(c->aux_stream) != cudaSuccess) { c->last_err = "cudaStreamSynchronize (late shards)"; (void)cudaGetLastError(); rc = ZSB_E_CUDA; }
        if (c->trace) {
            fprintf(stderr, "[zsb pipe] host: all shards enqueued at %.2f ms, finished at %.2f ms\n", t_enq - t0, now_ms() - t0);
            for (int k = 0; k < n; k++) if (sh[k].on) {
                float a = 0, b = 0, d = 0;
                cudaEventElapsedTime(&a, c->ev_tr[0], c->subs[k]->ev_tr[1]); cudaEventElapsedTime(&b, c->ev_tr[0], c->subs[k]->ev_tr[2]); cudaEventElapsedTime(&d, c->ev_tr[0], c->subs[k]->ev_tr[3]);
                fprintf(stderr, "[zsb pipe] shard %d: enqueue starts %.2f (host) | uploads done %.2f, output written %.2f, download done %.2f ms (device)\n", k, sh[k].host_ms, a, b, d);
            }
            (void)cudaGetLastError();
        }
        if (rc == ZSB_E_CUDA) return rc;
        if (bad) return 1;
        if (dst_total) *dst_total = total;
        return ZSB_OK;
    }
};

static int decode_pipelined(zsb_ctx *c, const uint8_t *src, size_t n, const zsb_frame *frames, size_t nf, const zsb_block *blocks, size_t nb,
                            uint8_t *dst, size_t dst_cap, uint64_t *dst_off, uint64_t *dst_len, int32_t *status, uint32_t *xxh32,
                            uint8_t *checksum_ok, uint64_t *dst_total, uint32_t flags) {
    (void)n;
    // expected output: declared sizes, 2.4 x the compressed size where a frame does not declare one (text at level 3)
    auto est = [&](const zsb_frame &fr) { return fr.kind == 0 && !fr.has_content_size ? (uint64_t)(2.4 * (double)fr.src_len) : frame_out_bytes(fr, blocks, flags); };
    uint64_t expect_total = 0; bool all_sized = true;
    for (size_t f = 0; f < nf; f++) {
        if (frames[f].status != ZSB_OK) return 1;
        if (frames[f].kind == 0 && !frames[f].has_content_size) all_sized = false;
        expect_total += est(frames[f]);
    }
    if ((all_sized && expect_total > dst_cap) || expect_total < (32u << 20)) return 1;
    double wts[kPipeShardsMax]; int n_fast = 0;
    const int ns = pipe_plan(expect_total / nf >= (64u << 10), wts, n_fast);
    double wsum = 0; for (int k = 0; k < ns; k++) wsum += wts[k];
    Pipe P(c, src, dst, dst_cap, flags);
    // boundaries: cumulative output bytes cut at the cumulative weights
    size_t f = 0; double acc = 0, cum = 0;
    bool ok = true;
    for (int k = 0; k < ns && ok; k++) {
        cum += wts[k];
        const double lim = (double)expect_total * (cum / wsum);
        const size_t f0 = f;
        while (f < nf && (k == ns - 1 || acc < lim)) { acc += (double)est(frames[f]); f++; }
        ok = P.dispatch(frames, nf, blocks, nb, f0, f, k < n_fast);
    }
    const int rc = P.collect(dst_off, dst_len, status, xxh32, checksum_ok, dst_total);
    if (rc == ZSB_E_CUDA) return rc;
    return (ok && rc == ZSB_OK) ? ZSB_OK : 1;
}

extern "C" int zsb_decode(zsb_ctx *c, const uint8_t *src, size_t n, const zsb_frame *frames, size_t nf, const zsb_block *blocks, size_t nb,
                          uint8_t *dst, size_t dst_cap, uint64_t *dst_off, uint64_t *dst_len, int32_t *status, uint32_t *xxh32,
                          uint8_t *checksum_ok, uint64_t *dst_total, uint32_t flags) {
    if (c && !c->is_sub && !(flags & (ZSB_SRC_ON_DEVICE | ZSB_DST_ON_DEVICE)) && nf >= kPipeMinFrames && src && dst && frames && blocks &&
        host_pinned(src) && host_pinned(dst)) {
        const int prc = decode_pipelined(c, src, n, frames, nf, blocks, nb, dst, dst_cap, dst_off, dst_len, status, xxh32, checksum_ok, dst_total, flags);
        if (prc != 1) return prc;
    }
    int rc = zsb_decode_prepare(c, src, n, frames, nf, blocks, nb, dst, dst_cap, flags);
    if (rc) return rc;
    rc = zsb_decode_launch(c);
    if (rc) return rc;
    return zsb_decode_finish(c, dst_off, dst_len, status, xxh32, checksum_ok, dst_total);
}

extern "C" int zsb_scan_device(zsb_ctx *c, const uint8_t *d_src, size_t n, uint32_t flags, uint64_t max_window,
                               zsb_frame **frames_out, size_t *n_frames, zsb_block **blocks_out, size_t *n_blocks, uint64_t *err_a, uint64_t *err_b);

// zsb_scan + zsb_decode in one call on host buffers, with the host walk overlapped: the walk stops at every shard boundary
// (a fraction of the compressed bytes) and the frames found so far start uploading and decoding while the rest of the buffer
// is still being walked.  Results are those of zsb_scan followed by zsb_decode.  Anything that keeps the pipelined path from
// applying (a malformed frame, an output that does not fit, a frame that fails or disagrees with its declared size in a
// shard that was placed early) ends in one plain batch over the completed scan.
extern "C" int zsb_scan_decode(zsb_ctx *c, const uint8_t *src, size_t n, uint8_t *dst, size_t dst_cap, uint32_t flags, uint64_t max_window,
                               zsb_frame **frames_out, size_t *n_frames, zsb_block **blocks_out, size_t *n_blocks,
                               zsb_result **results_out, uint64_t *dst_total, uint64_t *err_a, uint64_t *err_b) {
    if (!c || c->is_sub || (!src && n) || !dst || !frames_out || !n_frames || !blocks_out || !n_blocks || !results_out) return ZSB_E_ARG;
    *frames_out = nullptr; *blocks_out = nullptr; *results_out = nullptr; *n_frames = 0; *n_blocks = 0;
    if (flags & ZSB_SRC_ON_DEVICE) {
        // the compressed bytes are resident in HBM: the walk runs there too (zsb_scan_device), then one batch; dst is a host or (ZSB_DST_ON_DEVICE) a device buffer
        uint64_t ea = 0, eb = 0, total = 0;
        const int scan_rc = zsb_scan_device(c, src, n, flags, max_window, frames_out, n_frames, blocks_out, n_blocks, &ea, &eb);
        if (!*frames_out) return scan_rc;
        const size_t nf = *n_frames;
        std::vector<uint64_t> off(nf + 1), len(nf + 1); std::vector<int32_t> st(nf + 1); std::vector<uint32_t> xh(nf + 1); std::vector<uint8_t> ck(nf + 1);
        zsb_result *res = (zsb_result *)malloc(sizeof(zsb_result) * (nf + 1));
        int rc = res ? zsb_decode(c, src, n, *frames_out, nf, *blocks_out, *n_blocks, dst, dst_cap, off.data(), len.data(), st.data(), xh.data(), ck.data(), &total, flags) : ZSB_E_NOMEM;
        if (rc) { free(res); zsb_free(*frames_out); zsb_free(*blocks_out); *frames_out = nullptr; *blocks_out = nullptr; *n_frames = 0; *n_blocks = 0; return rc; }
        for (size_t f = 0; f < nf; f++) {
            res[f].dst_off = off[f]; res[f].dst_len = len[f]; res[f].status = st[f]; res[f].xxh32 = xh[f]; res[f].checksum_ok = ck[f]; memset(res[f].pad, 0, sizeof res[f].pad);
            res[f].err_a = f < c->h_err_a.size() ? c->h_err_a[f] : 0; res[f].err_b = f < c->h_err_b.size() ? c->h_err_b[f] : 0;
        }
        *results_out = res;
        if (dst_total) *dst_total = total;
        if (err_a) *err_a = ea;
        if (err_b) *err_b = eb;
        return scan_rc;
    }
    flags &= ~ZSB_DST_ON_DEVICE;
    ZsbScanner sc(src, n, flags, max_window);
    Pipe P(c, src, dst, dst_cap, flags);
    bool streamed = n >= (16u << 20) && host_pinned(src) && host_pinned(dst);     // worth cutting up at all, and the copies asynchronous
    const bool tried = streamed;
    if (streamed) {
        double wts[kPipeShardsMax]; int n_fast = 0;
        const int ns = pipe_plan(true, wts, n_fast);
        double wsum = 0, cum = 0; for (int k = 0; k < ns; k++) wsum += wts[k];
        for (int k = 0; k < ns && streamed && !sc.done; k++) {
            cum += wts[k];
            const size_t lim = k == ns - 1 ? n : (size_t)((double)n * (cum / wsum));
            const size_t f0 = sc.frames.size();
            while (!sc.done && sc.pos < lim) if (!sc.next()) break;
            if (sc.code != ZSB_OK) { streamed = false; break; }          // the walk ended on a malformed frame
            // the first shard swallowed the whole buffer (a single large frame, C3): nothing to overlap, and the context's own path executes a
            // frame of many blocks with 1 024-thread CTAs and hashes it beside the execution, which a shard's context does not
            if (k == 0 && (sc.done || sc.pos >= n)) { streamed = false; break; }
            const size_t f1 = sc.frames.size();
            uint64_t out = 0;
            for (size_t f = f0; f < f1; f++)
                out += sc.frames[f].kind == 0 && !sc.frames[f].has_content_size ? (uint64_t)(2.4 * (double)sc.frames[f].src_len) : frame_out_bytes(sc.frames[f], sc.blocks.data(), flags);
            const bool big = f1 > f0 && out / (f1 - f0) >= (64u << 10);
            if (!P.dispatch(sc.frames.data(), f1, sc.blocks.data(), sc.blocks.size(), f0, f1, big && k < n_fast)) streamed = false;
        }
    }
    while (sc.next()) {}                                                 // whatever is left (nothing, unless the pipeline was given up)
    const size_t nf = sc.frames.size();
    std::vector<uint64_t> off(nf + 1), len(nf + 1); std::vector<int32_t> st(nf + 1); std::vector<uint32_t> xh(nf + 1); std::vector<uint8_t> ck(nf + 1);
    uint64_t total = 0;
    int rc = P.collect(off.data(), len.data(), st.data(), xh.data(), ck.data(), &total);     // also drains shards of a pipeline given up
    if (rc == ZSB_E_CUDA) return rc;
    if (!streamed || rc != ZSB_OK) {
        if (!tried) rc = zsb_decode(c, src, n, sc.frames.data(), nf, sc.blocks.data(), sc.blocks.size(), dst, dst_cap, off.data(), len.data(), st.data(), xh.data(), ck.data(), &total, flags);
        else {      // the pipeline was tried and given up: one plain batch (zsb_decode would try its own pipeline first)
            rc = zsb_decode_prepare(c, src, n, sc.frames.data(), nf, sc.blocks.data(), sc.blocks.size(), dst, dst_cap, flags);
            if (!rc) rc = zsb_decode_launch(c);
            if (!rc) rc = zsb_decode_finish(c, off.data(), len.data(), st.data(), xh.data(), ck.data(), &total);
        }
        if (rc) return rc;
    }
    zsb_result *res = (zsb_result *)malloc(sizeof(zsb_result) * (nf + 1));
    if (!res) return ZSB_E_NOMEM;
    for (size_t f = 0; f < nf; f++) {
        res[f].dst_off = off[f]; res[f].dst_len = len[f]; res[f].status = st[f]; res[f].xxh32 = xh[f]; res[f].checksum_ok = ck[f]; memset(res[f].pad, 0, sizeof res[f].pad);
        res[f].err_a = f < c->h_err_a.size() ? c->h_err_a[f] : 0; res[f].err_b = f < c->h_err_b.size() ? c->h_err_b[f] : 0;
    }
    rc = sc.release(frames_out, n_frames, blocks_out, n_blocks);
    if (rc) { free(res); return rc; }
    *results_out = res;
    if (dst_total) *dst_total = total;
    if (err_a) *err_a = sc.err_a;
    if (err_b) *err_b = sc.err_b;
    return sc.code;
}

// ======================================================================================= zsb_scan_device
// zsb_scan for a buffer that is resident in HBM: the kernels of zsb_dscan.cu find the frames and blocks, the descriptors come back to the host.
// Same arrays, status and payload as zsb_scan on the same bytes (tests/test_gpu_parity.py compares them byte for byte).
extern "C" int zsb_scan_device(zsb_ctx *c, const uint8_t *d_src, size_t n, uint32_t flags, uint64_t max_window,
                               zsb_frame **frames_out, size_t *n_frames, zsb_block **blocks_out, size_t *n_blocks, uint64_t *err_a, uint64_t *err_b) {
    if (!c || !frames_out || !n_frames || !blocks_out || !n_blocks || (!d_src && n)) return ZSB_E_ARG;
    *frames_out = nullptr; *blocks_out = nullptr; *n_frames = 0; *n_blocks = 0;
    if (err_a) *err_a = 0;
    if (err_b) *err_b = 0;
    flags &= (ZSB_REFERENCE_QUIRKS | ZSB_STRICT_DICT);
    if (!max_window) max_window = ZSB_MAX_WINDOW_DEFAULT;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint64_t kMaxCand = 1ull << 24;
    std::vector<zsb_frame> frames; std::vector<zsb_block> blocks;
    int code = ZSB_OK; uint64_t ea = 0, eb = 0;
    uint64_t tail_off = 0;            // where the chain of well-formed frames ends
    bool chain_failed = false;        // ... on a frame that fails (it is the last entry of `frames`)
    uint64_t fail_emitted = 0;

    auto host_walk = [&](uint64_t from) -> int {      // the host walker on a copy of src[from, n): frames from `from` on replace what follows
        std::vector<uint8_t> h(n - from + 1);
        CK(c, cudaMemcpyAsync(h.data(), d_src + from, n - from, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        ZsbScanner sc(h.data(), n - from, flags, max_window);
        const size_t f0 = frames.size(), b0 = blocks.size();
        while (sc.next()) {}
        for (zsb_frame f : sc.frames) { f.src_off += from; f.first_block += (uint32_t)b0; frames.push_back(f); }
        for (zsb_block b : sc.blocks) { b.src_off += from; b.frame += (uint32_t)f0; blocks.push_back(b); }
        code = sc.code; ea = sc.err_a; eb = sc.err_b;
        return ZSB_OK;
    };

    if (n >= 4) {
        // one pass finds the candidates when there are no more than a guess allows (one per 2 KiB of input, at least 65 536); a buffer with more
        // of them is searched again with room for all
        uint64_t pos_cap = n / 2048 > 65536 ? n / 2048 : 65536;
        CK(c, c->dscan_pos.ensure(256 + 8 * pos_cap));
        unsigned long long *d_count = (unsigned long long *)c->dscan_pos.p;   // the first 256 bytes of the position list
        uint64_t *d_pos = (uint64_t *)((uint8_t *)c->dscan_pos.p + 256);
        unsigned long long count = 0;
        CK(c, cudaMemsetAsync(d_count, 0, 8, st));
        zsbk_dscan_find(st, d_src, n, d_count, d_pos, pos_cap, c->n_sm);
        CK(c, cudaMemcpyAsync(&count, d_count, 8, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        if (count > pos_cap && count <= kMaxCand) {
            pos_cap = count;
            CK(c, c->dscan_pos.ensure(256 + 8 * pos_cap));
            d_count = (unsigned long long *)c->dscan_pos.p; d_pos = (uint64_t *)((uint8_t *)c->dscan_pos.p + 256);
            CK(c, cudaMemsetAsync(d_count, 0, 8, st));
            zsbk_dscan_find(st, d_src, n, d_count, d_pos, pos_cap, c->n_sm);
        }
        if (count > kMaxCand) {                                               // a buffer that is mostly magic numbers: not worth the tables
            const int rc = host_walk(0);
            if (rc) return rc;
            goto done;
        }
        if (count) {
            const uint32_t ncand = (uint32_t)count, n1 = ncand + 1;
            uint32_t levels = 1; while ((1ull << levels) < n1) levels++;
            uint32_t mask = 1; while (mask < 2 * ncand + 2) mask <<= 1; mask -= 1;
            size_t o = 0;
            auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
            const size_t o_cand = take(sizeof(ZsbDscanCand) * (size_t)ncand), o_keys = take(8ull * (mask + 1ull)), o_vals = take(4ull * (mask + 1ull)),
                         o_jump = take(4ull * (levels + 1ull) * n1), o_da = take(4ull * n1), o_db = take(4ull * n1), o_head = take(4), o_tail = take(8 * ZSB_DSCAN_TAIL_WORDS),
                         o_order = take(4ull * ncand), o_fr = take(sizeof(zsb_frame) * (size_t)ncand), o_fb = take(4ull * ncand);
            CK(c, c->dscan.ensure(o));
            uint8_t *A = (uint8_t *)c->dscan.p;
            ZsbDscanCand *d_cand = (ZsbDscanCand *)(A + o_cand);
            unsigned long long *d_keys = (unsigned long long *)(A + o_keys); uint32_t *d_vals = (uint32_t *)(A + o_vals), *d_jump = (uint32_t *)(A + o_jump);
            uint32_t *d_da = (uint32_t *)(A + o_da), *d_db = (uint32_t *)(A + o_db), *d_head = (uint32_t *)(A + o_head), *d_order = (uint32_t *)(A + o_order), *d_fb = (uint32_t *)(A + o_fb);
            uint64_t *d_tail = (uint64_t *)(A + o_tail); zsb_frame *d_fr = (zsb_frame *)(A + o_fr);
            CK(c, cudaMemsetAsync(d_keys, 0, 8ull * (mask + 1ull), st));
            zsbk_dscan_parse(st, d_src, n, flags, max_window, d_pos, ncand, d_cand, d_keys, d_vals, mask);
            zsbk_dscan_link(st, d_cand, ncand, n, d_keys, d_vals, mask, d_jump, levels, d_da, d_db, d_head);
            uint32_t head = 0xFFFFFFFFu, nfr = 0;
            CK(c, cudaMemcpyAsync(&head, d_head, 4, cudaMemcpyDeviceToHost, st));
            CK(c, cudaStreamSynchronize(st));
            if (head != 0xFFFFFFFFu) {
                CK(c, cudaMemcpyAsync(&nfr, ((levels & 1) ? d_db : d_da) + head, 4, cudaMemcpyDeviceToHost, st));
                CK(c, cudaStreamSynchronize(st));
            }
            if (nfr) {
                uint64_t tail[ZSB_DSCAN_TAIL_WORDS] = {};
                frames.resize(nfr);
                zsbk_dscan_order(st, d_jump, levels, ncand, head, nfr, d_cand, d_order, d_fr, d_tail);
                CK(c, cudaMemcpyAsync(frames.data(), d_fr, sizeof(zsb_frame) * (size_t)nfr, cudaMemcpyDeviceToHost, st));
                CK(c, cudaMemcpyAsync(tail, d_tail, sizeof tail, cudaMemcpyDeviceToHost, st));
                CK(c, cudaStreamSynchronize(st));
                std::vector<uint32_t> fb(nfr);
                uint64_t nb = 0;
                for (uint32_t f = 0; f < nfr; f++) { fb[f] = (uint32_t)nb; frames[f].first_block = (uint32_t)nb; nb += frames[f].n_blocks; }
                if (nb > 0xFFFFFFFFull) { c->last_err = "zsb_scan_device: more than 2^32 blocks"; return ZSB_E_ARG; }
                blocks.resize(nb);
                if (nb) {
                    CK(c, c->dscan_blocks.ensure(sizeof(zsb_block) * (size_t)nb));
                    CK(c, cudaMemcpyAsync(d_fb, fb.data(), 4ull * nfr, cudaMemcpyHostToDevice, st));
                    zsbk_dscan_emit(st, d_src, n, flags, max_window, d_fr, d_fb, nfr, (zsb_block *)c->dscan_blocks.p);
                    CK(c, cudaMemcpyAsync(blocks.data(), c->dscan_blocks.p, sizeof(zsb_block) * (size_t)nb, cudaMemcpyDeviceToHost, st));
                }
                CK(c, cudaGetLastError());
                CK(c, cudaStreamSynchronize(st));
                if (tail[0]) tail_off = tail[1];
                else { chain_failed = true; tail_off = frames[nfr - 1].src_off; code = frames[nfr - 1].status; ea = tail[3]; eb = tail[4]; fail_emitted = tail[5]; }
            }
        }
    }
    if (chain_failed) {
        // ZSB_REFERENCE_QUIRKS: a section error of a block read before the point of failure comes first (eager_sections, zsb_scan.cpp): the host walker
        // decides, on the bytes from the failing frame on
        if ((flags & ZSB_REFERENCE_QUIRKS) && frames.back().kind == 0 && fail_emitted) {
            frames.pop_back();
            const int rc = host_walk(tail_off);
            if (rc) return rc;
        }
    } else if (tail_off < n) {
        // no frame starts here: fewer than four bytes, or four bytes that are no magic number (Frame::parse frame.rs:61-77)
        zsb_frame f; memset(&f, 0, sizeof f);
        f.src_off = tail_off; f.src_len = n - tail_off; f.first_block = (uint32_t)blocks.size();
        if (n - tail_off < 4) { code = ZSB_E_NOT_ENOUGH_BYTES; ea = 4; eb = n - tail_off; }
        else {
            uint8_t m[4];
            CK(c, cudaMemcpyAsync(m, d_src + tail_off, 4, cudaMemcpyDeviceToHost, st));
            CK(c, cudaStreamSynchronize(st));
            f.magic = (uint32_t)m[0] | (uint32_t)m[1] << 8 | (uint32_t)m[2] << 16 | (uint32_t)m[3] << 24;
            if (zsb_is_frame_magic(f.magic)) { c->last_err = "zsb_scan_device: a frame start was not found as a candidate"; return ZSB_E_CUDA; }
            code = ZSB_E_UNRECOGNIZED_MAGIC; ea = f.magic; eb = 0;
        }
        f.status = code;
        frames.push_back(f);
    }
done:
    {
        zsb_frame *fo = (zsb_frame *)malloc(sizeof(zsb_frame) * (frames.size() + 1));
        zsb_block *bo = (zsb_block *)malloc(sizeof(zsb_block) * (blocks.size() + 1));
        if (!fo || !bo) { free(fo); free(bo); return ZSB_E_NOMEM; }
        if (!frames.empty()) memcpy(fo, frames.data(), sizeof(zsb_frame) * frames.size());
        if (!blocks.empty()) memcpy(bo, blocks.data(), sizeof(zsb_block) * blocks.size());
        *frames_out = fo; *n_frames = frames.size(); *blocks_out = bo; *n_blocks = blocks.size();
    }
    if (err_a) *err_a = ea;
    if (err_b) *err_b = eb;
    return code;
}

// src/main.rs:42-58 : all-or-nothing whole-buffer decode
extern "C" int zsb_decompress(zsb_ctx *c, const uint8_t *src, size_t n, uint32_t flags, uint8_t **out, size_t *out_len,
                              uint64_t *err_a, uint64_t *err_b) {
    if (!c || !out || !out_len) return ZSB_E_ARG;
    *out = nullptr; *out_len = 0;
    flags &= ~(ZSB_SRC_ON_DEVICE | ZSB_DST_ON_DEVICE);
    zsb_frame *frames = nullptr; zsb_block *blocks = nullptr; size_t nf = 0, nb = 0;
    const int scan_rc = zsb_scan(src, n, flags, 0, &frames, &nf, &blocks, &nb, err_a, err_b);
    // src/main.rs:43-53 decodes every frame as the iterator yields it: an error in decoding frame j precedes the parse error of a later frame k,
    // so the frames in front of a malformed one are decoded first and only then is the walk's error reported
    const size_t nf_ok = scan_rc ? (nf ? nf - 1 : 0) : nf;
    // capacity: content sizes where declared (never with ZSB_REFERENCE_QUIRKS: the reference does not compare the decoded length with
    // Frame_Content_Size, a frame may be longer than it says); otherwise blocks regenerate at most 128 KiB each
    uint64_t cap = 0;
    for (size_t f = 0; f < nf_ok; f++) {
        if (frames[f].kind == 1) { cap += blocks[frames[f].first_block].size; continue; }
        if (frames[f].has_content_size && !(flags & ZSB_REFERENCE_QUIRKS)) { cap += frames[f].content_size; continue; }
        for (uint32_t k = 0; k < frames[f].n_blocks; k++) {
            const zsb_block &b = blocks[frames[f].first_block + k];
            cap += b.type == ZSB_BT_COMPRESSED ? ZSB_BLOCK_MAX : b.size;
        }
    }
    uint8_t *dst = (uint8_t *)malloc(cap ? cap : 1);
    std::vector<int32_t> status(nf_ok + 1);
    std::vector<uint64_t> off(nf_ok + 1), len(nf_ok + 1);
    uint64_t total = 0;
    if (!dst) { zsb_free(frames); zsb_free(blocks); return ZSB_E_NOMEM; }
    size_t nb_ok = nb;
    if (scan_rc && nf) nb_ok = frames[nf - 1].first_block;              // (a failed frame has no blocks; its first_block is the count so far)
    int rc = zsb_decode(c, src, n, frames, nf_ok, blocks, nb_ok, dst, cap, off.data(), len.data(), status.data(), nullptr, nullptr, &total, flags);
    if (!rc) for (size_t f = 0; f < nf_ok && !rc; f++) {
        rc = status[f];                                                  // main.rs:51 first error aborts, no partial output
        if (rc) { uint32_t a = 0, b = 0; zsb_decode_errors(c, &a, &b, 0); if (f < c->h_err_a.size()) { a = c->h_err_a[f]; b = c->h_err_b[f]; } if (err_a) *err_a = a; if (err_b) *err_b = b; }
    }
    if (!rc) rc = scan_rc;
    zsb_free(frames); zsb_free(blocks);
    if (rc) { free(dst); return rc; }
    *out = dst; *out_len = total;
    return ZSB_OK;
}

// ======================================================================================= stage-level entry points
namespace {
int stage_sync(zsb_ctx *c) { CK(c, cudaGetLastError()); CK(c, cudaStreamSynchronize(c->stream)); return ZSB_OK; }
}
static int fse_stage(zsb_ctx *c, const uint8_t *desc, size_t n, int max_symbols, const int16_t *dist_in, size_t nd_in, int al_in,
                     uint8_t *al, uint16_t *table, size_t *consumed, int16_t *dist, size_t *n_dist) {
    CK(c, cudaSetDevice(c->device));
    const size_t off_res = 0, off_cells = 64, off_dist = off_cells + 512 * 4, off_in = off_dist + 256 * 2, total = off_in + (desc ? n : nd_in * 2) + 64;
    CK(c, c->stage.ensure(total));
    uint8_t *base = (uint8_t *)c->stage.p; cudaStream_t st = c->stream;
    if (desc) CK(c, cudaMemcpyAsync(base + off_in, desc, n, cudaMemcpyHostToDevice, st));
    else CK(c, cudaMemcpyAsync(base + off_in, dist_in, nd_in * 2, cudaMemcpyHostToDevice, st));
    zsbk_stage_fse(st, desc ? base + off_in : nullptr, (uint32_t)n, max_symbols, desc ? nullptr : (const int16_t *)(base + off_in), (int)nd_in, al_in,
                   (int *)(base + off_res), (uint32_t *)(base + off_cells), (int16_t *)(base + off_dist));
    int res[4]; uint32_t cells[512]; int16_t d[256];
    CK(c, cudaMemcpyAsync(res, base + off_res, sizeof res, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(cells, base + off_cells, sizeof cells, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(d, base + off_dist, sizeof d, cudaMemcpyDeviceToHost, st));
    int rc = stage_sync(c); if (rc) return rc;
    if (res[0]) return res[0];
    if (al) *al = (uint8_t)res[1];
    if (consumed) *consumed = (size_t)res[3];
    if (dist) { for (int i = 0; i < res[2] && i < 256; i++) dist[i] = d[i]; }
    if (n_dist) *n_dist = (size_t)res[2];
    if (table) for (int i = 0; i < (1 << res[1]); i++) {
        table[3 * i] = (uint16_t)ZSB_CELL_CODE(cells[i]); table[3 * i + 1] = (uint16_t)ZSB_CELL_BASE(cells[i]); table[3 * i + 2] = (uint16_t)ZSB_CELL_NB(cells[i]);
    }
    return ZSB_OK;
}
extern "C" int zsb_fse_table_parse(zsb_ctx *c, const uint8_t *desc, size_t n, int max_symbols, uint8_t *al, uint16_t *table,
                                   size_t *consumed, int16_t *dist, size_t *n_dist) {
    if (!c || !desc || !n) return c && !n ? ZSB_E_EMPTY_INPUT_DATA : ZSB_E_ARG;
    if (max_symbols <= 0 || max_symbols > 64) max_symbols = 64;       // the cell's code field holds 0..63
    return fse_stage(c, desc, n, max_symbols, nullptr, 0, 0, al, table, consumed, dist, n_dist);
}
extern "C" int zsb_fse_table_from_distribution(zsb_ctx *c, uint8_t al, const int16_t *dist, size_t n_dist, uint16_t *table) {
    if (!c || !dist || n_dist == 0 || n_dist > 64) return ZSB_E_ARG;
    return fse_stage(c, nullptr, 0, 64, dist, n_dist, al, nullptr, table, nullptr, nullptr, nullptr);
}
extern "C" int zsb_huffman_parse(zsb_ctx *c, const uint8_t *desc, size_t n, uint8_t lens[256], uint16_t codes[256], size_t *consumed, uint8_t *max_bits) {
    if (!c || !desc || !n) return ZSB_E_ARG;
    CK(c, cudaSetDevice(c->device));
    const size_t off_res = 0, off_lens = 64, off_lut = off_lens + 256, off_in = off_lut + 4096, total = off_in + n + 64;
    CK(c, c->stage.ensure(total));
    uint8_t *base = (uint8_t *)c->stage.p; cudaStream_t st = c->stream;
    CK(c, cudaMemcpyAsync(base + off_in, desc, n, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemsetAsync(base + off_in + n, 0, 64, st));
    zsbk_stage_huf(st, base + off_in, (uint32_t)n, (int *)(base + off_res), base + off_lens, (uint16_t *)(base + off_lut));
    int res[3]; uint8_t l[256]; static thread_local uint16_t lut[2048];
    CK(c, cudaMemcpyAsync(res, base + off_res, sizeof res, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(l, base + off_lens, 256, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(lut, base + off_lut, 4096, cudaMemcpyDeviceToHost, st));
    int rc = stage_sync(c); if (rc) return rc;
    if (res[0]) return res[0];
    const int mb = res[1];
    if (lens) memcpy(lens, l, 256);
    if (codes) {
        // code of a symbol = index of its first LUT cell, shortened to its length
        memset(codes, 0, 512);
        for (int i = (1 << mb) - 1; i >= 0; i--) { const int s = lut[i] & 255, nb = lut[i] >> 8; codes[s] = (uint16_t)(i >> (mb - nb)); }
    }
    if (consumed) *consumed = (size_t)res[2];
    if (max_bits) *max_bits = (uint8_t)mb;
    return ZSB_OK;
}
extern "C" int zsb_execute_sequences(zsb_ctx *c, const uint32_t *seqs, size_t n_seq, const uint8_t *literals, size_t n_lit,
                                     uint8_t *out, size_t out_cap, size_t *out_len) {
    if (!c || (!seqs && n_seq) || (!literals && n_lit) || !out_len) return ZSB_E_ARG;
    if (n_lit > ZSB_BLOCK_MAX || n_seq > ZSB_MAX_NSEQ_PER_BLOCK) return ZSB_E_BLOCK_TOO_LARGE;
    CK(c, cudaSetDevice(c->device));
    // one synthetic frame holding one compressed block with raw literals; the records come from the device too
    const size_t off_work = 0, off_fout = 512, off_frame = 640, off_block = 768, off_cnt = 832, off_list = 896, off_tri = 1024,
                 off_rec = off_tri + ((12 * n_seq + 15) & ~(size_t)15), off_lit = off_rec + 8 * (n_seq + 2), off_out = (off_lit + n_lit + 79) & ~(size_t)15,
                 total = off_out + ZSB_BLOCK_MAX + 64;
    CK(c, c->stage.ensure(total));
    uint8_t *base = (uint8_t *)c->stage.p; cudaStream_t st = c->stream;
    CK(c, cudaMemsetAsync(base, 0, off_tri, st));
    if (n_seq) CK(c, cudaMemcpyAsync(base + off_tri, seqs, 12 * n_seq, cudaMemcpyHostToDevice, st));
    if (n_lit) CK(c, cudaMemcpyAsync(base + off_lit, literals, n_lit, cudaMemcpyHostToDevice, st));
    ZsbBlockWork *w = (ZsbBlockWork *)(base + off_work);
    zsbk_stage_records(st, (const uint32_t *)(base + off_tri), (uint32_t)n_seq, (uint32_t)n_lit, (uint64_t *)(base + off_rec), w);
    ZsbBlockWork hw;
    CK(c, cudaMemcpyAsync(&hw, w, sizeof hw, cudaMemcpyDeviceToHost, st));
    int rc = stage_sync(c); if (rc) return rc;
    if (hw.status) return hw.status;
    hw.lit_type = ZSB_LT_RAW; hw.lit_regen = (uint32_t)n_lit; hw.lit_src = off_lit; hw.nseq = (uint32_t)n_seq; hw.seq_buf = 0; hw.out_off = 0;
    zsb_frame fr; memset(&fr, 0, sizeof fr); fr.kind = 0; fr.first_block = 0; fr.n_blocks = 1;
    zsb_block bl; memset(&bl, 0, sizeof bl); bl.type = ZSB_BT_COMPRESSED; bl.last = 1;
    ZsbFrameOut fo; memset(&fo, 0, sizeof fo); fo.dst_len = hw.out_size;
    uint32_t zero = 0;
    CK(c, cudaMemcpyAsync(w, &hw, sizeof hw, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + off_frame, &fr, sizeof fr, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + off_block, &bl, sizeof bl, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + off_fout, &fo, sizeof fo, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + off_list, &zero, 4, cudaMemcpyHostToDevice, st));
    zsbk_exec2(st, 1, base, (const zsb_frame *)(base + off_frame), (const zsb_block *)(base + off_block), w, (ZsbFrameOut *)(base + off_fout),
              (const uint32_t *)(base + off_list), (const ZsbCounters *)(base + off_cnt), (const uint64_t *)(base + off_rec), base, base + off_out);
    CK(c, cudaMemcpyAsync(&fo, base + off_fout, sizeof fo, cudaMemcpyDeviceToHost, st));
    rc = stage_sync(c); if (rc) return rc;
    if (fo.status) return fo.status;
    *out_len = hw.out_size;
    if (hw.out_size > out_cap) return ZSB_E_DST_TOO_SMALL;
    if (hw.out_size) CK(c, cudaMemcpy(out, base + off_out, hw.out_size, cudaMemcpyDeviceToHost));
    return ZSB_OK;
}
extern "C" int zsb_xxh64(zsb_ctx *c, const uint8_t *data, size_t n, uint64_t *hash) {
    if (!c || (!data && n) || !hash) return ZSB_E_ARG;
    CK(c, cudaSetDevice(c->device));
    const size_t off_fout = 0, off_cnt = 64, off_list = 128, off_data = 256, total = off_data + n + 64;
    CK(c, c->stage.ensure(total));
    uint8_t *base = (uint8_t *)c->stage.p; cudaStream_t st = c->stream;
    CK(c, cudaMemsetAsync(base, 0, off_data, st));
    ZsbFrameOut fo; memset(&fo, 0, sizeof fo); fo.dst_off = off_data; fo.dst_len = n;
    CK(c, cudaMemcpyAsync(base + off_fout, &fo, sizeof fo, cudaMemcpyHostToDevice, st));
    if (n) CK(c, cudaMemcpyAsync(base + off_data, data, n, cudaMemcpyHostToDevice, st));
    zsbk_xxh(st, 1, base, (ZsbFrameOut *)(base + off_fout), (const uint32_t *)(base + off_list), (const ZsbCounters *)(base + off_cnt));
    CK(c, cudaMemcpyAsync(&fo, base + off_fout, sizeof fo, cudaMemcpyDeviceToHost, st));
    int rc = stage_sync(c); if (rc) return rc;
    *hash = fo.xxh64;
    return ZSB_OK;
}

extern "C" const char *zsb_strerror(int s) {
    switch (s) {
    case ZSB_OK: return "ok";
    case ZSB_E_NOT_ENOUGH_BYTES: return "parsing::Error::NotEnoughBytes";
    case ZSB_E_NOT_ENOUGH_BITS: return "parsing::Error::NotEnoughBits";
    case ZSB_E_EMPTY_INPUT_DATA: return "parsing::Error::EmptyInputData";
    case ZSB_E_NULL_BYTE: return "parsing::Error::NullByte";
    case ZSB_E_EMPTY_SLICE: return "parsing::Error::EmptySliceError";
    case ZSB_E_LARGE_ACCURACY_LOG: return "decoders::Error::LargeAccuracyLog";
    case ZSB_E_CORRUPTED_TABLE: return "decoders::Error::CorruptedTable";
    case ZSB_E_SEQ_CODE_MAX: return "decoders::Error::SequenceCodeMaxValueExceeded";
    case ZSB_E_HUFFMAN_MISSING: return "literals::Error::HuffmanDecoderMissing";
    case ZSB_E_STREAMS_TOO_BIG: return "literals::Error::CorruptedStreamsSizeTooBig";
    case ZSB_E_SEQ_RESERVED: return "sequences::Error::ReservedSet";
    case ZSB_E_NO_PREVIOUS_DECODER: return "sequences::Error::NoPreviousDecoder";
    case ZSB_E_WINDOW_TOO_BIG: return "frame::Error::WindowSizeTooBig";
    case ZSB_E_NULL_OFFSET: return "decoding_context::Error::NullOffsetError";
    case ZSB_E_IMPOSSIBLE_VALUE: return "decoding_context::Error::ImpossibleValue";
    case ZSB_E_RESERVED_BLOCK: return "block::Error::ReservedBlockType";
    case ZSB_E_UNRECOGNIZED_MAGIC: return "frame::Error::UnrecognizedMagic";
    case ZSB_E_FRAME_RESERVED: return "frame::Error::ReservedSet";
    case ZSB_E_MISSING_CHECKSUM: return "frame::Error::MissingChecksum";
    case ZSB_E_CORRUPT: return "corrupt entropy data (RFC 8878)";
    case ZSB_E_BLOCK_TOO_LARGE: return "block regenerates more than 128 KiB";
    case ZSB_E_CONTENT_SIZE: return "decoded size differs from Frame_Content_Size";
    case ZSB_E_DST_TOO_SMALL: return "destination too small";
    case ZSB_E_DICTIONARY: return "frame needs a dictionary";
    case ZSB_E_PREVIOUS_FRAME: return "not decoded: an earlier error stopped the frame";
    case ZSB_E_CUDA: return "CUDA error";
    case ZSB_E_ARG: return "bad argument";
    case ZSB_E_NOMEM: return "out of memory";
    default: return "unknown status";
    }
}
extern "C" const char *zsb_version(void) { return "zsb 0.1 (sm_100a)"; }
