// zsb_kernels.h -- launch interface between the host orchestration (zsb_host.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include "zsb_common.h"

struct ZsbCounters {
    uint64_t lit_total;   // bytes of literal scratch the batch needs
    uint64_t seq_total;   // sequence records the batch needs
    uint64_t dst_total;   // bytes produced
    uint32_t n_huf;       // blocks with Huffman-coded literals
    uint32_t n_seq;       // blocks with at least one sequence
    uint32_t overflow;    // scratch too small: every later kernel exits, the host grows it and relaunches
    uint32_t n_slow;      // blocks the fast sequence path handed to the careful decoder
    uint32_t ticket1, ticket2;   // k_plan1 / k_plan2: CTAs done with the per-frame part (the last one runs the scans)
    uint32_t zero;               // always 0: k_seq can order its look-ahead loads behind its cell loads with a data dependency on it
    uint32_t refuse;             // k_seqx executed a block where its frame did not end up (a frame before it failed or lied about its size): the host runs the batch again without k_seqx
    uint64_t lit_over;           // ZSB_REFERENCE_QUIRKS: bytes handed out of the overflow region behind the literal scratch (blocks whose streams
                                 // regenerate more literals than Regenerated_Size announced)
};

cudaError_t zsbk_init();
void zsbk_parse(cudaStream_t st, const uint8_t *src, const zsb_block *blocks, ZsbBlockWork *work, uint32_t nb, uint32_t flags);
void zsbk_plan1(cudaStream_t st, const zsb_frame *frames, uint32_t nf, const zsb_block *blocks, uint32_t nb, ZsbBlockWork *work,
                ZsbFrameOut *fout, uint32_t *huf_list, uint32_t *seq_list, ZsbCounters *cnt, uint64_t lit_cap, uint64_t seq_cap, uint32_t flags);
void zsbk_huf(cudaStream_t st, uint32_t ncomp, const uint8_t *src, uint64_t src_len, ZsbBlockWork *work, const uint32_t *huf_list,
              ZsbCounters *cnt, uint8_t *lit_pool, uint64_t lit_cap, uint64_t over_cap, uint32_t flags);
void zsbk_seq(cudaStream_t st, uint32_t ncomp, const uint8_t *src, ZsbBlockWork *work, const uint32_t *seq_list, ZsbCounters *cnt,
              uint64_t *seq_pool, uint32_t *slow_list, bool shared_device, uint32_t chains_hint, const zsb_frame *frames, const zsb_block *blocks,
              const uint64_t *pre_off /* per frame: where the host expects it in dst, ~0 = unknown; nullptr: k_seq, else k_seqx */,
              const uint8_t *lit_pool, uint8_t *dst);
void zsbk_seq_slow(cudaStream_t st, uint32_t ncomp, const uint8_t *src, uint64_t src_len, ZsbBlockWork *work, const uint32_t *slow_list,
                   const ZsbCounters *cnt, uint64_t *seq_pool);
void zsbk_plan2(cudaStream_t st, const zsb_frame *frames, uint32_t nf, const zsb_block *blocks, ZsbBlockWork *work, ZsbFrameOut *fout,
                ZsbCounters *cnt, uint64_t dst_cap, uint32_t flags, const uint64_t *pre_off);
void zsbk_rawrle(cudaStream_t st, uint32_t n, const uint8_t *src, const zsb_block *blocks, const ZsbBlockWork *work, const ZsbFrameOut *fout,
                 const uint32_t *list, const ZsbCounters *cnt, uint8_t *dst);
void zsbk_exec(cudaStream_t st, uint32_t n, const uint8_t *src, const zsb_frame *frames, const zsb_block *blocks, const ZsbBlockWork *work,
               ZsbFrameOut *fout, const uint32_t *exec_list, const ZsbCounters *cnt, const uint64_t *seq_pool, const uint8_t *lit_pool, uint8_t *dst, bool shared_device,
               uint32_t flags, void *wave /* 32 zeroed bytes per listed frame, or nullptr */, uint32_t *blk_done /* one zeroed word per block of the batch */,
               uint32_t ctas_per_frame /* > 1 with wave: wavefront mode */);
void zsbk_exec2(cudaStream_t st, uint32_t n, const uint8_t *src, const zsb_frame *frames, const zsb_block *blocks, const ZsbBlockWork *work,
                ZsbFrameOut *fout, const uint32_t *exec_list, const ZsbCounters *cnt, const uint64_t *seq_pool, const uint8_t *lit_pool, uint8_t *dst);
// frames of many blocks: one entry per output byte (k_link_init) and pointer jumping (k_link_resolve).  link_frames: n_frames x {uint64 entry offset,
// uint32 frame, uint32 0}; link_blocks: n_blocks x {uint32 block, uint32 index into link_frames}; tickets: n_frames zeroed words
void zsbk_link(cudaStream_t st, uint32_t n_frames, uint32_t n_blocks, const uint8_t *src, const zsb_block *blocks, const ZsbBlockWork *work, ZsbFrameOut *fout,
               const void *link_blocks, const void *link_frames, const ZsbCounters *cnt, const uint64_t *seq_pool, const uint8_t *lit_pool,
               uint32_t *ent, uint32_t *tickets, uint8_t *dst, int n_sm, int which /* 1 k_link_init, 2 k_link_resolve, 3 both */);
void zsbk_xxh_one(cudaStream_t st, uint32_t n_frames, const uint8_t *dst, ZsbFrameOut *fout, const zsb_frame *frames, const void *link_frames, const ZsbCounters *cnt);
void zsbk_publish(cudaStream_t st, void *host_dev_ptr, const ZsbCounters *cnt, const ZsbFrameOut *fout, uint32_t nf);
void zsbk_xxh(cudaStream_t st, uint32_t n, const uint8_t *dst, ZsbFrameOut *fout, const uint32_t *list, const ZsbCounters *cnt);
void zsbk_stage_fse(cudaStream_t st, const uint8_t *desc, uint32_t n, int max_sym, const int16_t *dist_in, int ndist_in, int al_in,
                    int *res, uint32_t *cells, int16_t *dist_out);
void zsbk_stage_huf(cudaStream_t st, const uint8_t *desc, uint32_t n, int *res, uint8_t *lens, uint16_t *lut);
void zsbk_stage_records(cudaStream_t st, const uint32_t *tri, uint32_t nseq, uint32_t nlit, uint64_t *rec, ZsbBlockWork *w);

// ---- zsb_dscan.cu: the frame / block walk on the device (zsb_scan_device)
struct ZsbDscanCand { zsb_frame f; uint64_t end, ea, eb; uint32_t next, n_emitted; int32_t ok; uint32_t pad; };      // a failed walk: f.status, (ea, eb) = the error and its payload, n_emitted = blocks read before it
void zsbk_dscan_find(cudaStream_t st, const uint8_t *src, uint64_t n, unsigned long long *count, uint64_t *pos_out, uint64_t cap, int n_sm);
void zsbk_dscan_parse(cudaStream_t st, const uint8_t *src, uint64_t n, uint32_t flags, uint64_t max_window, const uint64_t *pos_list, uint32_t ncand,
                      ZsbDscanCand *cand, unsigned long long *keys, uint32_t *vals, uint32_t mask);
// jump: (levels + 1) x (ncand + 1) words; after the call the distances (frames from a node to the end of its chain) are in dist_a if `levels` is even, else dist_b
void zsbk_dscan_link(cudaStream_t st, ZsbDscanCand *cand, uint32_t ncand, uint64_t n, const unsigned long long *keys, const uint32_t *vals, uint32_t mask,
                     uint32_t *jump, uint32_t levels, uint32_t *dist_a, uint32_t *dist_b, uint32_t *head);
void zsbk_dscan_order(cudaStream_t st, const uint32_t *jump, uint32_t levels, uint32_t ncand, uint32_t start, uint32_t nfr, const ZsbDscanCand *cand,
                      uint32_t *order, zsb_frame *frames_out, uint64_t *tail);
#define ZSB_DSCAN_TAIL_WORDS 6   // tail of the chain: {ok, end, node, ea, eb, n_emitted} of its last frame
void zsbk_dscan_emit(cudaStream_t st, const uint8_t *src, uint64_t n, uint32_t flags, uint64_t max_window, const zsb_frame *frames, const uint32_t *first_block,
                     uint32_t nfr, zsb_block *blocks_out);
