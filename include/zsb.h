/*
 * zsb.h -- C ABI of the B200-native Zstandard decode path ("zsb" = zstd on Blackwell).
 *
 * This header is the drop-in boundary for the decode path of the Rust crate
 * AchilleBailly/zstd-decompressor.  The crate has no FFI of its own; the boundary is its public
 * Rust API, and every entry point below names the reference interface it replaces
 * (paths relative to /root/reference/zstd-decompressor/src).  A Rust `-sys` crate binds this
 * header unchanged.  Plain pointers and sizes only; nothing here throws or
 * aborts across the boundary; there is NO CPU fallback: every decode entry point needs a CUDA
 * device and returns ZSB_E_CUDA if none is usable.
 */
#ifndef ZSB_H
#define ZSB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes -------------------------------------------------------------------------
 * 1..62 are the reference's error variants, same numbering as the CPU oracle (oracle/refcpu.h) so
 * parity tests compare codes directly.  >= 100 are conditions the reference has no variant for. */
enum {
    ZSB_OK = 0,
    ZSB_E_NOT_ENOUGH_BYTES = 1,        /* parsing::Error::NotEnoughBytes          parsing.rs:14 */
    ZSB_E_NOT_ENOUGH_BITS = 2,         /* parsing::Error::NotEnoughBits           parsing.rs:16 */
    ZSB_E_EMPTY_INPUT_DATA = 4,        /* parsing::Error::EmptyInputData          parsing.rs:20 */
    ZSB_E_NULL_BYTE = 5,               /* parsing::Error::NullByte                parsing.rs:22 */
    ZSB_E_EMPTY_SLICE = 6,             /* parsing::Error::EmptySliceError (ZSB_REFERENCE_QUIRKS only) */
    ZSB_E_LARGE_ACCURACY_LOG = 10,     /* decoders::Error::LargeAccuracyLog       decoders/mod.rs:18 */
    ZSB_E_CORRUPTED_TABLE = 11,        /* decoders::Error::CorruptedTable         decoders/mod.rs:20 */
    ZSB_E_SEQ_CODE_MAX = 12,           /* decoders::Error::SequenceCodeMaxValueExceeded */
    ZSB_E_HUFFMAN_MISSING = 20,        /* literals::Error::HuffmanDecoderMissing  literals.rs:14 */
    ZSB_E_STREAMS_TOO_BIG = 21,        /* literals::Error::CorruptedStreamsSizeTooBig */
    ZSB_E_SEQ_RESERVED = 30,           /* sequences::Error::ReservedSet           sequences.rs:18 */
    ZSB_E_NO_PREVIOUS_DECODER = 31,    /* sequences::Error::NoPreviousDecoder     sequences.rs:22 */
    ZSB_E_WINDOW_TOO_BIG = 40,         /* frame::Error::WindowSizeTooBig          frame.rs:32 */
    ZSB_E_NULL_OFFSET = 41,            /* decoding_context::Error::NullOffsetError */
    ZSB_E_IMPOSSIBLE_VALUE = 42,       /* decoding_context::Error::ImpossibleValue decoding_context.rs:14 */
    ZSB_E_RESERVED_BLOCK = 50,         /* block::Error::ReservedBlockType         block.rs:14 */
    ZSB_E_UNRECOGNIZED_MAGIC = 60,     /* frame::Error::UnrecognizedMagic         frame.rs:16 */
    ZSB_E_FRAME_RESERVED = 61,         /* frame::Error::ReservedSet               frame.rs:20 */
    ZSB_E_MISSING_CHECKSUM = 62,       /* frame::Error::MissingChecksum           frame.rs:28 */
    /* no reference variant: RFC 8878 violations the reference would panic on or silently accept */
    ZSB_E_CORRUPT = 100,               /* malformed entropy data (bad weights, stream length mismatch ...) */
    ZSB_E_BLOCK_TOO_LARGE = 101,       /* a compressed block regenerates more than 128 KiB */
    ZSB_E_CONTENT_SIZE = 102,          /* decoded size != Frame_Content_Size */
    ZSB_E_DST_TOO_SMALL = 103,
    ZSB_E_DICTIONARY = 104,            /* frame needs a dictionary (only when ZSB_STRICT_DICT) */
    ZSB_E_PREVIOUS_FRAME = 105,        /* frame not decoded because the scan stopped at an earlier error */
    ZSB_E_CUDA = 200,                  /* no device / CUDA runtime failure: see zsb_last_cuda_error */
    ZSB_E_ARG = 201,
    ZSB_E_NOMEM = 202
};

/* ---- flags ---------------------------------------------------------------------------------- */
#define ZSB_PRINT_SKIPPABLE   0x01u  /* src/main.rs:22-24,45-49 : skippable payloads become output      */
#define ZSB_VERIFY_CHECKSUM   0x02u  /* compute XXH64 of every frame that stores one (frame.rs:239-259) */
#define ZSB_REFERENCE_QUIRKS  0x04u  /* reject exactly what the reference rejects */
#define ZSB_SRC_ON_DEVICE     0x08u  /* src is a device pointer (compressed bytes resident in HBM).  The kernels read whole aligned 128-byte
                                        lines: the allocation must be readable up to the next 128-byte boundary behind src + n (any
                                        cudaMalloc'd buffer is; a sub-range of one is if 127 more bytes follow it)                   */
#define ZSB_DST_ON_DEVICE     0x10u  /* dst is a device pointer (output stays in HBM)                   */
#define ZSB_STRICT_DICT       0x20u  /* fail frames carrying a dict id (the reference ignores it)       */

#define ZSB_MAX_WINDOW_DEFAULT ((uint64_t)8 << 20)   /* frame::MAX_WIN_SIZE frame.rs:44 */
#define ZSB_BLOCK_MAX 131072u

/* ---- frame / block descriptors: result of walking the container ------------------------------
 * == enum Frame { ZStandardFrame, SkippableFrame } frame.rs:47-56 and struct Header frame.rs:103-108,
 *    enum Block block.rs:29-40 (type + extent only; section parsing happens on the GPU). */
typedef struct zsb_frame {
    uint32_t kind;              /* 0 = ZStandardFrame, 1 = SkippableFrame */
    uint32_t magic;
    uint64_t src_off, src_len;  /* whole frame, magic to checksum */
    uint64_t window_size;       /* Header::window_size (descriptor, else content size) */
    uint64_t content_size;      /* Header::content_size, valid if has_content_size */
    uint64_t dict_id;           /* Header::dictionnary_id, valid if has_dict_id (parsed, ignored) */
    uint32_t stored_checksum;   /* ZStandard::checksum(), valid if has_checksum */
    uint32_t first_block, n_blocks;
    int32_t  status;            /* scan status of this frame */
    uint8_t  has_content_size, has_checksum, has_dict_id, single_segment;
    uint32_t reserved;
} zsb_frame;

typedef struct zsb_block {
    uint64_t src_off;           /* payload (after the 3-byte header); skippable frame: its data */
    uint32_t size;              /* Block_Size field: payload bytes, or repeat count for RLE */
    uint32_t frame;             /* owning frame index */
    uint8_t  type;              /* 0 raw, 1 RLE, 2 compressed (block.rs:51-69); 4 = skippable payload */
    uint8_t  last;
    uint8_t  pad[6];
} zsb_block;

/* == ForwardByteParser::new(data).iter() + Frame::parse for every frame (parsing.rs:30-36,
 *    frame.rs:61-100,198-230 minus the section parsing).  Host only, never touches the GPU.
 * Walks until the buffer is exhausted or a frame is malformed.  On a malformed frame the walk
 * stops (like the reference's iterator, whose cursor is undefined after an error): that frame is
 * appended with its error in `status`, *err_pos is the byte offset the reference's error payload
 * refers to, and the function returns that status.  Arrays are malloc'd; free with zsb_free.
 * max_window = 0 means ZSB_MAX_WINDOW_DEFAULT.  flags: ZSB_REFERENCE_QUIRKS, ZSB_STRICT_DICT. */
int zsb_scan(const uint8_t *src, size_t n, uint32_t flags, uint64_t max_window,
             zsb_frame **frames, size_t *n_frames, zsb_block **blocks, size_t *n_blocks,
             uint64_t *err_a, uint64_t *err_b);
void zsb_free(void *p);

/* == what Block::parse keeps of a compressed block besides its extent (block.rs:63-69): LiteralsSection (literals.rs:19-35: type, regenerated
 *    size, jump table, Huffman tree, data) and Sequences (sequences.rs:41-48: number of sequences, the three SymbolCompressionModes with
 *    their FSE tables, bitstream), parsed on the HOST from the block's bytes -- for `--info`, which prints the parsed frame and decodes
 *    nothing (src/main.rs:35-40), and for callers that inspect a block.  The tables the decode uses are built on the GPU.
 *    Returns the first parse error in the reference's order (status, err_a, err_b also in *out). */
typedef struct zsb_fse_state { uint16_t output, baseline, bits_to_read; } zsb_fse_state;      /* fse.rs:72-76 */
typedef struct zsb_sections {
    int32_t  status; uint32_t err_a, err_b;
    uint8_t  lit_type;                 /* 0 RawLiteralsBlock, 1 RLELiteralsBlock, 2 CompressedLiteralsBlock with a tree, 3 without (treeless) */
    uint8_t  rle_byte, max_bits, pad0;
    uint32_t regenerated_size;
    uint16_t jump_table[4];
    uint64_t lit_data_off, lit_data_len;      /* raw: the literals; compressed: `data`, the streams behind tree and jump table */
    uint8_t  code_len[256];                   /* Huffman code length per symbol, 0 = not in the tree */
    uint16_t code[256];                       /* its code, read MSB first (left = 0) */
    uint32_t number_of_sequences;
    uint8_t  mode[3], rle_symbol[3];          /* LL, OF, ML: 0 PredefinedMode, 1 RLEMode(symbol), 2 FseCompressedMode(table), 3 RepeatMode */
    uint8_t  al[3], pad1[3];
    zsb_fse_state table[3][512];              /* 1 << al[t] states when mode[t] == 2 */
    uint64_t bitstream_off, bitstream_len;
} zsb_sections;
int zsb_block_sections(const uint8_t *src, size_t n, const zsb_block *block, uint32_t flags, zsb_sections *out);

/* ---- sharding by frame over the GPUs of one box (no reference counterpart: the crate is single threaded; frames are
 *      independent because ZStandard::decode creates a fresh DecodingContext, frame.rs:233).  Host only.
 * zsb_shard_plan: first[0..n_shards] = contiguous frame ranges [first[s], first[s+1]) balanced on decompressed bytes
 * (Frame_Content_Size where declared, else 2.4 x compressed size).
 * zsb_shard_extract: descriptors of frames [f0, f1) rebased onto the sub-buffer src[*src_off, +*src_len), ready for
 * zsb_decode on one rank.  Arrays are malloc'd; free with zsb_free. */
int zsb_shard_plan(const zsb_frame *frames, size_t n_frames, int n_shards, size_t *first);
int zsb_shard_extract(const zsb_frame *frames, size_t n_frames, const zsb_block *blocks, size_t n_blocks, size_t f0, size_t f1,
                      zsb_frame **out_frames, zsb_block **out_blocks, size_t *out_n_blocks, uint64_t *src_off, uint64_t *src_len);

/* ---- decode context: owns a CUDA stream and the device scratch of one GPU ---------------------
 * == DecodingContext (decoding_context.rs:17-47), except that one zsb_ctx serves a whole batch of
 *    frames: the per-frame state (repeat offsets [1,4,8], Huffman table, repeat tables, output)
 *    lives in device arrays indexed by frame/block. */
typedef struct zsb_ctx zsb_ctx;
int  zsb_ctx_create(zsb_ctx **ctx, int device);
void zsb_ctx_destroy(zsb_ctx *ctx);
/* Run all kernels on an existing CUDA stream (cudaStream_t cast to void*), e.g. torch's current
 * stream so that the caller can bracket the call with its own events.  NULL = the ctx's stream. */
int  zsb_ctx_set_stream(zsb_ctx *ctx, void *cuda_stream);
const char *zsb_last_cuda_error(const zsb_ctx *ctx);

/* == Frame::decode(self) -> Result<Vec<u8>> for every frame of the batch (frame.rs:79-84,232-260
 *    -> block.rs:74-99 -> literals.rs:49-86, sequences.rs:191-237, decoding_context.rs:50-106).
 * src/n: the same buffer zsb_scan walked (host, or device with ZSB_SRC_ON_DEVICE).
 * dst/dst_cap: output, frames are written back to back in order (skippable payloads only with
 * ZSB_PRINT_SKIPPABLE, as src/main.rs:43-53 concatenates them).
 * Per frame (arrays of n_frames, host memory, any may be NULL): dst_off/dst_len = where its bytes
 * are; status = ZSB_OK or the first error of the frame (one bad frame does not fail the batch;
 * its dst_len is 0); xxh32 = low 32 bits of XXH64(seed 0) of its content when ZSB_VERIFY_CHECKSUM
 * and the frame stores a checksum; checksum_ok = 1 if equal to the stored value (a mismatch is
 * reported, not an error -- the reference only prints a warning, frame.rs:251-254).
 * *dst_total = bytes produced.  Returns ZSB_OK if the batch ran (inspect status[]), else an error.
 * With host buffers and >= 512 frames the batch is cut into shards by frame (a small first shard, then growing), each with its
 * own stream and scratch: upload, kernels and download of different shards overlap (same results).  Frames that declare
 * Frame_Content_Size are placed at once and leave right behind their kernels; from the first frame without one on, a shard
 * decodes into a device buffer of its own and is sent to the host once the sizes before it are known.  Only for page-locked src and dst (zsb_host_alloc, cudaHostAlloc, cudaHostRegister): copies from and to pageable
 * memory block the caller, so pageable buffers are decoded as one batch. */
int zsb_decode(zsb_ctx *ctx, const uint8_t *src, size_t n,
               const zsb_frame *frames, size_t n_frames, const zsb_block *blocks, size_t n_blocks,
               uint8_t *dst, size_t dst_cap, uint64_t *dst_off, uint64_t *dst_len,
               int32_t *status, uint32_t *xxh32, uint8_t *checksum_ok,
               uint64_t *dst_total, uint32_t flags);

/* zsb_scan + zsb_decode on host buffers in one call, the host walk overlapped with the GPU work: the walk stops at shard
 * boundaries and the frames found so far are already uploading and decoding while the rest of the buffer is walked (== the
 * reference's FrameIterator feeding Frame::decode, frame.rs:94-99 / main.rs:42-53).  Returns what zsb_scan returns; frames /
 * blocks as from zsb_scan, results[f] as the per-frame arrays of zsb_decode (all three malloc'd: zsb_free). */
typedef struct zsb_result {
    uint64_t dst_off, dst_len;
    int32_t  status;
    uint32_t xxh32;
    uint32_t err_a, err_b;      /* payload of the frame's error variant where the reference's has one (NotEnoughBytes {requested, available} ...) */
    uint8_t  checksum_ok, pad[7];
} zsb_result;
int zsb_scan_decode(zsb_ctx *ctx, const uint8_t *src, size_t n, uint8_t *dst, size_t dst_cap, uint32_t flags, uint64_t max_window,
                    zsb_frame **frames, size_t *n_frames, zsb_block **blocks, size_t *n_blocks,
                    zsb_result **results, uint64_t *dst_total, uint64_t *err_a, uint64_t *err_b);

/* zsb_scan for compressed bytes that are resident in HBM (d_src: device pointer on the context's device, readable from the 16-byte
 * boundary at or below d_src up to the next 128-byte boundary behind d_src + n -- any cudaMalloc'd buffer or sub-range of one is): the walk itself runs on the GPU.  == ForwardByteParser::iter + Frame::parse like zsb_scan
 * (parsing.rs:29-112, frame.rs:61-230, block.rs:43-72), but "each header says where the next one starts" is not followed as one
 * dependent chain: every frame magic in the buffer is a candidate start, a lane per candidate walks its frame, pointer doubling over
 * the successor lists orders the chain that begins at offset 0 (csrc/zsb_dscan.cu).  Same descriptor arrays (host memory, zsb_free),
 * return value and error payload as zsb_scan on the same bytes.  Only the descriptors cross the host link; when the chain ends in a
 * malformed frame under ZSB_REFERENCE_QUIRKS the bytes from that frame on are fetched for the reference's eager section parsing.
 * zsb_scan_decode with ZSB_SRC_ON_DEVICE uses it, followed by one zsb_decode batch (dst on the host, or on the device with
 * ZSB_DST_ON_DEVICE). */
int zsb_scan_device(zsb_ctx *ctx, const uint8_t *d_src, size_t n, uint32_t flags, uint64_t max_window,
                    zsb_frame **frames, size_t *n_frames, zsb_block **blocks, size_t *n_blocks,
                    uint64_t *err_a, uint64_t *err_b);

/* Page-locked host buffers for src/dst of zsb_decode: copies from and to pageable memory are staged by the driver and
 * block the calling thread, which serialises the shards of the pipelined path (a binding would back its input and
 * output Vec<u8> with these).  NULL on failure.  Not for zsb_free. */
void *zsb_host_alloc(size_t n);
void  zsb_host_free(void *p);

/* Split zsb_decode for callers that keep data resident and time the GPU work only:
 * prepare enqueues the upload of the descriptors (and of src unless ZSB_SRC_ON_DEVICE: src must stay unchanged until finish)
 * and sizes the scratch;
 * launch enqueues every kernel of the batch on the ctx stream and returns without synchronising;
 * finish synchronises and returns the per-frame results. */
int zsb_decode_prepare(zsb_ctx *ctx, const uint8_t *src, size_t n,
                       const zsb_frame *frames, size_t n_frames, const zsb_block *blocks, size_t n_blocks,
                       uint8_t *dst, size_t dst_cap, uint32_t flags);
int zsb_decode_launch(zsb_ctx *ctx);
int zsb_decode_finish(zsb_ctx *ctx, uint64_t *dst_off, uint64_t *dst_len, int32_t *status,
                      uint32_t *xxh32, uint8_t *checksum_ok, uint64_t *dst_total);
/* Per-frame payloads of the errors the last zsb_decode / zsb_decode_finish reported in status[] (parsing::Error::NotEnoughBytes
 * {requested, available} of a block's sections, tests/block.rs:72-78): err_a[f], err_b[f], 0 where the variant carries none. */
int zsb_decode_errors(const zsb_ctx *ctx, uint32_t *err_a, uint32_t *err_b, size_t n_frames);
/* Number of kernels enqueued by the last zsb_decode_launch, and per-kernel device times (ms) of the
 * last launch when profiling was enabled with zsb_ctx_set_profile(ctx, 1).  names[i] are static. */
int zsb_ctx_set_profile(zsb_ctx *ctx, int enable);
int zsb_last_launch_count(const zsb_ctx *ctx);
/* How the sequence stage of the last batch ran: 0 = k_seq (records to HBM, executed afterwards), 1 = k_seqx (the first block of
 * every frame with a known place executed by the sequence kernel itself), 2 = k_seqx, refused (a frame did not end up where its
 * declared sizes put it) and the batch run again with k_seq.  The results are the same in all three. */
int zsb_last_seqx_state(const zsb_ctx *ctx);
int zsb_last_kernel_times(const zsb_ctx *ctx, const char **names, float *ms, int cap);
/* Synchronises, then averages each kernel over the launches recorded since profiling was switched on
 * (at most the last 32). */
int zsb_kernel_times_avg(zsb_ctx *ctx, const char **names, float *ms, int cap, int *n_launches);

/* ---- several GPUs from one process (no reference counterpart: src/main.rs:27-60 is one thread on one CPU; frames are independent,
 *      frame.rs:233).  zsb_multi owns one zsb_ctx per device and decodes ONE host buffer on all of them in one call: the container is
 *      walked once, the frames are cut into one contiguous range per device, balanced on decompressed bytes and weighted by what each
 *      device's host link delivers (zsb_multi_calibrate measures it with all devices copying at once; zsb_multi_set_weights sets it),
 *      and every range runs through its device's context on a worker thread of its own (upload, kernels, download pipelined per
 *      device as in zsb_decode).  No collective: every device writes its own slab of dst.  Results as zsb_scan_decode.  Frames without
 *      Frame_Content_Size, a frame that fails or disagrees with its header, or a malformed container end in one plain decode on the
 *      first device (same results).  src / dst: page-locked host memory (zsb_host_alloc) keeps the copies of all devices asynchronous. */
typedef struct zsb_multi zsb_multi;
int  zsb_multi_create(zsb_multi **m, const int *device_ids /* NULL: 0 .. n-1 */, int n_devices);
void zsb_multi_destroy(zsb_multi *m);
int  zsb_multi_device_count(const zsb_multi *m);
zsb_ctx *zsb_multi_ctx(zsb_multi *m, int i);            /* the context of device i (owned by m), e.g. for resident decodes per device */
int  zsb_multi_calibrate(zsb_multi *m, size_t bytes, int reps, double *gbs /* n_devices, may be NULL */);
int  zsb_multi_set_weights(zsb_multi *m, const double *weights /* n_devices; NULL: equal */);
int  zsb_multi_get_weights(const zsb_multi *m, double *weights);
const char *zsb_multi_last_error(const zsb_multi *m);
int  zsb_multi_scan_decode(zsb_multi *m, const uint8_t *src, size_t n, uint8_t *dst, size_t dst_cap, uint32_t flags, uint64_t max_window,
                           zsb_frame **frames, size_t *n_frames, zsb_block **blocks, size_t *n_blocks,
                           zsb_result **results, uint64_t *dst_total, uint64_t *err_a, uint64_t *err_b);
/* NVLink gather for a single-stream consumer: device-resident slabs (slab i: sizes[i] bytes at src_ptrs[i] on src_devices[i]) ->
 * dst_ptr on dst_device, back to back, one cudaMemcpyPeerAsync per slab; *ms = device time of the whole gather.  Timed and reported
 * apart from any decode figure. */
int  zsb_gather_peer(int n, const int *src_devices, const void *const *src_ptrs, const size_t *sizes, int dst_device, void *dst_ptr, float *ms);

/* == whole-program behaviour of src/main.rs:42-58 : scan + decode + concatenate into a malloc'd
 *    buffer; all-or-nothing like the CLI (first error => no output).  Convenience for bindings. */
int zsb_decompress(zsb_ctx *ctx, const uint8_t *src, size_t n, uint32_t flags,
                   uint8_t **out, size_t *out_len, uint64_t *err_a, uint64_t *err_b);

/* ---- stage-level entry points (the reference's tests call these stages directly) -------------- */
/* == parse_fse_table (fse.rs:16-69) + FseTable::from_distribution (fse.rs:110-202) on the GPU.
 *    in: description bytes; out: al, table of (1<<al) x {output, baseline, bits_to_read} u16 triples,
 *    consumed = ForwardBitParser::bytes_read().  dist (optional, 64 entries) gets the counts. */
int zsb_fse_table_parse(zsb_ctx *ctx, const uint8_t *desc, size_t n, int max_symbols,
                        uint8_t *al, uint16_t *table, size_t *consumed, int16_t *dist, size_t *n_dist);
/* == FseTable::from_distribution (fse.rs:110-202) on the GPU */
int zsb_fse_table_from_distribution(zsb_ctx *ctx, uint8_t al, const int16_t *dist, size_t n_dist, uint16_t *table);
/* == HuffmanDecoder::parse (huffman.rs:80-130) + from_weights (huffman.rs:177-203) on the GPU:
 *    per symbol code length (0 = absent) and canonical code value, as the reference's tree assigns. */
int zsb_huffman_parse(zsb_ctx *ctx, const uint8_t *desc, size_t n, uint8_t lens[256], uint16_t codes[256],
                      size_t *consumed, uint8_t *max_bits);
/* == DecodingContext::execute_sequences (decoding_context.rs:78-106) on the GPU with a fresh context
 *    (offsets [1,4,8]); seqs = n_seq triples (literal_length, offset_value, match_length) as u32. */
int zsb_execute_sequences(zsb_ctx *ctx, const uint32_t *seqs, size_t n_seq, const uint8_t *literals, size_t n_lit,
                          uint8_t *out, size_t out_cap, size_t *out_len);
/* XXH64 (seed 0) of a host buffer, computed on the GPU (twox_hash::XxHash64 at frame.rs:240-244) */
int zsb_xxh64(zsb_ctx *ctx, const uint8_t *data, size_t n, uint64_t *hash);

const char *zsb_strerror(int status);
const char *zsb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ZSB_H */
