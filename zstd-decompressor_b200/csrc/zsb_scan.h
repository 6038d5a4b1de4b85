// zsb_scan.h -- the host walk as a resumable object (internal; the C ABI is zsb_scan in include/zsb.h).
// zsb_scan runs it to the end; zsb_scan_decode (zsb_host.cu) runs it shard by shard so that the first shards are already
// uploading and decoding while the rest of the buffer is still being walked.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector>
#include "zsb_common.h"

struct ZsbScanner {
    const uint8_t *src; size_t n; size_t pos = 0;
    uint32_t flags; uint64_t max_window;
    std::vector<zsb_frame> frames; std::vector<zsb_block> blocks;
    int code = ZSB_OK; uint64_t err_a = 0, err_b = 0;
    bool done = false;
    ZsbScanner(const uint8_t *s, size_t len, uint32_t fl, uint64_t mw) : src(s), n(len), flags(fl), max_window(mw ? mw : ZSB_MAX_WINDOW_DEFAULT) {}
    bool next();
    int release(zsb_frame **frames_out, size_t *n_frames, zsb_block **blocks_out, size_t *n_blocks) const;
};
