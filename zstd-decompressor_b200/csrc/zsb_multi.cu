// zsb_multi.cu -- one host buffer, several GPUs, one call (include/zsb.h, "several GPUs from one process").
//
// The reference is one thread on one CPU (src/main.rs:27-60); frames are independent (ZStandard::decode creates a fresh
// DecodingContext, frame.rs:233), so a caller that owns a whole box decodes one buffer on all of its GPUs: the container is
// walked once on the host, the frames are cut into one contiguous range per device -- balanced on decompressed bytes, weighted
// by what each device's host link delivers (zsb_multi_calibrate) -- and every range goes through that device's own context
// (upload, kernels and download pipelined per device exactly as in zsb_decode), one worker thread per device.  No collective:
// every device writes its own slab of the caller's output buffer.  zsb_gather_peer is the optional NVLink gather of
// device-resident slabs into one device for a single-stream consumer: one cudaMemcpyPeerAsync per slab.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include "zsb_common.h"

struct zsb_multi {
    std::vector<int> devices;
    std::vector<zsb_ctx *> ctx;
    std::vector<double> weight;      // relative host-link rate per device (1.0 = equal)
    std::string last_err;
};

extern "C" int zsb_multi_create(zsb_multi **out, const int *device_ids, int n_devices) {
    if (!out || n_devices <= 0 || n_devices > 64) return ZSB_E_ARG;
    *out = nullptr;
    zsb_multi *m = new (std::nothrow) zsb_multi();
    if (!m) return ZSB_E_NOMEM;
    for (int i = 0; i < n_devices; i++) {
        zsb_ctx *c = nullptr;
        const int dev = device_ids ? device_ids[i] : i;
        const int rc = zsb_ctx_create(&c, dev);
        if (rc != ZSB_OK) { for (zsb_ctx *x : m->ctx) zsb_ctx_destroy(x); delete m; return rc; }
        m->devices.push_back(dev); m->ctx.push_back(c); m->weight.push_back(1.0);
    }
    *out = m;
    return ZSB_OK;
}
extern "C" void zsb_multi_destroy(zsb_multi *m) {
    if (!m) return;
    for (zsb_ctx *c : m->ctx) zsb_ctx_destroy(c);
    delete m;
}
extern "C" int zsb_multi_device_count(const zsb_multi *m) { return m ? (int)m->ctx.size() : 0; }
extern "C" zsb_ctx *zsb_multi_ctx(zsb_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
extern "C" const char *zsb_multi_last_error(const zsb_multi *m) { return m ? m->last_err.c_str() : "no context"; }
extern "C" int zsb_multi_set_weights(zsb_multi *m, const double *w) {
    if (!m) return ZSB_E_ARG;
    for (size_t i = 0; i < m->weight.size(); i++) m->weight[i] = (w && w[i] > 0) ? w[i] : 1.0;
    return ZSB_OK;
}
extern "C" int zsb_multi_get_weights(const zsb_multi *m, double *w) {
    if (!m || !w) return ZSB_E_ARG;
    for (size_t i = 0; i < m->weight.size(); i++) w[i] = m->weight[i];
    return ZSB_OK;
}

// What each device's host link delivers while ALL of them copy at once (the situation of a decode call): `bytes` are downloaded
// from every device simultaneously, `reps` times; weight = GB/s of the device relative to the slowest.  (On the 8 x B200 box
// of this project the downloads of GPUs 4-7 run 1.5 x faster than those of GPUs 0-3 when all eight are busy.)
extern "C" int zsb_multi_calibrate(zsb_multi *m, size_t bytes, int reps, double *gbs_out) {
    if (!m) return ZSB_E_ARG;
    const size_t n = m->ctx.size();
    if (bytes < (1u << 20)) bytes = 64u << 20;
    if (reps <= 0) reps = 3;
    std::vector<void *> d(n, nullptr), h(n, nullptr);
    std::vector<cudaStream_t> st(n, nullptr);
    std::vector<cudaEvent_t> e0(n, nullptr), e1(n, nullptr);
    bool ok = true;
    for (size_t i = 0; i < n && ok; i++) {
        ok = cudaSetDevice(m->devices[i]) == cudaSuccess && cudaMalloc(&d[i], bytes) == cudaSuccess && cudaHostAlloc(&h[i], bytes, cudaHostAllocDefault) == cudaSuccess &&
             cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&e0[i]) == cudaSuccess && cudaEventCreate(&e1[i]) == cudaSuccess;
    }
    std::vector<double> gbs(n, 0.0);
    if (ok) {
        for (size_t i = 0; i < n; i++) { cudaSetDevice(m->devices[i]); cudaMemcpyAsync(h[i], d[i], bytes, cudaMemcpyDeviceToHost, st[i]); }     // warm-up
        for (size_t i = 0; i < n; i++) { cudaSetDevice(m->devices[i]); cudaStreamSynchronize(st[i]); }
        for (size_t i = 0; i < n; i++) {
            cudaSetDevice(m->devices[i]);
            cudaEventRecord(e0[i], st[i]);
            for (int r = 0; r < reps; r++) cudaMemcpyAsync(h[i], d[i], bytes, cudaMemcpyDeviceToHost, st[i]);
            cudaEventRecord(e1[i], st[i]);
        }
        for (size_t i = 0; i < n; i++) {
            cudaSetDevice(m->devices[i]);
            cudaStreamSynchronize(st[i]);
            float ms = 0;
            if (cudaEventElapsedTime(&ms, e0[i], e1[i]) == cudaSuccess && ms > 0) gbs[i] = (double)bytes * reps / (ms * 1e-3) / 1e9;
        }
    }
    for (size_t i = 0; i < n; i++) {
        cudaSetDevice(m->devices[i]);
        if (d[i]) cudaFree(d[i]);
        if (h[i]) cudaFreeHost(h[i]);
        if (st[i]) cudaStreamDestroy(st[i]);
        if (e0[i]) cudaEventDestroy(e0[i]);
        if (e1[i]) cudaEventDestroy(e1[i]);
    }
    (void)cudaGetLastError();
    if (!ok) { m->last_err = "zsb_multi_calibrate: allocation failed"; return ZSB_E_CUDA; }
    double lo = 0;
    for (double g : gbs) if (g > 0 && (lo == 0 || g < lo)) lo = g;
    for (size_t i = 0; i < n; i++) { m->weight[i] = (lo > 0 && gbs[i] > 0) ? gbs[i] / lo : 1.0; if (gbs_out) gbs_out[i] = gbs[i]; }
    return ZSB_OK;
}

// contiguous frame ranges whose decompressed bytes are proportional to the device weights
static void weighted_plan(const zsb_frame *frames, size_t nf, const std::vector<double> &w, std::vector<size_t> &first) {
    const size_t n = w.size();
    std::vector<double> cum(nf + 1, 0.0);
    for (size_t f = 0; f < nf; f++) {
        const zsb_frame &fr = frames[f];
        cum[f + 1] = cum[f] + 1.0 + (fr.kind == 0 ? (fr.has_content_size ? (double)fr.content_size : 2.4 * (double)fr.src_len) : (double)fr.src_len);
    }
    double wsum = 0; for (double x : w) wsum += x;
    first.assign(n + 1, 0);
    size_t f = 0; double acc = 0;
    for (size_t s = 1; s < n; s++) {
        acc += w[s - 1];
        const double target = cum[nf] * acc / wsum;
        while (f < nf && cum[f + 1] <= target) f++;
        if (f < nf && target - cum[f] > cum[f + 1] - target) f++;
        if (f < first[s - 1]) f = first[s - 1];
        first[s] = f;
    }
    first[n] = nf;
}

extern "C" int zsb_multi_scan_decode(zsb_multi *m, const uint8_t *src, size_t n, uint8_t *dst, size_t dst_cap, uint32_t flags, uint64_t max_window,
                                     zsb_frame **frames_out, size_t *n_frames, zsb_block **blocks_out, size_t *n_blocks, zsb_result **results_out,
                                     uint64_t *dst_total, uint64_t *err_a, uint64_t *err_b) {
    if (!m || (!src && n) || !dst || !frames_out || !n_frames || !blocks_out || !n_blocks || !results_out) return ZSB_E_ARG;
    *frames_out = nullptr; *blocks_out = nullptr; *results_out = nullptr; *n_frames = 0; *n_blocks = 0;
    flags &= ~(ZSB_SRC_ON_DEVICE | ZSB_DST_ON_DEVICE);
    zsb_frame *frames = nullptr; zsb_block *blocks = nullptr; size_t nf = 0, nb = 0;
    const int scan_rc = zsb_scan(src, n, flags, max_window, &frames, &nf, &blocks, &nb, err_a, err_b);
    // early placement needs every frame's size from its header; otherwise (or with a malformed tail) one device decodes the lot
    bool sized = scan_rc == ZSB_OK;
    uint64_t total = 0;
    std::vector<uint64_t> off(nf + 1, 0);
    for (size_t f = 0; f < nf && sized; f++) {
        off[f] = total;
        if (frames[f].kind == 1) total += (flags & ZSB_PRINT_SKIPPABLE) ? blocks[frames[f].first_block].size : 0;
        else if (frames[f].has_content_size) total += frames[f].content_size;
        else sized = false;
    }
    off[nf] = total;
    if (sized && total > dst_cap) sized = false;
    zsb_result *res = (zsb_result *)calloc(nf + 1, sizeof(zsb_result));
    if (!res) { zsb_free(frames); zsb_free(blocks); return ZSB_E_NOMEM; }
    const size_t nd = m->ctx.size();
    bool placed = false;
    uint64_t out_total = 0;
    int rc = ZSB_OK;
    if (sized && nd > 1 && nf >= nd) {
        std::vector<size_t> first;
        weighted_plan(frames, nf, m->weight, first);
        struct Job { int rc = ZSB_OK; bool exact = true; uint64_t total = 0; };
        std::vector<Job> job(nd);
        std::vector<std::thread> th;
        for (size_t d = 0; d < nd; d++) {
            th.emplace_back([&, d]() {
                const size_t f0 = first[d], f1 = first[d + 1];
                Job &J = job[d];
                if (f0 == f1) return;
                zsb_frame *sf = nullptr; zsb_block *sb = nullptr; size_t snb = 0; uint64_t so = 0, sl = 0;
                J.rc = zsb_shard_extract(frames, nf, blocks, nb, f0, f1, &sf, &sb, &snb, &so, &sl);
                if (J.rc != ZSB_OK) return;
                const size_t cnt = f1 - f0;
                std::vector<uint64_t> o(cnt), l(cnt); std::vector<int32_t> s(cnt); std::vector<uint32_t> x(cnt); std::vector<uint8_t> k(cnt);
                const uint64_t slab = off[f1] - off[f0];
                J.rc = zsb_decode(m->ctx[d], src + so, sl, sf, cnt, sb, snb, dst + off[f0], slab, o.data(), l.data(), s.data(), x.data(), k.data(), &J.total, flags);
                if (J.rc == ZSB_OK) {
                    for (size_t i = 0; i < cnt; i++) {
                        zsb_result &R = res[f0 + i];
                        R.dst_off = off[f0] + o[i]; R.dst_len = l[i]; R.status = s[i]; R.xxh32 = x[i]; R.checksum_ok = k[i];

                        if (s[i] != ZSB_OK || R.dst_off != off[f0 + i]) J.exact = false;      // a frame failed or was not as long as declared: the slabs no longer abut
                    }
                    if (J.total != slab) J.exact = false;
                }
                zsb_free(sf); zsb_free(sb);
            });
        }
        for (std::thread &t : th) t.join();
        placed = true;
        for (size_t d = 0; d < nd; d++) {
            if (job[d].rc != ZSB_OK) { rc = job[d].rc; m->last_err = zsb_last_cuda_error(m->ctx[d]); placed = false; }
            else if (!job[d].exact) placed = false;
        }
        if (rc == ZSB_E_CUDA) { free(res); zsb_free(frames); zsb_free(blocks); return rc; }
        out_total = total;
    }
    if (!placed) {
        // one device, the plain way: frames of unknown size, a frame that failed or disagreed with its header (the output of the frames
        // behind it moves), a malformed container, or a single device
        std::vector<uint64_t> o(nf + 1), l(nf + 1); std::vector<int32_t> s(nf + 1); std::vector<uint32_t> x(nf + 1); std::vector<uint8_t> k(nf + 1);
        rc = zsb_decode(m->ctx[0], src, n, frames, nf, blocks, nb, dst, dst_cap, o.data(), l.data(), s.data(), x.data(), k.data(), &out_total, flags);
        if (rc != ZSB_OK) { m->last_err = zsb_last_cuda_error(m->ctx[0]); free(res); zsb_free(frames); zsb_free(blocks); return rc; }
        std::vector<uint32_t> ea(nf + 1, 0), eb(nf + 1, 0);
        zsb_decode_errors(m->ctx[0], ea.data(), eb.data(), nf);
        for (size_t f = 0; f < nf; f++) { res[f].dst_off = o[f]; res[f].dst_len = l[f]; res[f].status = s[f]; res[f].xxh32 = x[f]; res[f].checksum_ok = k[f]; res[f].err_a = ea[f]; res[f].err_b = eb[f]; }
    }
    *frames_out = frames; *n_frames = nf; *blocks_out = blocks; *n_blocks = nb; *results_out = res;
    if (dst_total) *dst_total = out_total;
    return scan_rc;
}

// NVLink gather: slab i (sizes[i] bytes at src_ptrs[i] on device src_devices[i]) -> dst_ptr + sum(sizes[0..i)) on dst_device, one
// cudaMemcpyPeerAsync per slab, each on a stream of its own on the destination device.  *ms = device time from the first copy to the
// last (events on the destination device), i.e. what a single-stream consumer on dst_device waits for; never part of a decode figure.
extern "C" int zsb_gather_peer(int n, const int *src_devices, const void *const *src_ptrs, const size_t *sizes, int dst_device, void *dst_ptr, float *ms) {
    if (n <= 0 || !src_devices || !src_ptrs || !sizes || !dst_ptr) return ZSB_E_ARG;
    if (cudaSetDevice(dst_device) != cudaSuccess) { (void)cudaGetLastError(); return ZSB_E_CUDA; }
    for (int i = 0; i < n; i++) {
        if (src_devices[i] == dst_device) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dst_device, src_devices[i]) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src_devices[i], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); }
        }
        (void)cudaGetLastError();
    }
    std::vector<cudaStream_t> st(n, nullptr);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t main_st = nullptr;
    bool ok = cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess && cudaStreamCreateWithFlags(&main_st, cudaStreamNonBlocking) == cudaSuccess;
    std::vector<cudaEvent_t> done(n, nullptr);
    for (int i = 0; i < n && ok; i++) ok = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
        // the sources must be complete: the caller synchronised its decodes (zsb_decode_finish); start all copies behind one event
        cudaEventRecord(e0, main_st);
        size_t at = 0;
        for (int i = 0; i < n && ok; i++) {
            cudaStreamWaitEvent(st[i], e0, 0);
            if (sizes[i]) ok = cudaMemcpyPeerAsync((uint8_t *)dst_ptr + at, dst_device, src_ptrs[i], src_devices[i], sizes[i], st[i]) == cudaSuccess;
            cudaEventRecord(done[i], st[i]);
            cudaStreamWaitEvent(main_st, done[i], 0);
            at += sizes[i];
        }
        cudaEventRecord(e1, main_st);
        ok = ok && cudaStreamSynchronize(main_st) == cudaSuccess;
        float t = 0;
        if (ok && cudaEventElapsedTime(&t, e0, e1) == cudaSuccess && ms) *ms = t;
    }
    for (int i = 0; i < n; i++) { if (st[i]) cudaStreamDestroy(st[i]); if (done[i]) cudaEventDestroy(done[i]); }
    if (main_st) cudaStreamDestroy(main_st);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (!ok) { (void)cudaGetLastError(); return ZSB_E_CUDA; }
    return ZSB_OK;
}
