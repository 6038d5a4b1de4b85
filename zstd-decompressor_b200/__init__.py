"""zstd-decompressor_b200 -- host-side mirror of the `zstd_decompressor` crate over the zsb C ABI.

The names and call shapes follow the reference crate (paths relative to
/root/reference/zstd-decompressor/src) so that tests read like the reference's own:

    ForwardByteParser(data).iter()  -> FrameIterator -> Frame      parsing.rs:30-36, frame.rs:87-100
    Frame.decode() -> bytes                                         frame.rs:79-84
    ZStandard.header() / .checksum() / .blocks()                   frame.rs:262-272
    Header{content_checksum_flag, window_size, dictionnary_id, content_size}   frame.rs:103-108
    DecodingContext(window_size).execute_sequences(seqs, literals) decoding_context.rs:29,78
    MAX_WIN_SIZE                                                   frame.rs:44
    decompress(data, print_skippable) == src/main.rs:42-58

plus `Decoder`, the batch interface the CUDA path is built for (all frames of a buffer in one go,
host or device-resident buffers).  Everything that decodes goes through libzsb.so (sm_100a kernels);
there is no CPU decode path here and importing the package without the built library fails loudly.
"""
import ctypes as C
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzsb.so")

MAX_WIN_SIZE = 8 << 20          # frame.rs:44

# flags (include/zsb.h)
PRINT_SKIPPABLE, VERIFY_CHECKSUM, REFERENCE_QUIRKS, SRC_ON_DEVICE, DST_ON_DEVICE, STRICT_DICT = 1, 2, 4, 8, 16, 32
OK = 0
E_DST_TOO_SMALL = 103
E_CUDA = 200


class ZsbFrame(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("magic", C.c_uint32), ("src_off", C.c_uint64), ("src_len", C.c_uint64),
                ("window_size", C.c_uint64), ("content_size", C.c_uint64), ("dict_id", C.c_uint64),
                ("stored_checksum", C.c_uint32), ("first_block", C.c_uint32), ("n_blocks", C.c_uint32), ("status", C.c_int32),
                ("has_content_size", C.c_uint8), ("has_checksum", C.c_uint8), ("has_dict_id", C.c_uint8), ("single_segment", C.c_uint8),
                ("reserved", C.c_uint32)]


class ZsbBlock(C.Structure):
    _fields_ = [("src_off", C.c_uint64), ("size", C.c_uint32), ("frame", C.c_uint32), ("type", C.c_uint8), ("last", C.c_uint8),
                ("pad", C.c_uint8 * 6)]


class ZsbResult(C.Structure):
    _fields_ = [("dst_off", C.c_uint64), ("dst_len", C.c_uint64), ("status", C.c_int32), ("xxh32", C.c_uint32), ("err_a", C.c_uint32), ("err_b", C.c_uint32),
                ("checksum_ok", C.c_uint8), ("pad", C.c_uint8 * 7)]


class ZsbError(Exception):
    """Carries the zsb status code; codes 1..62 are the reference's error variants (include/zsb.h)."""
    def __init__(self, code, a=0, b=0, what=""):
        self.code, self.a, self.b = code, a, b
        msg = lib().zsb_strerror(code).decode()
        super().__init__(f"{msg} (code {code}, a={a}, b={b}) {what}".strip())


def _build_module():
    spec = importlib.util.spec_from_file_location("_zsb_build", os.path.join(_HERE, "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def build(force=False, verbose=False):
    """Compile libzsb.so for sm_100a (nvcc).  Raises if nvcc is missing or the build fails."""
    return _build_module().build(force=force, verbose=verbose)


def build_cli(force=False):
    """Build cli/zstd-decompressor (flag compatible with the reference's src/main.rs)."""
    return _build_module().build_cli(force=force)


_lib = None


def lib():
    """The loaded C ABI.  Builds it first if the sources are newer.  No fallback: raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("ZSB_LIB_PATH") and _build_module().needs_build():
        build()                                            # raises if nvcc is missing or the build fails: a stale library is never loaded silently
    L = C.CDLL(os.environ.get("ZSB_LIB_PATH") or _SO)      # ZSB_LIB_PATH: development knob, a differently tuned build of the same library
    u8p, sz, vp = C.POINTER(C.c_uint8), C.c_size_t, C.c_void_p
    u64p, i32p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    L.zsb_scan.argtypes = [vp, sz, C.c_uint32, C.c_uint64, C.POINTER(C.POINTER(ZsbFrame)), C.POINTER(sz),
                           C.POINTER(C.POINTER(ZsbBlock)), C.POINTER(sz), u64p, u64p]
    L.zsb_free.argtypes = [vp]
    L.zsb_host_alloc.argtypes = [C.c_size_t]; L.zsb_host_alloc.restype = vp
    L.zsb_host_free.argtypes = [vp]; L.zsb_host_free.restype = None
    L.zsb_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.zsb_ctx_destroy.argtypes = [vp]
    L.zsb_ctx_set_stream.argtypes = [vp, vp]
    L.zsb_last_cuda_error.restype = C.c_char_p; L.zsb_last_cuda_error.argtypes = [vp]
    L.zsb_ctx_set_profile.argtypes = [vp, C.c_int]
    L.zsb_last_launch_count.argtypes = [vp]
    L.zsb_last_seqx_state.argtypes = [vp]
    L.zsb_last_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.zsb_kernel_times_avg.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    L.zsb_decode.argtypes = [vp, vp, sz, C.POINTER(ZsbFrame), sz, C.POINTER(ZsbBlock), sz, vp, sz, u64p, u64p, i32p, u32p, u8p, u64p, C.c_uint32]
    L.zsb_decode_prepare.argtypes = [vp, vp, sz, C.POINTER(ZsbFrame), sz, C.POINTER(ZsbBlock), sz, vp, sz, C.c_uint32]
    L.zsb_decode_launch.argtypes = [vp]
    L.zsb_decode_errors.argtypes = [vp, u32p, u32p, sz]
    L.zsb_decode_finish.argtypes = [vp, u64p, u64p, i32p, u32p, u8p, u64p]
    L.zsb_scan_decode.argtypes = [vp, vp, sz, vp, sz, C.c_uint32, C.c_uint64, C.POINTER(C.POINTER(ZsbFrame)), C.POINTER(sz),
                                  C.POINTER(C.POINTER(ZsbBlock)), C.POINTER(sz), C.POINTER(C.POINTER(ZsbResult)), u64p, u64p, u64p]
    L.zsb_scan_device.argtypes = [vp, vp, sz, C.c_uint32, C.c_uint64, C.POINTER(C.POINTER(ZsbFrame)), C.POINTER(sz),
                                  C.POINTER(C.POINTER(ZsbBlock)), C.POINTER(sz), u64p, u64p]
    L.zsb_decompress.argtypes = [vp, vp, sz, C.c_uint32, C.POINTER(vp), C.POINTER(sz), u64p, u64p]
    L.zsb_fse_table_parse.argtypes = [vp, C.c_char_p, sz, C.c_int, u8p, C.POINTER(C.c_uint16), C.POINTER(sz), C.POINTER(C.c_int16), C.POINTER(sz)]
    L.zsb_fse_table_from_distribution.argtypes = [vp, C.c_uint8, C.POINTER(C.c_int16), sz, C.POINTER(C.c_uint16)]
    L.zsb_huffman_parse.argtypes = [vp, C.c_char_p, sz, u8p, C.POINTER(C.c_uint16), C.POINTER(sz), u8p]
    L.zsb_execute_sequences.argtypes = [vp, u32p, sz, C.c_char_p, sz, vp, sz, C.POINTER(sz)]
    L.zsb_xxh64.argtypes = [vp, C.c_char_p, sz, u64p]
    L.zsb_shard_plan.argtypes = [C.POINTER(ZsbFrame), sz, C.c_int, C.POINTER(sz)]
    L.zsb_shard_extract.argtypes = [C.POINTER(ZsbFrame), sz, C.POINTER(ZsbBlock), sz, sz, sz, C.POINTER(C.POINTER(ZsbFrame)), C.POINTER(C.POINTER(ZsbBlock)),
                                    C.POINTER(sz), u64p, u64p]
    dp = C.POINTER(C.c_double)
    L.zsb_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.zsb_multi_destroy.argtypes = [vp]; L.zsb_multi_destroy.restype = None
    L.zsb_multi_device_count.argtypes = [vp]
    L.zsb_multi_ctx.argtypes = [vp, C.c_int]; L.zsb_multi_ctx.restype = vp
    L.zsb_multi_calibrate.argtypes = [vp, sz, C.c_int, dp]
    L.zsb_multi_set_weights.argtypes = [vp, dp]
    L.zsb_multi_get_weights.argtypes = [vp, dp]
    L.zsb_multi_last_error.argtypes = [vp]; L.zsb_multi_last_error.restype = C.c_char_p
    L.zsb_multi_scan_decode.argtypes = [vp, vp, sz, vp, sz, C.c_uint32, C.c_uint64, C.POINTER(C.POINTER(ZsbFrame)), C.POINTER(sz),
                                        C.POINTER(C.POINTER(ZsbBlock)), C.POINTER(sz), C.POINTER(C.POINTER(ZsbResult)), u64p, u64p, u64p]
    L.zsb_gather_peer.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(vp), C.POINTER(sz), C.c_int, vp, C.POINTER(C.c_float)]
    L.zsb_strerror.restype = C.c_char_p; L.zsb_strerror.argtypes = [C.c_int]
    L.zsb_version.restype = C.c_char_p
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "zsb_scan", "zsb_free", "zsb_ctx_create", "zsb_ctx_destroy", "zsb_ctx_set_stream", "zsb_last_cuda_error", "zsb_ctx_set_profile",
    "zsb_last_launch_count", "zsb_last_seqx_state", "zsb_last_kernel_times", "zsb_kernel_times_avg", "zsb_decode", "zsb_decode_prepare", "zsb_decode_launch", "zsb_decode_finish",
    "zsb_decompress", "zsb_fse_table_parse", "zsb_fse_table_from_distribution", "zsb_huffman_parse", "zsb_execute_sequences",
    "zsb_xxh64", "zsb_strerror", "zsb_version", "zsb_shard_plan", "zsb_shard_extract", "zsb_host_alloc", "zsb_host_free", "zsb_scan_decode",
    "zsb_multi_create", "zsb_multi_destroy", "zsb_multi_device_count", "zsb_multi_ctx", "zsb_multi_calibrate", "zsb_multi_set_weights", "zsb_multi_get_weights",
    "zsb_multi_last_error", "zsb_multi_scan_decode", "zsb_gather_peer", "zsb_decode_errors", "zsb_block_sections", "zsb_scan_device"]


# ------------------------------------------------------------------------------------------ scan
class Scan:
    """Result of zsb_scan: frame and block descriptor arrays (owned; freed with the object)."""
    def __init__(self, data, flags=0, max_window=0):
        self._keep = data
        self.buf, self.n = _as_buffer(data)
        fp, bp = C.POINTER(ZsbFrame)(), C.POINTER(ZsbBlock)()
        nf, nb, ea, eb = C.c_size_t(), C.c_size_t(), C.c_uint64(), C.c_uint64()
        self.status = lib().zsb_scan(self.buf, self.n, flags, max_window, C.byref(fp), C.byref(nf), C.byref(bp), C.byref(nb), C.byref(ea), C.byref(eb))
        self.frames, self.blocks, self.n_frames, self.n_blocks = fp, bp, nf.value, nb.value
        self.err_a, self.err_b = ea.value, eb.value

    def error(self):
        return ZsbError(self.status, self.err_a, self.err_b) if self.status else None

    def __del__(self):
        try:
            if getattr(self, "frames", None): lib().zsb_free(self.frames)
            if getattr(self, "blocks", None): lib().zsb_free(self.blocks)
        except Exception:
            pass


class DeviceScan(Scan):
    """zsb_scan_device: the walk over a buffer resident in HBM, run on the GPU.  Same attributes as Scan; buf is the device pointer."""
    def __init__(self, ctx, dev_ptr, n, flags=0, max_window=0):
        self._keep = None
        self.buf, self.n = C.c_void_p(dev_ptr), n
        fp, bp = C.POINTER(ZsbFrame)(), C.POINTER(ZsbBlock)()
        nf, nb, ea, eb = C.c_size_t(), C.c_size_t(), C.c_uint64(), C.c_uint64()
        self.status = lib().zsb_scan_device(ctx.h, self.buf, n, flags, max_window, C.byref(fp), C.byref(nf), C.byref(bp), C.byref(nb), C.byref(ea), C.byref(eb))
        self.frames, self.blocks, self.n_frames, self.n_blocks = fp, bp, nf.value, nb.value
        self.err_a, self.err_b = ea.value, eb.value
        if not fp and self.status:
            raise ZsbError(self.status, what="zsb_scan_device" + (": " + ctx.cuda_error() if self.status == E_CUDA else ""))


def _as_buffer(data):
    """-> (ctypes pointer-compatible object, length) for bytes / bytearray / memoryview / numpy / (ptr, n)."""
    if isinstance(data, tuple):
        return C.c_void_p(data[0]), data[1]
    if isinstance(data, bytes):
        return C.cast(C.c_char_p(data), C.c_void_p), len(data)
    mv = memoryview(data).cast("B")
    if mv.readonly:
        b = bytes(mv)
        return C.cast(C.c_char_p(b), C.c_void_p), len(b)
    arr = (C.c_uint8 * len(mv)).from_buffer(mv)
    return C.cast(arr, C.c_void_p), len(mv)


# ------------------------------------------------------------------------------------------ context
class Context:
    """zsb_ctx: one CUDA stream + device scratch on one GPU."""
    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib().zsb_ctx_create(C.byref(h), device)
        if rc:
            raise ZsbError(rc, what="zsb_ctx_create: no usable CUDA device (this library has no CPU path)")
        self.h = h

    def set_stream(self, cuda_stream_ptr):
        lib().zsb_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr))

    def set_profile(self, on=True):
        lib().zsb_ctx_set_profile(self.h, int(on))

    def last_launch_count(self):
        return lib().zsb_last_launch_count(self.h)

    def last_seqx_state(self):
        """0: k_seq, 1: k_seqx executed the placed blocks, 2: k_seqx refused and the batch ran again with k_seq."""
        return lib().zsb_last_seqx_state(self.h)

    def kernel_times(self):
        names = (C.c_char_p * 16)(); ms = (C.c_float * 16)()
        n = lib().zsb_last_kernel_times(self.h, names, ms, 16)
        return [(names[i].decode(), ms[i]) for i in range(n)]

    def kernel_times_avg(self):
        """[(kernel, mean ms)] over the launches since set_profile(True), and how many launches that was"""
        names = (C.c_char_p * 16)(); ms = (C.c_float * 16)(); nl = C.c_int()
        n = lib().zsb_kernel_times_avg(self.h, names, ms, 16, C.byref(nl))
        return [(names[i].decode(), ms[i]) for i in range(n)], nl.value

    def cuda_error(self):
        return lib().zsb_last_cuda_error(self.h).decode()

    def close(self):
        if getattr(self, "h", None):
            lib().zsb_ctx_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class BatchResult:
    def __init__(self, nf):
        self.dst_off = (C.c_uint64 * max(nf, 1))(); self.dst_len = (C.c_uint64 * max(nf, 1))()
        self.status = (C.c_int32 * max(nf, 1))(); self.xxh32 = (C.c_uint32 * max(nf, 1))(); self.checksum_ok = (C.c_uint8 * max(nf, 1))()
        self.total = C.c_uint64(); self.nf = nf

    def errors(self, ctx):
        """[(err_a, err_b)] per frame: the payloads of the reference's error variants (zsb_decode_errors)"""
        a = (C.c_uint32 * max(self.nf, 1))(); b = (C.c_uint32 * max(self.nf, 1))()
        lib().zsb_decode_errors(ctx.h, a, b, self.nf)
        return [(a[i], b[i]) for i in range(self.nf)]

    def first_error(self):
        raw = bytes(self.status)[:4 * self.nf]
        if raw.count(0) == len(raw):            # the common case at C speed
            return None
        for i in range(self.nf):
            if self.status[i]:
                return i, self.status[i]
        return None


class ScanDecode:
    """zsb_scan_decode: walk + decode of a host buffer into a host buffer in one call (the walk overlaps the GPU work).
    src / dst are (pointer, length) pairs; page-locked memory (zsb_host_alloc, torch pinned tensors) keeps the copies asynchronous.
    Attributes as Scan (frames, blocks, n_frames, n_blocks, status = what zsb_scan would return) plus results[f] (ZsbResult) and total."""
    def __init__(self, ctx, src, dst, flags=VERIFY_CHECKSUM, max_window=0):
        fp, bp, rp = C.POINTER(ZsbFrame)(), C.POINTER(ZsbBlock)(), C.POINTER(ZsbResult)()
        nf, nb, tot, ea, eb = C.c_size_t(), C.c_size_t(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.status = lib().zsb_scan_decode(ctx.h, C.c_void_p(src[0]), src[1], C.c_void_p(dst[0]), dst[1], flags, max_window,
                                            C.byref(fp), C.byref(nf), C.byref(bp), C.byref(nb), C.byref(rp), C.byref(tot), C.byref(ea), C.byref(eb))
        self.frames, self.blocks, self.results = fp, bp, rp
        self.n_frames, self.n_blocks, self.total = nf.value, nb.value, tot.value
        self.err_a, self.err_b = ea.value, eb.value
        if not rp and self.status:                      # the call itself failed (arguments, CUDA, memory): nothing was produced
            raise ZsbError(self.status, what="zsb_scan_decode" + (": " + ctx.cuda_error() if self.status == E_CUDA else ""))

    def first_error(self):
        n = self.n_frames
        raw = C.string_at(self.results, C.sizeof(ZsbResult) * n) if n else b""
        if all(raw[16 + k::C.sizeof(ZsbResult)].count(0) == n for k in range(4)):      # every status zero: the common case at C speed
            return None
        for i in range(n):
            if self.results[i].status:
                return i, self.results[i].status
        return None

    def __del__(self):
        try:
            for p in ("frames", "blocks", "results"):
                if getattr(self, p, None): lib().zsb_free(getattr(self, p))
        except Exception:
            pass


class MultiContext:
    """zsb_multi: one context per GPU of the box, one host buffer decoded on all of them in one call (frames sharded by contiguous ranges
    weighted by each device's host link).  src / dst are (pointer, length) pairs of page-locked host memory."""
    def __init__(self, devices):
        h = C.c_void_p()
        ids = (C.c_int * len(devices))(*devices)
        rc = lib().zsb_multi_create(C.byref(h), ids, len(devices))
        if rc:
            raise ZsbError(rc, what="zsb_multi_create")
        self.h, self.n = h, len(devices)

    def calibrate(self, nbytes=64 << 20, reps=3):
        g = (C.c_double * self.n)()
        rc = lib().zsb_multi_calibrate(self.h, nbytes, reps, g)
        if rc:
            raise ZsbError(rc, what="zsb_multi_calibrate")
        return list(g)

    def weights(self):
        w = (C.c_double * self.n)(); lib().zsb_multi_get_weights(self.h, w); return list(w)

    def set_weights(self, w):
        lib().zsb_multi_set_weights(self.h, (C.c_double * self.n)(*w) if w else None)

    def scan_decode(self, src, dst, flags=VERIFY_CHECKSUM, max_window=0):
        """-> ScanDecode-like object (frames, blocks, results, total, status)"""
        r = ScanDecode.__new__(ScanDecode)
        fp, bp, rp = C.POINTER(ZsbFrame)(), C.POINTER(ZsbBlock)(), C.POINTER(ZsbResult)()
        nf, nb, tot, ea, eb = C.c_size_t(), C.c_size_t(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        r.status = lib().zsb_multi_scan_decode(self.h, C.c_void_p(src[0]), src[1], C.c_void_p(dst[0]), dst[1], flags, max_window,
                                               C.byref(fp), C.byref(nf), C.byref(bp), C.byref(nb), C.byref(rp), C.byref(tot), C.byref(ea), C.byref(eb))
        r.frames, r.blocks, r.results = fp, bp, rp
        r.n_frames, r.n_blocks, r.total = nf.value, nb.value, tot.value
        r.err_a, r.err_b = ea.value, eb.value
        if not rp and r.status:
            raise ZsbError(r.status, what="zsb_multi_scan_decode: " + lib().zsb_multi_last_error(self.h).decode())
        return r

    def close(self):
        if getattr(self, "h", None):
            lib().zsb_multi_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_peer(slabs, dst_device, dst_ptr):
    """slabs: [(device, device pointer, bytes)] -> milliseconds of the NVLink gather into dst_ptr on dst_device (zsb_gather_peer)"""
    n = len(slabs)
    devs = (C.c_int * n)(*[s[0] for s in slabs]); ptrs = (C.c_void_p * n)(*[s[1] for s in slabs]); sizes = (C.c_size_t * n)(*[s[2] for s in slabs])
    ms = C.c_float()
    rc = lib().zsb_gather_peer(n, devs, ptrs, sizes, dst_device, C.c_void_p(dst_ptr), C.byref(ms))
    if rc:
        raise ZsbError(rc, what="zsb_gather_peer")
    return ms.value


class Decoder:
    """Batch decode of every frame of a buffer (== Frame::decode for each frame of the iterator)."""
    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()

    def _check(self, rc, what):
        if rc:
            raise ZsbError(rc, what=what + (": " + self.ctx.cuda_error() if rc == E_CUDA else ""))

    def decode(self, data, flags=VERIFY_CHECKSUM, dst_cap=None, scan=None):
        """Host buffers in, host bytes out.  Returns (output bytes, Scan, BatchResult)."""
        sc = scan or Scan(data, flags)
        cap = dst_cap if dst_cap is not None else capacity_bound(sc, flags)
        out = C.create_string_buffer(max(cap, 1))
        r = BatchResult(sc.n_frames)
        rc = lib().zsb_decode(self.ctx.h, sc.buf, sc.n, sc.frames, sc.n_frames, sc.blocks, sc.n_blocks, out, cap,
                              r.dst_off, r.dst_len, r.status, r.xxh32, r.checksum_ok, C.byref(r.total), flags)
        self._check(rc, "zsb_decode")
        return out.raw[:r.total.value], sc, r

    def scan_decode(self, data, flags=VERIFY_CHECKSUM, dst_cap=None):
        """decode() with the walk overlapped (zsb_scan_decode).  dst_cap: output capacity; default 4 x the input + 1 MiB, grown and
        retried when a frame reports that the output did not fit.  Returns (output bytes, ScanDecode)."""
        buf, n = _as_buffer(data)
        cap = dst_cap if dst_cap is not None else 4 * n + (1 << 20)
        while True:
            out = C.create_string_buffer(max(cap, 1))
            sd = ScanDecode(self.ctx, (C.cast(buf, C.c_void_p).value or 0, n), (C.addressof(out), cap), flags)
            sd._keep = (data, buf)
            err = sd.first_error()
            if dst_cap is None and err is not None and err[1] == E_DST_TOO_SMALL and cap < (1 << 40):
                cap *= 4
                continue
            return out.raw[:sd.total], sd

    # ---- resident path: device pointers, explicit prepare / launch / finish (used by bench.py)
    def prepare(self, src_ptr, n, scan, dst_ptr, dst_cap, flags):
        self._scan = scan
        self._check(lib().zsb_decode_prepare(self.ctx.h, C.c_void_p(src_ptr), n, scan.frames, scan.n_frames, scan.blocks, scan.n_blocks,
                                             C.c_void_p(dst_ptr), dst_cap, flags), "zsb_decode_prepare")

    def launch(self):
        self._check(lib().zsb_decode_launch(self.ctx.h), "zsb_decode_launch")

    def finish(self):
        r = BatchResult(self._scan.n_frames)
        self._check(lib().zsb_decode_finish(self.ctx.h, r.dst_off, r.dst_len, r.status, r.xxh32, r.checksum_ok, C.byref(r.total)), "zsb_decode_finish")
        return r


def capacity_bound(scan, flags=0):
    """Upper bound of the output size of a scanned buffer: content sizes where declared, else
    128 KiB per compressed block (Block_Maximum_Size) and the stored size of raw / RLE blocks."""
    cap = 0
    for i in range(scan.n_frames):
        f = scan.frames[i]
        if f.kind == 1:
            cap += scan.blocks[f.first_block].size if f.n_blocks else 0
        elif f.has_content_size and not (flags & REFERENCE_QUIRKS):
            cap += f.content_size
        else:
            for k in range(f.n_blocks):
                b = scan.blocks[f.first_block + k]
                cap += 131072 if b.type == 2 else b.size
    return cap


# ------------------------------------------------------------------------------------------ reference API mirror
class Header:
    """frame.rs:103-108"""
    def __init__(self, f):
        self.content_checksum_flag = bool(f.has_checksum)
        self.window_size = f.window_size
        self.dictionnary_id = f.dict_id if f.has_dict_id else None
        self.content_size = f.content_size if f.has_content_size else None


class Block:
    """enum Block (block.rs:29-40): type and extent; section parsing happens on the GPU."""
    RAW, RLE, COMPRESSED = 0, 1, 2
    def __init__(self, b, data):
        self.type, self.last, self.size = b.type, bool(b.last), b.size
        self.src_off = b.src_off
        self._data = data


class Frame:
    """enum Frame { ZStandardFrame(ZStandard), SkippableFrame(Skippable) }  frame.rs:47-56"""
    def __init__(self, parser, index):
        self._p, self._i = parser, index
        f = parser._scan.frames[index]
        self.is_skippable = f.kind == 1
        self.magic = f.magic
        self.src_off, self.src_len = f.src_off, f.src_len
        if self.is_skippable:
            b = parser._scan.blocks[f.first_block]
            self.data = bytes(parser._bytes[b.src_off:b.src_off + b.size])

    # ZStandard accessors frame.rs:262-272
    def header(self):
        return Header(self._p._scan.frames[self._i])

    def checksum(self):
        f = self._p._scan.frames[self._i]
        return f.stored_checksum if f.has_checksum else None

    def blocks(self):
        f = self._p._scan.frames[self._i]
        return [Block(self._p._scan.blocks[f.first_block + k], self._p._bytes) for k in range(f.n_blocks)]

    def decode(self, ctx=None):
        """Frame::decode frame.rs:79-84 -> decoded bytes (skippable: its data).  Raises ZsbError."""
        if self.is_skippable:
            return self.data
        f = self._p._scan.frames[self._i]
        sub = self._p._bytes[f.src_off:f.src_off + f.src_len]
        out, sc, r = Decoder(ctx).decode(sub, self._p._flags | VERIFY_CHECKSUM)
        if sc.status:
            raise sc.error()
        if r.status[0]:
            raise ZsbError(r.status[0])
        self.computed_checksum = r.xxh32[0] if f.has_checksum else None
        self.checksum_ok = bool(r.checksum_ok[0]) if f.has_checksum else None
        return out


class FrameIterator:
    """frame.rs:87-100: yields Frame; raises the frame's error where the reference yields Err."""
    def __init__(self, parser):
        self._p, self._i = parser, 0

    def __iter__(self):
        return self

    def __next__(self):
        sc = self._p._scan
        if self._i >= sc.n_frames:
            raise StopIteration
        f = sc.frames[self._i]
        if f.status:
            self._i = sc.n_frames
            raise ZsbError(f.status, sc.err_a, sc.err_b)
        fr = Frame(self._p, self._i)
        self._i += 1
        return fr


class ForwardByteParser:
    """parsing.rs:9,29-36.  `quirks=True` keeps the reference's accept/reject behaviour."""
    def __init__(self, data, quirks=True):
        self._bytes = bytes(data)
        self._flags = REFERENCE_QUIRKS if quirks else 0
        self._scan = Scan(self._bytes, self._flags)

    def iter(self):
        return FrameIterator(self)

    def len(self):
        return len(self._bytes)


class DecodingContext:
    """decoding_context.rs:17-106 for the stage the reference's tests call directly."""
    def __init__(self, window_size, ctx=None):
        if window_size > MAX_WIN_SIZE:
            raise ZsbError(40, MAX_WIN_SIZE, window_size)
        self.window_size = window_size
        self.decoded = b""
        self._ctx = ctx or default_context()

    def execute_sequences(self, sequences, literals):
        """sequences: iterable of (literal_length, offset_value, match_length)."""
        seqs = list(sequences)
        flat = (C.c_uint32 * max(3 * len(seqs), 1))(*[int(x) for s in seqs for x in s])
        literals = bytes(literals)
        out = C.create_string_buffer(131072 + 64); ol = C.c_size_t()
        rc = lib().zsb_execute_sequences(self._ctx.h, flat, len(seqs), literals, len(literals), out, 131072, C.byref(ol))
        if rc:
            raise ZsbError(rc)
        self.decoded += out.raw[:ol.value]


def fse_table_parse(desc, ctx=None, max_symbols=64):
    """parse_fse_table + FseTable::from_distribution (fse.rs:16-69,110-202) on the GPU.
    -> (accuracy_log, distribution, [(output, baseline, bits_to_read)], bytes_read)"""
    ctx = ctx or default_context(); desc = bytes(desc)
    al = C.c_uint8(); tbl = (C.c_uint16 * (3 * 512))(); cons = C.c_size_t(); dist = (C.c_int16 * 256)(); nd = C.c_size_t()
    rc = lib().zsb_fse_table_parse(ctx.h, desc, len(desc), max_symbols, C.byref(al), tbl, C.byref(cons), dist, C.byref(nd))
    if rc:
        raise ZsbError(rc)
    return al.value, list(dist[:nd.value]), [(tbl[3 * i], tbl[3 * i + 1], tbl[3 * i + 2]) for i in range(1 << al.value)], cons.value


def fse_table_from_distribution(al, dist, ctx=None):
    """FseTable::from_distribution (fse.rs:110-202) on the GPU."""
    ctx = ctx or default_context()
    d = (C.c_int16 * len(dist))(*dist); tbl = (C.c_uint16 * (3 * 512))()
    rc = lib().zsb_fse_table_from_distribution(ctx.h, al, d, len(dist), tbl)
    if rc:
        raise ZsbError(rc)
    return [(tbl[3 * i], tbl[3 * i + 1], tbl[3 * i + 2]) for i in range(1 << al)]


def huffman_parse(desc, ctx=None):
    """HuffmanDecoder::parse (huffman.rs:80-203) on the GPU -> ({symbol: (nbits, code)}, consumed, max_bits)."""
    ctx = ctx or default_context(); desc = bytes(desc)
    lens = (C.c_uint8 * 256)(); codes = (C.c_uint16 * 256)(); cons = C.c_size_t(); mb = C.c_uint8()
    rc = lib().zsb_huffman_parse(ctx.h, desc, len(desc), lens, codes, C.byref(cons), C.byref(mb))
    if rc:
        raise ZsbError(rc)
    return {s: (lens[s], codes[s]) for s in range(256) if lens[s]}, cons.value, mb.value


def xxh64(data, ctx=None):
    ctx = ctx or default_context(); data = bytes(data); h = C.c_uint64()
    rc = lib().zsb_xxh64(ctx.h, data, len(data), C.byref(h))
    if rc:
        raise ZsbError(rc)
    return h.value


def decompress(data, print_skippable=False, quirks=True, verify=True, ctx=None):
    """src/main.rs:42-58: decode every frame and concatenate; any error aborts with no output."""
    flags = (PRINT_SKIPPABLE if print_skippable else 0) | (REFERENCE_QUIRKS if quirks else 0) | (VERIFY_CHECKSUM if verify else 0)
    out, sc, r = Decoder(ctx).decode(data, flags)
    if sc.status:
        # the reference reports the first error in stream order: a frame decoded before the bad one may fail first
        e = r.first_error()
        if e and e[0] < sc.n_frames - 1:
            raise ZsbError(e[1])
        raise sc.error()
    e = r.first_error()
    if e:
        raise ZsbError(e[1])
    return out


# ------------------------------------------------------------------------------------------ sharding by frame (one process per GPU)
def shard_plan(scan, n_shards):
    """Contiguous frame ranges balanced on decompressed bytes: [first[s], first[s+1]) for shard s (zsb_shard_plan)."""
    first = (C.c_size_t * (n_shards + 1))()
    rc = lib().zsb_shard_plan(scan.frames, scan.n_frames, n_shards, first)
    if rc:
        raise ZsbError(rc)
    return list(first)


class Shard:
    """Frames [f0, f1) of a scanned buffer as a self-contained batch: the sub-buffer and its rebased descriptors."""
    def __init__(self, scan, f0, f1):
        fp, bp = C.POINTER(ZsbFrame)(), C.POINTER(ZsbBlock)()
        nb, off, ln = C.c_size_t(), C.c_uint64(), C.c_uint64()
        rc = lib().zsb_shard_extract(scan.frames, scan.n_frames, scan.blocks, scan.n_blocks, f0, f1, C.byref(fp), C.byref(bp), C.byref(nb), C.byref(off), C.byref(ln))
        if rc:
            raise ZsbError(rc)
        self.frames, self.blocks, self.n_frames, self.n_blocks = fp, bp, f1 - f0, nb.value
        self.src_off, self.src_len, self.f0, self.f1 = off.value, ln.value, f0, f1
        self.status = 0
        base = C.cast(scan.buf, C.c_void_p).value or 0
        self.buf, self.n = C.c_void_p(base + self.src_off), self.src_len
        self._keep = scan

    def __del__(self):
        try:
            if getattr(self, "frames", None): lib().zsb_free(self.frames)
            if getattr(self, "blocks", None): lib().zsb_free(self.blocks)
        except Exception:
            pass


def decode_shard(data, rank, world, flags=VERIFY_CHECKSUM, ctx=None, scan=None):
    """What rank `rank` of `world` does with a buffer every rank holds: scan (host), take its frame range, decode it on
    its own GPU.  Returns (output bytes of the shard, Shard, BatchResult).  No collective is involved."""
    sc = scan or Scan(data, flags)
    first = shard_plan(sc, world)
    sh = Shard(sc, first[rank], first[rank + 1])
    out, _, r = Decoder(ctx).decode(None, flags, scan=sh)
    return out, sh, r


def gather_outputs(local_bytes, group=None, device=None):
    """Optional gather for a single-stream consumer: every rank receives the concatenation of all shards' outputs in rank
    order (torch.distributed all_gather over NCCL/NVLink on GPUs, gloo on CPU).  Not part of the decode path."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device or "cpu"
    n = torch.tensor([len(local_bytes)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if len(local_bytes):
        mine[:len(local_bytes)] = torch.frombuffer(bytearray(local_bytes), dtype=torch.uint8).to(dev)
    parts = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return b"".join(bytes(p[:sz].cpu().numpy().tobytes()) for p, sz in zip(parts, sizes))
