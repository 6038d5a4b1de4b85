"""Randomised check of the pipelined host paths on a GPU box: batches of frames of random sizes and kinds (empty, tiny, multi-block,
incompressible -> raw blocks, runs -> RLE blocks, skippable, with and without content size / checksum), decoded by zsb_decode and
zsb_scan_decode on page-locked buffers and by zsb_decode on pageable ones; all three must equal the plaintext.
   python tools/probes/fuzz_pipe.py [iterations] [seed]"""
import ctypes as C, os, random, struct, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen_corpus as G
import zstd_decompressor_b200 as Z

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 6
r = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
dec = Z.Decoder(Z.Context(0)); L = Z.lib()
text = G.moby_text()
Q, SKIP, VER = Z.REFERENCE_QUIRKS, Z.PRINT_SKIPPABLE, Z.VERIFY_CHECKSUM

def plain_of(kind):
    if kind == 0: n = r.choice([1, 2, 7, 100, 1000]); o = r.randrange(len(text) - n); return text[o:o + n]
    if kind == 1: n = r.randrange(1 << 10, 1 << 17); o = r.randrange(len(text) - n); return text[o:o + n]
    if kind == 2: n = r.randrange(1 << 17, 1 << 20); o = r.randrange(len(text) - n); return text[o:o + n]          # several blocks
    if kind == 3: return r.randbytes(r.randrange(1 << 8, 1 << 16))                                                    # raw blocks
    if kind == 4: return bytes([r.randrange(256)]) * r.randrange(1 << 8, 1 << 18)                                     # RLE
    n = r.randrange(1 << 14, 1 << 16); o = r.randrange(len(text) - n); return text[o:o + n]

bad = 0
t0 = time.time()
for it in range(n_iter):
    fcs_all = r.random() < 0.7
    frames, plains = [], []
    total_c = 0
    target = r.choice([17 << 20, 24 << 20, 40 << 20])
    while total_c < target:
        if r.random() < 0.03:
            pay = r.randbytes(r.randrange(1, 300)); f = struct.pack("<II", 0x184D2A50 + r.randrange(16), len(pay)) + pay
            frames.append(f); plains.append((pay, True)); total_c += len(f); continue
        p = plain_of(r.choice([0, 1, 1, 1, 2, 3, 4, 5, 5]))
        f = G.compress(p, level=r.choice([1, 3, 3, 5]), checksum=r.random() < 0.8, content_size=fcs_all or r.random() < 0.9)
        frames.append(f); plains.append((p, False)); total_c += len(f)
    blob = b"".join(frames)
    for flags in (VER | SKIP, VER):          # RFC mode: the reference's quirks reject some tiny valid frames
        want = b"".join(p for p, s in plains if (flags & SKIP) or not s)
        cap = len(want) + 64
        src = L.zsb_host_alloc(len(blob)); dst = L.zsb_host_alloc(cap)
        C.memmove(src, blob, len(blob))
        sc = Z.Scan((src, len(blob)), flags)
        res = Z.BatchResult(sc.n_frames)
        rc = L.zsb_decode(dec.ctx.h, C.c_void_p(src), len(blob), sc.frames, sc.n_frames, sc.blocks, sc.n_blocks, C.c_void_p(dst), cap,
                          res.dst_off, res.dst_len, res.status, res.xxh32, res.checksum_ok, C.byref(res.total), flags)
        ok1 = rc == 0 and res.first_error() is None and C.string_at(dst, res.total.value) == want
        C.memset(dst, 0, cap)
        sd = Z.ScanDecode(dec.ctx, (src, len(blob)), (dst, cap), flags)
        ok2 = sd.status == 0 and sd.first_error() is None and sd.n_frames == len(frames) and C.string_at(dst, sd.total) == want
        out3, sc3, r3 = dec.decode(blob, flags)
        ok3 = out3 == want
        L.zsb_host_free(src); L.zsb_host_free(dst)
        if not (ok1 and ok2 and ok3):
            bad += 1
            print(f"MISMATCH it={it} flags={flags} frames={len(frames)} fcs_all={fcs_all}: decode={ok1} scan_decode={ok2} pageable={ok3}", flush=True)
    print(f"it {it}: {len(frames)} frames, {len(blob) >> 20} MiB in, fcs_all={fcs_all}", flush=True)
print(f"{n_iter} batches in {time.time() - t0:.1f} s: {bad} MISMATCHES")
