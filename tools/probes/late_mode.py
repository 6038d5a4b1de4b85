"""zsb_scan_decode on C2 with and without Frame_Content_Size (early vs late placement of the shards): ms per call.
   python tools/probes/late_mode.py [frames]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import random
from concurrent.futures import ThreadPoolExecutor
import gen_corpus as G
import zstd_decompressor_b200 as Z
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
text = G.moby_text(); r = random.Random(3)
offs = [r.randrange(len(text) - 131072) for _ in range(n)]
def comp(fcs):
    def work(idx):
        c = G.Compressor(level=3, content_size=fcs)
        return [c.compress(text[offs[i]:offs[i] + 131072]) for i in idx]
    with ThreadPoolExecutor(16) as ex:
        parts = list(ex.map(work, [range(i, n, 16) for i in range(16)]))
    out = [None] * n
    for t, p in enumerate(parts):
        for j, i in enumerate(range(t, n, 16)): out[i] = p[j]
    return b"".join(out)
want_len = n * 131072
dec = Z.Decoder(Z.Context(0)); L = Z.lib()
for fcs in (True, False):
    blob = comp(fcs)
    src = L.zsb_host_alloc(len(blob)); dst = L.zsb_host_alloc(want_len + 64)
    C.memmove(src, blob, len(blob))
    for _ in range(3): sd = Z.ScanDecode(dec.ctx, (src, len(blob)), (dst, want_len + 64), Z.VERIFY_CHECKSUM)
    assert sd.status == 0 and sd.first_error() is None and sd.total == want_len
    t = time.perf_counter()
    for _ in range(10): sd = Z.ScanDecode(dec.ctx, (src, len(blob)), (dst, want_len + 64), Z.VERIFY_CHECKSUM)
    dt = (time.perf_counter() - t) / 10
    ok = C.string_at(dst, want_len) == b"".join(text[o:o + 131072] for o in offs)
    print(f"content size declared: {fcs}: {dt * 1e3:.2f} ms per call, {want_len / dt / 1e9:.1f} GB/s, output {'ok' if ok else 'WRONG'}", flush=True)
    L.zsb_host_free(src); L.zsb_host_free(dst)
