"""Per-CTA timeline of k_huf (library built with -DZSB_SEQ_TIMING, named by ZSB_LIB_PATH): table set-up against stream decode."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gen_corpus as G
import zstd_decompressor_b200 as Z
blob, exp = G.make_c2(4096, seed=2)
ctx = Z.Context(0); dec = Z.Decoder(ctx)
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
src = torch.cat([torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda(), torch.zeros(128, dtype=torch.uint8, device="cuda")])
dst = torch.empty(len(exp) + 64, dtype=torch.uint8, device="cuda")
sc = Z.Scan(blob, 6)
dec.prepare(src.data_ptr(), len(blob), sc, dst.data_ptr(), len(exp), 6 | 8 | 16)
for _ in range(3):
    dec.launch(); r = dec.finish()
buf = (C.c_longlong * (1024 * 8))()
assert Z.lib().zsb_debug_huf_timing(buf) == 0
rows = [r for r in ([buf[i * 8 + j] for j in range(8)] for i in range(1024)) if r[0]]      # (CTAs that ran)
t0 = min(r[0] for r in rows)
import statistics as S
for name, f in (("start - first start", lambda r: r[0] - t0), ("tables", lambda r: r[1] - r[0]), ("  weights", lambda r: r[3] - r[0]), ("  lut", lambda r: r[1] - r[3]), ("    plan", lambda r: r[4] - r[3]), ("    cell starts", lambda r: r[5] - r[4]), ("    T1", lambda r: r[6] - r[5]), ("    pairs", lambda r: r[1] - r[6]), ("streams", lambda r: r[2] - r[1]), ("end - first start", lambda r: r[2] - t0)):
    c = [f(r) for r in rows]
    print(f"{name:24s} min {min(c):10d} median {S.median(c):12.1f} max {max(c):10d}")
