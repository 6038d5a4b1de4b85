"""Per-CTA timeline of k_seq (library built with -DZSB_SEQ_TIMING, named by ZSB_LIB_PATH): where the sequence stage spends its time."""
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
    dec.launch(); r = dec.finish() if not os.environ.get("NOFIN") else None
buf = (C.c_longlong * (160 * 8))()
assert Z.lib().zsb_debug_seq_timing(buf) == 0
rows = [[buf[i * 8 + j] for j in range(8)] for i in range(147)]
t0 = min(r[0] for r in rows)
import statistics as S
def col(f): return [f(r) for r in rows]
for name, f in (("start - first start", lambda r: r[0] - t0), ("table build", lambda r: r[1] - r[0]), ("state init", lambda r: r[2] - r[1]), ("producer loop", lambda r: r[3] - r[2]),
                ("producer waited for free windows", lambda r: r[4]), ("helper 1 waited", lambda r: r[5]), ("helper 1 worked", lambda r: r[6]), ("windows", lambda r: r[7]),
                ("loop cycles per step (windows*128)", lambda r: (r[3] - r[2] - r[4]) / max(r[7] * 128, 1)), ("end - first start", lambda r: r[3] - t0)):
    c = col(f)
    print(f"{name:42s} min {min(c):12.1f} median {S.median(c):12.1f} max {max(c):12.1f}")
