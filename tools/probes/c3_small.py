"""C3-style single frame (default 64 MiB), resident decode, per-kernel times (development probe; ncu target for k_exec)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gen_corpus as G
import zstd_decompressor_b200 as Z
total = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 6
blob, exp = G.make_c3(total=total)
ctx = Z.Context(0); dec = Z.Decoder(ctx)
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
src = torch.cat([torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda(), torch.zeros(128, dtype=torch.uint8, device="cuda")])
dst = torch.empty(len(exp) + 64, dtype=torch.uint8, device="cuda")
sc = Z.Scan(blob, flags)
dec.prepare(src.data_ptr(), len(blob), sc, dst.data_ptr(), len(exp), flags | 8 | 16)
dec.launch(); r = dec.finish()
ok = r.first_error() is None and dst[:len(exp)].cpu().numpy().tobytes() == exp
ctx.set_profile(True)
for _ in range(2): dec.launch()
t, n = ctx.kernel_times_avg()
tot = sum(v for _, v in t)
print("ok" if ok else "WRONG", round(tot, 3), "ms", round(len(exp) / tot / 1e6, 3), "GB/s", {k: round(v, 3) for k, v in t if v > 0.01})
