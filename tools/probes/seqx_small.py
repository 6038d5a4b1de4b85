"""C2-style batch (default 1 024 frames), resident decode: which sequence kernel ran, is the output right, per-kernel times
(development probe for k_seqx; ZSB_DEBUG=1 prints the host's placement decisions)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gen_corpus as G
import zstd_decompressor_b200 as Z
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 2
blob, exp = G.make_c2(n_frames=nf, seed=2)
ctx = Z.Context(0); dec = Z.Decoder(ctx)
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
src = torch.cat([torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda(), torch.zeros(128, dtype=torch.uint8, device="cuda")])
dst = torch.zeros(len(exp) + 64, dtype=torch.uint8, device="cuda")
sc = Z.Scan(blob, flags)
dec.prepare(src.data_ptr(), len(blob), sc, dst.data_ptr(), len(exp), flags | 8 | 16)
dec.launch(); r = dec.finish()
print("seqx state after the first run:", ctx.last_seqx_state(), "first error:", r.first_error())
dec.prepare(src.data_ptr(), len(blob), sc, dst.data_ptr(), len(exp), flags | 8 | 16)
dst.zero_()
dec.launch(); r = dec.finish()
out = dst[:len(exp)].cpu().numpy().tobytes()
ok = r.first_error() is None and out == exp
if not ok:
    import numpy as np
    a = np.frombuffer(out, dtype=np.uint8); b = np.frombuffer(exp, dtype=np.uint8)
    bad = np.nonzero(a != b)[0]
    print("mismatches:", len(bad), "first at", bad[:8], "frame", bad[:8] // 131072, "pos in frame", bad[:8] % 131072)
print("seqx state:", ctx.last_seqx_state())
ctx.set_profile(True)
for _ in range(3): dec.launch()
t, n = ctx.kernel_times_avg()
tot = sum(v for _, v in t)
print("ok" if ok else "WRONG", round(tot, 3), "ms", round(len(exp) / tot / 1e6, 3), "GB/s", {k: round(v, 3) for k, v in t if v > 0.004})
