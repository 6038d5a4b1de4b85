"""Per-kernel CUDA-event times of one frame of many blocks (C3-like, resident buffers) for the library named by ZSB_LIB_PATH.
   python tools/probes/ktimes_c3.py [MiB]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gen_corpus as G
import zstd_decompressor_b200 as Z

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
blob, exp = G.make_c3(total=mib << 20)
ctx = Z.Context(0); dec = Z.Decoder(ctx)
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
src = torch.cat([torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda(), torch.zeros(64, dtype=torch.uint8, device="cuda")])
dst = torch.empty(len(exp) + 64, dtype=torch.uint8, device="cuda")
sc = Z.Scan(blob, 6)
dec.prepare(src.data_ptr(), len(blob), sc, dst.data_ptr(), len(exp), 6 | 8 | 16)
dec.launch(); r = dec.finish()
ok = r.first_error() is None and dst[:len(exp)].cpu().numpy().tobytes() == exp
ctx.set_profile(True)
for _ in range(3): dec.launch()
t, n = ctx.kernel_times_avg()
tot = sum(v for _, v in t)
print(os.environ.get("ZSB_LIB_PATH", "default"), "ok" if ok else "WRONG", f"{sc.n_frames} frame(s), {sc.n_blocks} blocks, {tot:.3f} ms = {len(exp) / tot / 1e6:.2f} GB/s", {k: round(v, 4) for k, v in t if v > 0.01})
