"""C3-style single frame through zsb_scan_decode on page-locked host buffers: wall time per call against the kernels' time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gen_corpus as G
import zstd_decompressor_b200 as Z
total = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 6
blob, exp = G.make_c3(total=total)
ctx = Z.Context(0)
hs = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
hd = torch.empty(len(exp) + 64, dtype=torch.uint8).pin_memory()
ctx.set_profile(True)
for it in range(4):
    t = time.perf_counter()
    sd = Z.ScanDecode(ctx, (hs.data_ptr(), len(blob)), (hd.data_ptr(), len(exp)), flags)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) * 1e3
    kt = ctx.kernel_times()
    print(it, "status", sd.status, "total", sd.total, f"{dt:.1f} ms wall, kernels {sum(v for _, v in kt):.1f} ms", {k: round(v, 1) for k, v in kt if v > 1}, "launches", ctx.last_launch_count())
print("ok" if hd[:len(exp)].numpy().tobytes() == exp else "WRONG")
