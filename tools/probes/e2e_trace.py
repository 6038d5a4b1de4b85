"""C2 through zsb_scan_decode on page-locked buffers with ZSB_PIPE_TRACE=1: the device timeline of every shard (stderr), wall time per call."""
import os, sys, time
os.environ["ZSB_PIPE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gen_corpus as G
import zstd_decompressor_b200 as Z
blob, exp = G.make_c2(4096, seed=2)
ctx = Z.Context(0)
hs = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
hd = torch.empty(len(exp) + 64, dtype=torch.uint8).pin_memory()
fl = Z.VERIFY_CHECKSUM | Z.REFERENCE_QUIRKS
for it in range(5):
    t = time.perf_counter()
    sd = Z.ScanDecode(ctx, (hs.data_ptr(), len(blob)), (hd.data_ptr(), len(exp)), fl)
    dt = (time.perf_counter() - t) * 1e3
    print(f"call {it}: {dt:.2f} ms", file=sys.stderr)
