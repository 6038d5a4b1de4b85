"""zsb_scan_device (the walk on the GPU over a buffer resident in HBM) beside the alternatives for such a buffer: zsb_scan on a host copy that
   already exists, and the device-to-host copy that the host walk would need first.  Wall clock of the calls, best of 5.
   python tools/probes/dscan_timing.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gen_corpus as G
import zstd_decompressor_b200 as Z

ctx = Z.Context(0)


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def run(name, blob):
    src = torch.cat([torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda(), torch.zeros(256, dtype=torch.uint8, device="cuda")])
    pin = torch.empty(len(blob), dtype=torch.uint8, pin_memory=True)
    ds = Z.DeviceScan(ctx, src.data_ptr(), len(blob), 4)
    hs = Z.Scan(blob, 4)
    same = (ds.status, ds.n_frames, ds.n_blocks) == (hs.status, hs.n_frames, hs.n_blocks)
    t_dev = best(lambda: Z.DeviceScan(ctx, src.data_ptr(), len(blob), 4))
    t_host = best(lambda: Z.Scan(blob, 4))
    t_d2h = best(lambda: pin.copy_(src[:len(blob)], non_blocking=True))
    print(f"{name}: {len(blob) / 1e6:.1f} MB compressed, {hs.n_frames} frames, {hs.n_blocks} blocks, {'same' if same else 'DIFFERENT'}: "
          f"zsb_scan_device {t_dev:.3f} ms ({len(blob) / t_dev / 1e6:.1f} GB/s of compressed bytes) | zsb_scan on a host copy {t_host:.3f} ms | D2H of the buffer {t_d2h:.3f} ms", flush=True)


run("C2 (4 096 frames of 128 KiB)", G.make_c2(4096, seed=2)[0])
run("C5 share (8 192 frames)", G.make_c2(8192, seed=5)[0])
run("C3 (one frame, 256 MiB)", G.make_c3(total=256 << 20)[0])
run("C4 (mixed)", G.make_c4()[0])
