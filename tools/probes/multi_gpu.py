"""One process, N GPUs: (1) zsb_multi_scan_decode of one C5-style corpus from one pinned host buffer (equal and calibrated shard weights),
(2) per-device resident decode of the same shards followed by the NVLink gather into device 0 (zsb_gather_peer), timed apart.
    gpurun --gpus N -- python tools/probes/multi_gpu.py [frames_total]
"""
import ctypes as C, hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gen_corpus as G
import zstd_decompressor_b200 as Z

n_dev = torch.cuda.device_count()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8192 * n_dev
blob, plain = G.make_c2(frames, seed=5)
want = hashlib.sha256(plain).digest()
L = Z.lib()
src = L.zsb_host_alloc(len(blob)); dst = L.zsb_host_alloc(len(plain) + 64)
C.memmove(src, blob, len(blob))
flags = Z.VERIFY_CHECKSUM | Z.REFERENCE_QUIRKS
out = {"devices": n_dev, "frames": frames, "compressed": len(blob), "decompressed": len(plain)}
m = Z.MultiContext(list(range(n_dev)))


def run(label, reps=4):
    for _ in range(2):
        r = m.scan_decode((src, len(blob)), (dst, len(plain)), flags)
        assert r.status == 0 and r.total == len(plain) and r.first_error() is None, (r.status, r.total, r.first_error())
    assert hashlib.sha256(memoryview((C.c_uint8 * len(plain)).from_address(dst))).digest() == want
    t = time.perf_counter()
    for _ in range(reps):
        r = m.scan_decode((src, len(blob)), (dst, len(plain)), flags)
    dt = (time.perf_counter() - t) / reps
    out[label] = {"ms": dt * 1e3, "GBps": len(plain) / dt / 1e9, "weights": m.weights()}


run("equal_shards")
gbs = m.calibrate()
out["link_GBps_all_busy"] = gbs
run("weighted_shards")

# ---- resident decode per device + NVLink gather to device 0
first = (C.c_size_t * (n_dev + 1))()
sc = Z.Scan(blob, flags)
assert L.zsb_shard_plan(sc.frames, sc.n_frames, n_dev, first) == 0
slabs, keep = [], []
for d in range(n_dev):
    f0, f1 = first[d], first[d + 1]
    fp, bp = C.POINTER(Z.ZsbFrame)(), C.POINTER(Z.ZsbBlock)()
    nb, so, sl = C.c_size_t(), C.c_uint64(), C.c_uint64()
    assert L.zsb_shard_extract(sc.frames, sc.n_frames, sc.blocks, sc.n_blocks, f0, f1, C.byref(fp), C.byref(bp), C.byref(nb), C.byref(so), C.byref(sl)) == 0
    n_out = (f1 - f0) * 131072
    with torch.cuda.device(d):
        d_src = torch.empty(sl.value + 128, dtype=torch.uint8, device=f"cuda:{d}")
        d_src[:sl.value].copy_(torch.frombuffer(bytearray(blob[so.value:so.value + sl.value]), dtype=torch.uint8)); d_src[sl.value:].zero_()
        d_dst = torch.empty(n_out + 64, dtype=torch.uint8, device=f"cuda:{d}")
        torch.cuda.synchronize(d)
    ctx = L.zsb_multi_ctx(m.h, d)
    tot = C.c_uint64()
    st = (C.c_int32 * (f1 - f0))()
    rc = L.zsb_decode(ctx, C.c_void_p(d_src.data_ptr()), sl.value, fp, f1 - f0, bp, nb.value, C.c_void_p(d_dst.data_ptr()), n_out, None, None, st, None, None, C.byref(tot),
                      flags | Z.SRC_ON_DEVICE | Z.DST_ON_DEVICE)
    assert rc == 0 and tot.value == n_out and not any(st), (rc, tot.value)
    slabs.append((d, d_dst.data_ptr(), n_out)); keep.append((d_src, d_dst))
with torch.cuda.device(0):
    g = torch.empty(len(plain) + 64, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize(0)
times = [Z.gather_peer(slabs, 0, g.data_ptr()) for _ in range(5)]
assert hashlib.sha256(g[:len(plain)].cpu().numpy().tobytes()).digest() == want
peer_bytes = sum(s[2] for s in slabs[1:])
out["nvlink_gather_to_device0"] = {"ms_best": min(times), "ms_all": times, "bytes_over_nvlink": peer_bytes, "GBps_into_device0": peer_bytes / (min(times) * 1e-3) / 1e9 if peer_bytes else None,
                                   "note": "one cudaMemcpyPeerAsync per slab; the slab of device 0 is a local copy; never part of a decode figure"}
print(json.dumps(out))
