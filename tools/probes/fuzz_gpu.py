"""Differential fuzz on a GPU box: mutated frames through the CUDA path and through the g++ build of the same lane-serial
code (tests/emul); statuses must agree frame by frame and decoded frames must be identical (development probe; the
permanent, smaller version is tests/test_gpu_parity.py::test_gpu_matches_cpu_build_of_device_code)."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpora
import emul_lib as E
import zstd_decompressor_b200 as Z

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 77
dec = Z.Decoder(Z.Context(0))
r = random.Random(seed)
srcs = dict(corpora.mutation_sources())
blob, _ = corpora.c2_small(3)
srcs["c2x3"] = blob
srcs["moby"] = corpora.fixture("moby-dick.txt.zst")
names = list(srcs)
bad = ok = err = 0
t0 = time.time()
for it in range(n_iter):
    name = r.choice(names)
    b = corpora.mutate(r, srcs[name])
    if r.random() < 0.3:                       # several frames in one batch, some of them damaged
        b = b + corpora.mutate(r, srcs[r.choice(names)]) + srcs[r.choice(names)]
    flags = 4 | 1 | (2 if r.random() < 0.5 else 0)
    sc0 = Z.Scan(b, flags)
    cap = Z.capacity_bound(sc0, flags)                 # the same capacity for both sides (ZSB_E_DST_TOO_SMALL depends on it)
    rc, eout, eframes, _ = E.decode(b, flags & ~2, cap=cap)
    out, sc, res = dec.decode(b, flags, dst_cap=cap, scan=sc0)
    same = sc.status == rc and [res.status[i] for i in range(sc.n_frames)] == [f[0] for f in eframes]
    for i, (st, off, ln) in enumerate(eframes):
        if same and st == 0:
            same = (res.dst_off[i], res.dst_len[i]) == (off, ln) and out[off:off + ln] == eout[off:off + ln]
    if not same:
        bad += 1
        if bad <= 5:
            open(os.path.join(ROOT, "gpurun_out", f"fuzz_fail_{seed}_{it}.zst"), "wb").write(b)
            print("MISMATCH", it, name, sc.status, rc, [res.status[i] for i in range(sc.n_frames)], [f[0] for f in eframes])
    elif any(f[0] for f in eframes) or rc:
        err += 1
    else:
        ok += 1
print(f"{n_iter} inputs in {time.time() - t0:.1f} s: {ok} decoded identically, {err} rejected identically, {bad} MISMATCHES")
