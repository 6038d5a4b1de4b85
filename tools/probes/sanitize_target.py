"""A small tour of the decode path (asserts every output; also the target for compute-sanitizer where the pool allows it): the fixtures, C4, 64 C2 frames, a 3 MiB C3 frame,
the same with ZSB_SEQX=1 and ZSB_WAVE=4, a host-buffer call and a few mutated inputs.
    compute-sanitizer --tool memcheck python tools/probes/sanitize_target.py"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import corpora
import zstd_decompressor_b200 as Z
Q, SKIP, VER = Z.REFERENCE_QUIRKS, Z.PRINT_SKIPPABLE, Z.VERIFY_CHECKSUM


def tour(dec, tag):
    n = 0
    for name in corpora.FIXTURE_NAMES:
        d = corpora.fixture(name)
        for fl in (Q | VER, VER | SKIP):
            out, sc, r = dec.decode(d, fl); n += 1
    blob, exp, exp_skip, _ = corpora.c4()
    out, sc, r = dec.decode(blob, Q | VER); assert out == exp; n += 1
    blob, exp = corpora.c2_small(64)
    out, sc, r = dec.decode(blob, Q | VER); assert out == exp; n += 1
    blob, exp = corpora.c3_small(3 << 20)
    out, sc, r = dec.decode(blob, Q | VER); assert out == exp; n += 1
    rnd = random.Random(7)
    for src in list(corpora.mutation_sources().values())[:3]:
        for _ in range(6):
            d = corpora.mutate(rnd, src)
            dec.decode(d, Q | VER); n += 1
    # the walk on the device (zsb_scan_device) on the same inputs, truncations and mutations
    import torch
    for d in [corpora.fixture(x) for x in corpora.FIXTURE_NAMES] + [corpora.c4()[0], corpora.c2_small(64)[0][:300_000], b"", b"abc", bytes(range(256))] + \
             [corpora.mutate(rnd, src) for src in list(corpora.mutation_sources().values())[:4] for _ in range(4)]:
        for cut in (len(d), len(d) // 2, max(len(d) - 3, 0)):
            t = torch.zeros(cut + 256, dtype=torch.uint8, device="cuda:0")
            if cut: t[:cut] = torch.frombuffer(bytearray(d[:cut]), dtype=torch.uint8).to("cuda:0")
            torch.cuda.synchronize()
            ds = Z.DeviceScan(dec.ctx, t.data_ptr(), cut, Q); hs = Z.Scan(d[:cut], Q)
            assert (ds.status, ds.n_frames, ds.n_blocks, ds.err_a, ds.err_b) == (hs.status, hs.n_frames, hs.n_blocks, hs.err_a, hs.err_b); n += 1
    print(tag, n, "decodes and walks")


tour(Z.Decoder(Z.Context(0)), "default")
for k, v in (("ZSB_SEQX", "1"), ("ZSB_WAVE", "4")):
    os.environ[k] = v
    try:
        d = Z.Decoder(Z.Context(0))
    finally:
        del os.environ[k]
    tour(d, f"{k}={v}")
print("done")
