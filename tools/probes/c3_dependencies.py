"""How parallel is the execution of a C3 frame?  (CPU analysis, no GPU.)

The sequences zstd -3 emits for the C3 plaintext (ZSTD_generateSequences: the same parser as the compressor behind make_c3) are replayed
as a dependency graph at byte granularity: every output byte gets the 'time' at which it can exist if a block may start as soon as the
scheduler allows and a match waits only for its source bytes.  Printed: match-offset histogram, and for a block executed beside its
predecessors the fraction of its sequences that (transitively) wait for bytes of the previous block -- the part a block-parallel
wavefront cannot start early."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_corpus as G

total = int(sys.argv[1]) if len(sys.argv) > 1 else 8 << 20
plain = G.shuffled_tiles(total)
z = G.libzstd()


class Seq(C.Structure):
    _fields_ = [("offset", C.c_uint), ("litLength", C.c_uint), ("matchLength", C.c_uint), ("rep", C.c_uint)]


z.ZSTD_createCCtx.restype = C.c_void_p
z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
z.ZSTD_generateSequences.restype = C.c_size_t
z.ZSTD_generateSequences.argtypes = [C.c_void_p, C.POINTER(Seq), C.c_size_t, C.c_char_p, C.c_size_t]
cctx = z.ZSTD_createCCtx()
z.ZSTD_CCtx_setParameter(cctx, 100, 3)      # compressionLevel
z.ZSTD_CCtx_setParameter(cctx, 101, 23)     # windowLog
cap = len(plain) // 3 + 1024
buf = (Seq * cap)()
n = z.ZSTD_generateSequences(cctx, buf, cap, plain, len(plain))
assert not z.ZSTD_isError(n), "generateSequences failed"
a = np.frombuffer(buf, dtype=np.uint32, count=4 * n).reshape(n, 4)
off, ll, ml = a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), a[:, 2].astype(np.int64)
real = ml > 0                                # (block delimiters have offset = matchLength = 0)
print(f"{len(plain) >> 20} MiB, {int(real.sum())} sequences, {ll.sum() / len(plain):.3f} of the bytes are literals, mean match {ml[real].mean():.1f} bytes")
edges = [0, 16, 256, 4096, 65536, 131072, 262144, 524288, 1 << 20, 2 << 20, 8 << 20]
h, _ = np.histogram(off[real], bins=edges)
hb, _ = np.histogram(off[real], bins=edges, weights=ml[real])
for i in range(len(h)):
    print(f"  offset {edges[i]:>8} .. {edges[i + 1]:>8}: {100 * h[i] / real.sum():5.1f} % of the matches, {100 * hb[i] / ml[real].sum():5.1f} % of the match bytes")

# transitive dependence on the previous block: dep[p] = 1 if byte p of the frame cannot exist before the block before its own is complete
BS = 131072
pos = 0
dep = np.zeros(len(plain), dtype=np.uint8)
wait_seq = np.zeros((len(plain) + BS - 1) // BS, dtype=np.int64)
n_seq = np.zeros_like(wait_seq)
for o, l, m in zip(off.tolist(), ll.tolist(), ml.tolist()):
    pos += l                                 # literals depend on nothing
    if m == 0:
        continue
    b0 = (pos // BS) * BS                    # start of the block this match is written in (matches do not straddle blocks in zstd's output)
    s = pos - o
    d = 0
    if s < b0:                               # source begins in an earlier block
        e = min(s + m, b0)
        if e > b0 - BS:                      # ... and touches the block right before
            d = 1
    if not d and s + m > b0:                 # source inside this block: inherits
        lo = max(s, b0)
        d = int(dep[lo:min(s + m, pos)].max()) if lo < pos else 0
    blk = pos // BS
    n_seq[blk] += 1; wait_seq[blk] += d
    if d:
        dep[pos:pos + m] = 1
    elif o < m:                              # overlapping copy of independent bytes stays independent
        pass
    pos += m
fr = wait_seq[1:-1] / np.maximum(n_seq[1:-1], 1)
print(f"blocks: {len(fr)}; sequences that wait for the previous block (directly or through bytes that do): median {100 * np.median(fr):.1f} %, "
      f"mean {100 * fr.mean():.1f} %, min {100 * fr.min():.1f} %, max {100 * fr.max():.1f} %")
byte_dep = np.add.reduceat(dep, np.arange(0, len(dep), BS))[1:-1] / BS
print(f"bytes of a block that wait for the previous block: median {100 * np.median(byte_dep):.1f} %, mean {100 * byte_dep.mean():.1f} %")
