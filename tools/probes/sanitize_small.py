"""Small decode workload for compute-sanitizer (memcheck / racecheck / synccheck): every fixture, the mixed corpus C4,
a few C2 frames and mutated inputs (development probe)."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpora
import zstd_decompressor_b200 as Z

dec = Z.Decoder(Z.Context(0))
n = 0
for name in corpora.FIXTURE_NAMES:
    out, sc, r = dec.decode(corpora.fixture(name), 6 | 1); n += 1
blob, exp, exp_skip, _ = corpora.c4()
out, sc, r = dec.decode(blob, 6); assert out == exp; n += 1
blob, exp = corpora.c2_small(40)
out, sc, r = dec.decode(blob, 6); assert out == exp; n += 1
rr = random.Random(3)
for name, d in list(corpora.mutation_sources().items())[:6]:
    for _ in range(6):
        dec.decode(corpora.mutate(rr, d), 6 | 1); n += 1
print("decoded", n, "inputs")
