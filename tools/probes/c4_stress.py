"""C4 decoded over and over with both flag sets (ITERS, default 20): per-frame checksum verdicts and the output must never change (found a divergent polling loop in the trailing hash)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import corpora, zstd_decompressor_b200 as Z
dec = Z.Decoder(Z.Context(0))
blob, exp, exp_skip, parts = corpora.c4()
for it in range(int(os.environ.get("ITERS", "20"))):
    for fl in (Z.VERIFY_CHECKSUM | Z.PRINT_SKIPPABLE, Z.VERIFY_CHECKSUM | Z.REFERENCE_QUIRKS):
        out, sc, r = dec.decode(blob, fl)
        bad = [i for i in range(sc.n_frames) if sc.frames[i].kind == 0 and sc.frames[i].has_checksum and not r.checksum_ok[i]]
        if bad or out != (exp_skip if fl & Z.PRINT_SKIPPABLE else exp): print(it, fl, "bad", bad, [(sc.frames[i].n_blocks, r.dst_len[i], hex(r.xxh32[i]), hex(sc.frames[i].stored_checksum)) for i in bad], "out ok", out == (exp_skip if fl & Z.PRINT_SKIPPABLE else exp))
print("done")
