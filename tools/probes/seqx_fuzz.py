"""ZSB_SEQX=1 on mutated inputs, each in its own process (a CUDA fault is sticky): prints the inputs that fault or differ from the default path
and saves them under gpurun_out/."""
import os, random, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import zstd_decompressor_b200 as Z
    d = open(sys.argv[2], "rb").read()
    fl = Z.REFERENCE_QUIRKS | Z.VERIFY_CHECKSUM
    out0, sc0, r0 = Z.Decoder(Z.Context(0)).decode(d, fl)
    os.environ["ZSB_SEQX"] = "1"
    dx = Z.Decoder(Z.Context(0))
    out1, sc1, r1 = dx.decode(d, fl)
    same = out0 == out1 and list(r0.status[:sc0.n_frames]) == list(r1.status[:sc1.n_frames])
    print("same" if same else "DIFFERENT", "seqx state", dx.ctx.last_seqx_state(), "status", list(r1.status[:sc1.n_frames]))
    sys.exit(0 if same else 4)
import corpora
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
bad = 0
srcs = list(corpora.mutation_sources().items())
if os.environ.get("TOUR"):          # the inputs of tools/probes/sanitize_target.py, in its order
    order = [(name, src, k) for name, src in srcs[:3] for k in range(6)]
else:
    order = [(name, src, k) for k in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6) for name, src in srcs]
for name, src, rnd_i in order:
    if True:
        d = corpora.mutate(rnd, src)
        p = os.path.join(ROOT, "gpurun_out", f"fuzz_{name}_{rnd_i}.zst")
        open(p, "wb").write(d)
        r = subprocess.run([sys.executable, __file__, "one", p], capture_output=True, text=True)
        if r.returncode != 0:
            bad += 1
            print(name, rnd_i, "rc", r.returncode, (r.stdout + r.stderr).strip().split("\n")[-1][:300])
        else:
            os.remove(p)
print("bad:", bad)
