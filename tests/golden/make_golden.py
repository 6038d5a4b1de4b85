#!/usr/bin/env python
"""Writes tests/golden/corpora.json: for every deterministic corpus of the parity tests the SHA-256 of the compressed bytes (as
produced by libzstd 1.5.5 / the hand assembler with the fixed seeds of tools/gen_corpus.py) and of the decoded bytes (as produced
by the CPU oracle, oracle/refcpu.c, which restates the reference; the Rust reference itself cannot be run in this image).

    python tests/golden/make_golden.py          # regenerate (run in the build container)

The five reference fixtures are data copied from the reference (`resources/*.zst`); their decoded hashes are what both the oracle
and libzstd produce and, for moby-dick, what SURVEY.md 8(c) records (sha256 61d5ab6a3910fab6...)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import corpora  # noqa: E402
import refcpu as R  # noqa: E402


def sha(b):
    return hashlib.sha256(b).hexdigest()


def entries():
    for name in corpora.FIXTURE_NAMES:
        d = corpora.fixture(name)
        yield f"fixture:{name}", d, R.main_decode(d), R.main_decode(d, print_skippable=True)
    blob, exp, exp_skip, _ = corpora.c4()
    yield "c4(seed=4)", blob, R.main_decode(blob), R.main_decode(blob, print_skippable=True)
    for n in (16, 64):
        blob, exp = corpora.c2_small(n)
        yield f"c2(frames={n},seed=2)", blob, R.main_decode(blob), None
    blob, exp = corpora.c3_small(3 << 20)
    yield "c3(total=3MiB)", blob, R.main_decode(blob), None


def main():
    out = {}
    for name, blob, dec, dec_skip in entries():
        out[name] = {"compressed_bytes": len(blob), "compressed_sha256": sha(blob), "decoded_bytes": len(dec), "decoded_sha256": sha(dec)}
        if dec_skip is not None and dec_skip != dec:
            out[name]["decoded_with_skippable_sha256"] = sha(dec_skip)
    with open(os.path.join(HERE, "corpora.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
