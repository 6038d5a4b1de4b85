"""tests/golden/corpora.json (written by tests/golden/make_golden.py): the deterministic corpora reproduce on this machine, the
oracle decodes them to the recorded bytes, and -- on a GPU -- so does the CUDA path."""
import hashlib
import json
import os

import pytest

import corpora
import refcpu as R

GOLDEN = json.load(open(os.path.join(corpora.ROOT, "tests", "golden", "corpora.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def blobs():
    for name in corpora.FIXTURE_NAMES:
        yield f"fixture:{name}", corpora.fixture(name)
    yield "c4(seed=4)", corpora.c4()[0]
    for n in (16, 64):
        yield f"c2(frames={n},seed=2)", corpora.c2_small(n)[0]
    yield "c3(total=3MiB)", corpora.c3_small(3 << 20)[0]


def test_moby_dick_hash_is_the_surveyed_one():
    assert GOLDEN["fixture:moby-dick.txt.zst"]["decoded_sha256"].startswith("61d5ab6a3910fab6")


def test_oracle_reproduces_the_golden_hashes():
    for name, blob in blobs():
        g = GOLDEN[name]
        if sha(blob) != g["compressed_sha256"]:
            assert not name.startswith("fixture:"), name
            pytest.skip(f"{name}: libzstd here emits different bytes than libzstd 1.5.5 (the decoded bytes are checked by the other tests)")
        dec = R.main_decode(blob)
        assert (len(dec), sha(dec)) == (g["decoded_bytes"], g["decoded_sha256"]), name
        if "decoded_with_skippable_sha256" in g:
            assert sha(R.main_decode(blob, print_skippable=True)) == g["decoded_with_skippable_sha256"], name


@pytest.mark.gpu
def test_gpu_reproduces_the_golden_hashes():
    import zstd_decompressor_b200 as Z
    dec = Z.Decoder(Z.Context(0))
    for name, blob in blobs():
        g = GOLDEN[name]
        if sha(blob) != g["compressed_sha256"]:
            continue
        out, sc, r = dec.decode(blob, Z.REFERENCE_QUIRKS | Z.VERIFY_CHECKSUM)
        assert sc.status == 0 and r.first_error() is None and sha(out) == g["decoded_sha256"], name
        if "decoded_with_skippable_sha256" in g:
            out, sc, r = dec.decode(blob, Z.REFERENCE_QUIRKS | Z.VERIFY_CHECKSUM | Z.PRINT_SKIPPABLE)
            assert sha(out) == g["decoded_with_skippable_sha256"], name
