"""C-ABI library: loads without a GPU, exports every symbol include/zsb.h declares, refuses to decode
without a device (no CPU fallback), and its host-side frame/block walk matches the reference."""
import ctypes as C
import os
import random
import re

import pytest

import corpora
import refcpu as R
import zstd_decompressor_b200 as Z
import zstd_inspect as I

ROOT = corpora.ROOT


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "zsb.h")).read()
    declared = sorted(set(re.findall(r"\b(zsb_[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(Z.EXPORTED_SYMBOLS)
    L = Z.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(Z.ZsbError) as ei:
        Z.Context(0)
    assert ei.value.code == Z.E_CUDA
    with pytest.raises(Z.ZsbError):
        Z.decompress(corpora.fixture("welcome.zst"))


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zstd-decompressor_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("librefcpu", "refcpu.h", "import refcpu", "emul_lib", "libzsb_emul", "emul.cpp", "oracle/_", "/oracle\""):
                    assert needle not in src, (f, needle)


def test_scan_matches_inspector_on_corpora():
    for blob in (corpora.c4()[0], corpora.fixture("moby-dick.txt.zst"), corpora.fixture("skippables.zst")):
        sc = Z.Scan(blob, Z.REFERENCE_QUIRKS)
        want = I.inspect(blob)
        assert sc.status == 0 and sc.n_frames == len(want)
        for i, w in enumerate(want):
            f = sc.frames[i]
            assert (f.kind == 1) == (w.kind == "skippable") and f.src_off == w.src_off and f.src_len == w.src_len and f.magic == w.magic
            if w.kind == "zstd":
                assert f.n_blocks == len(w.blocks) and f.window_size == w.window_size and bool(f.has_checksum) == w.has_checksum
                assert (f.content_size if f.has_content_size else None) == w.content_size
                assert (f.dict_id if f.has_dict_id else None) == w.dict_id
                if w.has_checksum:
                    assert f.stored_checksum == w.checksum
                for k, wb in enumerate(w.blocks):
                    b = sc.blocks[f.first_block + k]
                    assert (b.type, bool(b.last), b.size, b.src_off) == (I.BLOCK_TYPES.index(wb.type), wb.last, wb.size, wb.src_off)


# tests/frame.rs and tests/block.rs error vectors through the host walk
@pytest.mark.parametrize("data,code,a,b", [
    (bytes([0x10, 0x20, 0x30, 0x40]), 60, 0x40302010, 0),                                   # parsing_error_on_unknown_frame
    (bytes([0x53, 0x2a, 0x4d, 0x18, 0x03, 0, 0, 0, 0x10, 0x20]), 1, 3, 2),                    # truncated skippable data
    (bytes([0x53, 0x2a, 0x4d, 0x18, 0x03, 0, 0]), 1, 4, 3),                                   # truncated length
    (bytes([0x53, 0x2a, 0x4d]), 1, 4, 3),                                                     # truncated magic
    (bytes([0x28, 0xB5, 0x2F, 0xFD, 0x24, 0x04, 0x21, 0, 0, 0x10, 0x20, 0x30, 0x40, 0x42]), 62, 4, 1),   # parse_no_checksum_error
    (bytes([0x28, 0xB5, 0x2F, 0xFD, 0x04, 0xff, 0x04, 0x05, 0x21, 0, 0, 1, 2, 3, 4]), 40, 8 << 20, (1 << 41) + 7 * (1 << 38)),
    (bytes([0x28, 0xB5, 0x2F, 0xFD, 0x20, 0x00, 0x27, 0, 0, 1, 2, 3, 4, 5]), 50, 0, 0),       # reserved_block_error_test
    (bytes([0x28, 0xB5, 0x2F, 0xFD, 0x20, 0x00, 0x21, 0, 0, 0x10, 0x20, 0x30]), 1, 4, 3),     # not_enough_bytes_error_test
    (bytes([0x28, 0xB5, 0x2F, 0xFD, 0x08]), 61, 0, 0),                                        # reserved bit
])
def test_scan_error_vectors(data, code, a, b):
    sc = Z.Scan(data, Z.REFERENCE_QUIRKS)
    _, _, oerr = R.decode_frames(data)
    assert (sc.status, sc.err_a, sc.err_b) == (code, a, b)
    assert (oerr.code, oerr.a, oerr.b) == (code, a, b)


def test_scan_structural_errors_match_reference_on_truncations():
    """cut valid buffers at every position of their headers: same error variant and payload as the oracle
    whenever the reference fails in the container walk (not inside a block's sections)."""
    r = random.Random(3)
    blob = corpora.fixture("welcome.zst") + corpora.fixture("skippables.zst")
    for cut in range(1, len(blob)):
        d = blob[:cut]
        sc = Z.Scan(d, Z.REFERENCE_QUIRKS)
        _, _, oerr = R.decode_frames(d)
        assert (sc.status != 0) == (oerr is not None)
        if oerr is not None:
            assert (sc.status, sc.err_a, sc.err_b) == (oerr.code, oerr.a, oerr.b), cut


def test_reference_api_mirror_without_gpu():
    p = Z.ForwardByteParser(corpora.fixture("welcome.zst"))
    frames = list(p.iter())
    assert frames[0].is_skippable and frames[0].magic == 0x184D2A57 and len(frames[0].decode()) == 48
    h = frames[1].header()
    assert h.content_checksum_flag and h.content_size == 126 and h.window_size == 126 and h.dictionnary_id is None
    assert frames[1].checksum() == 0x9f5d2e9e and len(frames[1].blocks()) == 4
    assert Z.MAX_WIN_SIZE == 8 << 20


def test_result_struct_layout_matches_header():
    """zsb_result (include/zsb.h) as the Python mirror sees it: 40 bytes, fields at the offsets ScanDecode.first_error relies on."""
    import ctypes as C
    assert C.sizeof(Z.ZsbResult) == 40
    assert (Z.ZsbResult.dst_off.offset, Z.ZsbResult.dst_len.offset, Z.ZsbResult.status.offset, Z.ZsbResult.xxh32.offset,
            Z.ZsbResult.err_a.offset, Z.ZsbResult.err_b.offset, Z.ZsbResult.checksum_ok.offset) == (0, 8, 16, 20, 24, 28, 32)
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "zsb.h")).read()
    assert "typedef struct zsb_result" in hdr and "zsb_scan_decode" in hdr and "zsb_host_alloc" in hdr
