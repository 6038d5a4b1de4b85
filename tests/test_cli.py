"""cli/zstd-decompressor: the flags and behaviour of the reference's src/main.rs:7-60.
`--info` runs on the host only (CPU tests); decoding needs a GPU."""
import os
import subprocess

import pytest

import corpora
import refcpu as R
import zstd_decompressor_b200 as Z

FIX = os.path.join(corpora.ROOT, "tests", "fixtures")


@pytest.fixture(scope="module")
def cli():
    return Z.build_cli()


def run(cli, *args):
    return subprocess.run([cli, *args], capture_output=True)


def rust_hex_list(data, depth):
    """`{:#x?}` of a byte slice at an indentation depth"""
    pad = "    " * depth
    return "[\n" + "".join(f"{pad}    {b:#x},\n" for b in data) + pad + "]"


def test_info_skippable_frames_like_rust_debug(cli):
    r = run(cli, "--info", os.path.join(FIX, "skippables.zst"))
    assert r.returncode == 0
    want = ""
    for magic, payload in ((0x184d2a53, b"\x10\x20\x30"), (0x184d2a51, b"\x42")):
        want += f"SkippableFrame(\n    Skippable {{\n        magic: {magic:#x},\n        data: {rust_hex_list(payload, 2)},\n    }},\n)\n"
    assert r.stdout.decode() == want


def test_info_zstandard_frame_header_blocks_checksum(cli):
    r = run(cli, "-i", os.path.join(FIX, "welcome.zst"))
    out = r.stdout.decode()
    assert r.returncode == 0
    # second frame of welcome.zst: RLE, Raw, RLE, Raw blocks, single segment (window = content size = 126), checksum 0x9f5d2e9e
    z = out[out.index("ZStandardFrame("):]
    assert z.startswith("ZStandardFrame(\n    ZStandard {\n        header: Header {\n            content_checksum_flag: true,\n"
                        "            window_size: 0x7e,\n            dictionnary_id: None,\n            content_size: Some(\n                0x7e,\n            ),\n        },\n"
                        "        blocks: [\n            RLEBlock {\n                byte: ")
    assert z.count("RLEBlock {") == 2 and z.count("RawBlock(") == 2
    assert z.endswith("        ],\n        checksum: Some(\n            0x9f5d2e9e,\n        ),\n    },\n)\n")


def test_info_compressed_blocks_huffman_tree_and_fse_tables(cli, tmp_path):
    """main.rs:35-40 prints the PARSED frame: the Huffman tree through the reference's own Debug impl (huffman.rs:60-77) and the FSE
    tables state by state (fse.rs:72-89).  Expected text: tests/rust_debug.py (container parsed in Python, tree and tables from the CPU
    oracle), byte for byte; romeo.txt.zst also against the committed dump tests/golden/romeo_info.txt (written by the same module)."""
    import rust_debug as D
    import zasm
    import gen_corpus as G
    r = run(cli, "--info", os.path.join(FIX, "romeo.txt.zst"))
    assert r.returncode == 0
    assert r.stdout.decode() == open(os.path.join(corpora.ROOT, "tests", "golden", "romeo_info.txt")).read() == D.dump(corpora.fixture("romeo.txt.zst"))
    cases = [zasm.frame_rle_modes(1)[0], zasm.frame_huffman_direct(2)[0], zasm.frame_huffman_direct(3, n=700, streams=1)[0], zasm.frame_treeless(3)[0],
             G.compress(G.moby_text()[:200]), G.compress(G.moby_text()[200000:230000], level=19), corpora.fixture("welcome.zst") + corpora.fixture("romeo3.txt.zst")]
    for i, d in enumerate(cases):
        p = tmp_path / f"c{i}.zst"; p.write_bytes(d)
        r = run(cli, "-i", str(p))
        assert r.returncode == 0 and r.stdout.decode() == D.dump(d), i


def test_info_section_error_prints_nothing_of_that_frame(cli, tmp_path):
    """frame.rs:210-223: the sections are parsed with the frame; a bad modes byte (sequences::Error::ReservedSet) fails the second frame"""
    d = bytearray(corpora.fixture("romeo.txt.zst"))
    import zstd_inspect as I
    blk = I.inspect(bytes(d))[0].blocks[0]
    # the modes byte follows the literals section (header 5 bytes for size format 3? use the inspector's numbers) and the 1-byte sequence count
    hdr = 3 if ((d[blk.src_off] >> 2) & 3) <= 1 else 4 if ((d[blk.src_off] >> 2) & 3) == 2 else 5
    pos = blk.src_off + hdr + blk.lit_csize + 1
    d[pos] |= 1
    p = tmp_path / "two.zst"; p.write_bytes(corpora.fixture("welcome.zst") + bytes(d))
    r = run(cli, "--info", str(p))
    want, _, oerr = R.decode_frames(p.read_bytes(), quirks=True)
    assert oerr is not None and oerr.code == 30
    assert r.returncode == 1 and b"ReservedSet" in r.stderr
    import rust_debug as D
    assert r.stdout.decode() == D.dump(corpora.fixture("welcome.zst"))


def test_info_reports_bad_magic_after_the_good_frames(cli, tmp_path):
    p = tmp_path / "bad.zst"
    p.write_bytes(corpora.fixture("skippables.zst") + b"\x01\x02\x03\x04\x05")
    r = run(cli, "--info", str(p))
    assert r.returncode == 1 and r.stdout.decode().count("SkippableFrame(") == 2 and b"UnrecognizedMagic" in r.stderr


def test_usage_errors(cli):
    assert run(cli).returncode == 2
    assert run(cli, "--nope", "x").returncode == 2
    assert run(cli, "/nonexistent/file.zst").returncode == 1


@pytest.mark.gpu
def test_decode_to_stdout_and_file(cli, tmp_path):
    for name in ("welcome.zst", "romeo3.txt.zst", "moby-dick.txt.zst"):
        path = os.path.join(FIX, name)
        want = R.main_decode(corpora.fixture(name))
        r = run(cli, path)
        assert r.returncode == 0 and r.stdout == want
        o = tmp_path / "out.txt"
        o.write_bytes(b"old contents that must be overwritten" * 1000)
        r = run(cli, path, "-o", str(o))
        assert r.returncode == 0 and r.stdout == b"" and o.read_bytes() == want
    r = run(cli, "-p", os.path.join(FIX, "welcome.zst"))
    assert r.returncode == 0 and r.stdout == R.main_decode(corpora.fixture("welcome.zst"), print_skippable=True)


@pytest.mark.gpu
def test_error_means_no_output_and_non_utf8_is_refused(cli, tmp_path):
    bad = bytearray(corpora.fixture("romeo.txt.zst")); bad[300] ^= 0x55
    p = tmp_path / "bad.zst"; p.write_bytes(corpora.fixture("romeo.txt.zst") + bytes(bad))
    want, _, oerr = R.decode_frames(p.read_bytes(), quirks=True)
    r = run(cli, str(p))
    if oerr is not None:
        assert r.returncode == 1 and r.stdout == b""                    # main.rs:51: the first error aborts, nothing is printed
    import gen_corpus as G
    q = tmp_path / "bin.zst"; q.write_bytes(G.compress(bytes(range(256)) * 8))
    r = run(cli, str(q))
    assert r.returncode == 101 and r.stdout == b""                      # main.rs:57 from_utf8().unwrap()
    r = run(cli, "--binary", str(q))
    assert r.returncode == 0 and r.stdout == bytes(range(256)) * 8
