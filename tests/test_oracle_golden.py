"""The CPU oracle (oracle/refcpu.c) against every known-answer vector of the reference's own tests.

Each test names the reference test it transcribes (paths relative to /root/reference/zstd-decompressor).
This is what pins the oracle; it runs without a GPU.
"""
import hashlib

import pytest

import refcpu as R
from conftest import fixture_bytes


def err_of(fn, *a, **k):
    with pytest.raises(R.RefError) as ei:
        fn(*a, **k)
    return ei.value


# ---------------------------------------------------------------- tests/parsing.rs : forward bits
def test_fwd_new_empty():                       # new_empty_data_nok
    assert err_of(R.FwdBits, b"").code == R.EmptyInputData

def test_fwd_len_and_empty():                   # is_empty_ok / is_empty_nok / len_ok / len_ok2
    p = R.FwdBits([1]); assert not p.is_empty() and p.len() == 8
    p.take(8); assert p.is_empty() and p.len() == 0

def test_fwd_take_whole_byte():                 # take_whole_byte_ok
    p = R.FwdBits([75]); assert p.take(8) == 75 and p.is_empty()

def test_fwd_take_byte_and_half():              # take_whole_byte_and_half_ok
    p = R.FwdBits([75, 0b0000_1111]); assert p.take(12) == (15 << 8) + 75 and p.len() == 4

def test_fwd_take_few():                        # take_few_ok
    p = R.FwdBits([0b0101_1010, 0b1100_0011])
    assert [p.take(3), p.take(3), p.take(4), p.take(6)] == [0b010, 0b011, 0b1101, 0b110000] and p.is_empty()

def test_fwd_take_more_than_64():               # take_more_than_64_nok
    p = R.FwdBits([1] * 10); e = err_of(p.take, 67)
    assert (e.code, e.a) == (R.MaximumReadableBitsExceeded, 67) and p.len() == 80

def test_fwd_take_more_than_available():        # take_more_than_available_nok
    p = R.FwdBits([1] * 6); e = err_of(p.take, 60)
    assert (e.code, e.a, e.b) == (R.NotEnoughBits, 60, 48) and p.len() == 48

def test_fwd_bytes_read():                      # parsing.rs:122-126
    p = R.FwdBits([0xff, 0xff, 0xff]); assert p.bytes_read() == 0
    p.take(1); assert p.bytes_read() == 1
    p.take(7); assert p.bytes_read() == 1
    p.take(1); assert p.bytes_read() == 2


# ---------------------------------------------------------------- tests/parsing.rs : backward bits
def test_bwd_null_byte():                       # new_null_byte_error_nok
    assert err_of(R.BwdBits, [0]).code == R.NullByte

def test_bwd_empty():                           # new_empty_data_error_nok
    assert err_of(R.BwdBits, b"").code == R.EmptyInputData

def test_bwd_is_empty_and_len():                # is_empty_ok / is_empty_nok / len_ok / len_ok2 / len_ok3
    assert R.BwdBits([1]).is_empty() and R.BwdBits([1]).len() == 0
    assert not R.BwdBits([2]).is_empty()
    assert R.BwdBits([0x5f, 1]).len() == 8
    assert R.BwdBits([0x5f, 0xff]).len() == 15

def test_bwd_take_whole_byte():                 # take_whole_byte_ok
    p = R.BwdBits([0x5f, 1]); assert p.take(8) == 0b0101_1111 and p.is_empty()

def test_bwd_take_byte_and_half():              # take_whole_byte_and_half_ok
    p = R.BwdBits([0b0000_1111, 0b0111_0101, 1]); assert p.take(12) == 0b0111_0101_0000 and p.len() == 4

def test_bwd_take_few():                        # take_few_ok
    p = R.BwdBits([0b0101_1010, 0b1100_0011, 1])
    assert [p.take(3), p.take(3), p.take(4), p.take(6)] == [0b110, 0, 0b1101, 0b011010] and p.is_empty()

def test_bwd_take_more_than_64():               # take_more_than_64_nok
    p = R.BwdBits([1] * 10); e = err_of(p.take, 67)
    assert (e.code, e.a) == (R.MaximumReadableBitsExceeded, 67) and p.len() == 72

def test_bwd_take_more_than_available():        # take_more_than_available_nok
    p = R.BwdBits([1] * 6); e = err_of(p.take, 60)
    assert (e.code, e.a, e.b) == (R.NotEnoughBits, 60, 40) and p.len() == 40

def test_bwd_take_zero_never_fails():           # weird_bug_ok_should_not_panic
    for i in range(15):
        p = R.BwdBits([0b10100000, 0b01111000])
        try:
            p.take(i)
        except R.RefError:
            pass
        assert p.take(0) == 0


# ---------------------------------------------------------------- tests/parsing.rs : byte parser (via frame/block parsing)
def test_byte_parser_errors_through_frames():   # u8 / slice / le_u32 error payloads
    out, frames, err = R.decode_frames(bytes([0x00, 0x00, 0x00]))
    assert (err.code, err.a, err.b) == (R.NotEnoughBytes, 4, 3)


# ---------------------------------------------------------------- tests/decoders/fse.rs
def test_parse_fse_table():                     # parse_fse_table_test_ok
    al, dist, bits_left, _ = R.parse_fse_table([0x30, 0x6f, 0x9b, 0x03])
    assert al == 5 and dist == [18, 6, 2, 2, 2, 1, 1] and bits_left == 6

def test_fse_table_from_distribution():         # fse_table_from_distribution_ok
    t = R.fse_from_distribution(5, [18, 6, 2, 2, 2, 1, 1])
    assert t[0xc] == (1, 0x18, 3)

def test_fse_table_from_description():          # fse_table_from_distribution2_ok
    data = [0x21, 0x9d, 0x51, 0xcc, 0x18, 0x42, 0x44, 0x81, 0x8c, 0x94, 0xb4, 0x50, 0x1e]
    al, dist, _, _ = R.parse_fse_table(data)
    t = R.fse_from_distribution(al, dist)
    assert al == 6 and len(dist) == 25
    assert t[0x3f] == (24, 0x10, 4) and t[0x2c] == (0, 0x34, 2)

HAND_TABLE = [(0, 1, 0), (3, 2, 1), (1, 0, 1), (0, 2, 1)]   # (output, baseline, bits_to_read), al = 2

def test_fse_decoder_full_run():                # run_full_decoder_ok
    syms, _, err = R.fse_run(HAND_TABLE, 2, [0b10100000, 0b11110000], 11)
    assert err is None and syms == [0, 0, 1, 0, 3, 1, 0, 3, 0, 1, 3]

def test_predefined_tables_spot_checks():       # SURVEY section 9 (sequences.rs:29-39 through from_distribution)
    LL = [4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1]
    OF = [1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1]
    ML = [1, 4, 3, 2, 2, 2, 2, 2, 2] + [1] * 37 + [-1] * 7
    ll, of, ml = R.fse_from_distribution(6, LL), R.fse_from_distribution(5, OF), R.fse_from_distribution(6, ML)
    assert ll[:4] == [(0, 0, 4), (0, 16, 4), (1, 32, 5), (3, 0, 5)] and ll[-2:] == [(33, 0, 6), (32, 0, 6)]
    assert of[:4] == [(0, 0, 5), (6, 0, 4), (9, 0, 5), (15, 0, 5)]
    assert ml[:4] == [(0, 0, 6), (1, 0, 4), (2, 32, 5), (3, 0, 5)]

def test_fse_large_accuracy_log():              # fse.rs:18-20
    e = err_of(R.parse_fse_table, [0x05, 0, 0, 0])      # al = 5 + 5 = 10
    assert (e.code, e.a) == (R.LargeAccuracyLog, 10)


# ---------------------------------------------------------------- tests/decoders/alternating.rs
def test_alternating_run():                     # alternating_initialize_test
    syms, _, err = R.fse_run(HAND_TABLE, 2, [0b1001_1_000, 0b0000_0001, 0b1111_1110], 22, alternating=True)
    assert err is None
    assert syms == [0, 0, 0, 0, 1, 1, 0, 0, 3, 3, 1, 1, 0, 0, 3, 3, 0, 0, 1, 1, 3, 3]


# ---------------------------------------------------------------- tests/decoders/huffman.rs
W_ABC = [0] * 65 + [1, 2]

def test_huffman_from_weights_codes():          # example_tree / insert_example_tree_ok
    t = R.huffman_from_weights(W_ABC)
    assert t == {ord("A"): (2, 0b00), ord("C"): (2, 0b01), ord("B"): (1, 0b1)}

def test_huffman_project_example():             # huffman_project_example (also tests/parsing.rs)
    t = R.huffman_from_weights(W_ABC)
    assert R.huffman_decode_stream(t, [0x97, 0x01]) == b"BABCBB"

def test_huffman_parse_direct():                # parse_direct_stream_ok
    packed = []
    for i in range(0, len(W_ABC), 2):
        c = W_ABC[i:i + 2]
        packed.append((c[0] << 4) + (c[1] & 0xf if len(c) > 1 else 0))
    data = bytes([127 + 67] + packed)
    t, consumed, weights = R.huffman_parse(data)
    assert consumed == len(data) and list(weights) == W_ABC
    assert R.huffman_decode_stream(t, [0x97, 0x01]) == b"BABCBB"


# ---------------------------------------------------------------- decoding_context.rs:109-122
def test_execute_sequences():
    out = R.execute_sequences(0x42, [(3, 5, 3), (2, 11, 1)], b"abcdefgh")
    assert out == bytes([0x61, 0x62, 0x63, 0x62, 0x63, 0x62, 0x64, 0x65, 0x61, 0x66, 0x67, 0x68])

def test_context_window_too_big():              # decoding_context.rs:30-35
    e = err_of(R.execute_sequences, (8 << 20) + 1, [], b"")
    assert (e.code, e.a, e.b) == (R.WindowSizeTooBig, 8 << 20, (8 << 20) + 1)


# ---------------------------------------------------------------- tests/block.rs  (wrapped in a minimal frame)
def frame_wrap(blocks, checksum=None, fhd=0x20, fcs=b"\x00"):
    """magic + single-segment header (FCS 1 byte, ignored by the reference) + blocks."""
    return b"\x28\xb5\x2f\xfd" + bytes([fhd | (4 if checksum is not None else 0)]) + fcs + blocks + (checksum or b"")

def test_block_raw_last():                      # decode_raw_block_last
    out, frames, err = R.decode_frames(frame_wrap(bytes([0x21, 0, 0, 0x10, 0x20, 0x30, 0x40])))
    assert err is None and out == bytes([0x10, 0x20, 0x30, 0x40]) and frames[0]["n_blocks"] == 1

def test_block_rle_not_last():                  # decode_rle_block_not_last: repeat 196612 > 128 KiB accepted
    out, frames, err = R.decode_frames(frame_wrap(bytes([0x22, 0x0, 0x18, 0x42]) + bytes([0x09, 0, 0, 0x50])))
    assert err is None and len(out) == 196612 + 1 and set(out[:196612]) == {0x42} and frames[0]["n_blocks"] == 2

def test_block_reserved_type():                 # reserved_block_error_test
    _, _, err = R.decode_frames(frame_wrap(bytes([0x27, 0, 0, 0x10, 0x20, 0x30, 0x40, 0x50])))
    assert err.code == R.ReservedBlockType

def test_block_not_enough_bytes():              # not_enough_bytes_error_test
    _, _, err = R.decode_frames(frame_wrap(bytes([0x21, 0, 0, 0x10, 0x20, 0x30])))
    assert (err.code, err.a, err.b) == (R.NotEnoughBytes, 4, 3)


# ---------------------------------------------------------------- tests/frame.rs
SKIP = bytes([0x53, 0x2a, 0x4d, 0x18, 0x03, 0, 0, 0, 0x10, 0x20, 0x30])
ZFRAME = bytes([0x28, 0xB5, 0x2F, 0xFD, 0b01_1_0_0_1_00, 0x04, 0x00, 0x21, 0, 0, 0x10, 0x20, 0x30, 0x40, 0x01, 0, 0, 0])

def test_frame_skippable():                     # parse_skippable_frame_ok / decode_skippable_frame_test
    out, frames, err = R.decode_frames(SKIP)
    assert err is None and out == bytes([0x10, 0x20, 0x30])
    assert frames[0]["kind"] == 1 and frames[0]["magic"] == 0x184d2a53 and frames[0]["src_len"] == 11

def test_frame_standard():                      # parse_standard_frame_ok / parse_with_checksum_ok
    out, frames, err = R.decode_frames(ZFRAME)
    f = frames[0]
    assert err is None and out == bytes([0x10, 0x20, 0x30, 0x40])
    assert f["has_checksum"] and f["stored_checksum"] == 1 and f["n_blocks"] == 1 and f["src_len"] == len(ZFRAME)
    # the reference never validates; a wrong stored checksum still decodes
    assert f["computed_xxh64_low32"] != 1

def test_frame_without_checksum():              # parse_without_checksum_ok
    d = bytes([0x28, 0xB5, 0x2F, 0xFD, 0b01_1_0_0_0_00, 0x04, 0x00, 0x21, 0, 0, 0x10, 0x20, 0x30, 0x40])
    out, frames, err = R.decode_frames(d)
    assert err is None and not frames[0]["has_checksum"] and frames[0]["src_len"] == len(d)

def test_frame_unknown_magic():                 # parsing_error_on_unknown_frame
    _, _, err = R.decode_frames(bytes([0x10, 0x20, 0x30, 0x40]))
    assert (err.code, err.a) == (R.UnrecognizedMagic, 0x40302010)

def test_skippable_truncated_data():            # parsing_error_on_truncated_data_frame
    _, _, err = R.decode_frames(bytes([0x53, 0x2a, 0x4d, 0x18, 0x03, 0, 0, 0, 0x10, 0x20]))
    assert (err.code, err.a, err.b) == (R.NotEnoughBytes, 3, 2)

def test_skippable_truncated_length():          # parsing_error_on_truncated_length_frame
    _, _, err = R.decode_frames(bytes([0x53, 0x2a, 0x4d, 0x18, 0x03, 0, 0]))
    assert (err.code, err.a, err.b) == (R.NotEnoughBytes, 4, 3)

def test_skippable_truncated_magic():           # parsing_error_on_truncated_magic_frame
    _, _, err = R.decode_frames(bytes([0x53, 0x2a, 0x4d]))
    assert (err.code, err.a, err.b) == (R.NotEnoughBytes, 4, 3)

def test_header_fcs2():                         # simple_valid_data_ok
    h, n = R.header_parse(bytes([0b01_1_0_0_0_00, 0xcc, 0xcc]))
    assert not h.content_checksum_flag and h.window_size == 0xcccc + 256 and h.content_size == 0xcccc + 256 and not h.has_dict_id

def test_header_window_and_fcs2():              # simple_valid_data_2_ok
    h, n = R.header_parse(bytes([0b01_0_0_0_0_00, 0x00, 0xcc, 0xdd]))
    assert h.window_size == 1024 and h.content_size == 0xddcc + 256 and not h.has_dict_id

def test_header_reserved_bit():                 # reserved_bit_set_should_throw_error
    assert err_of(R.header_parse, bytes([0b01_0_0_1_0_00])).code == R.FrameReservedSet

def test_header_dict_id():                      # simple_valid_data_with_dict_id_ok
    h, n = R.header_parse(bytes([0b01_0_0_0_0_10, 0x00, 0xef, 0xab, 0xcc, 0xdd]))
    assert h.window_size == 1024 and h.content_size == 0xddcc + 256 and h.has_dict_id and h.dictionnary_id == 0xabef

def test_header_long_dict_and_fcs():            # simple_valid_data_long_dict_and_fc_ok
    h, n = R.header_parse(bytes([0b11_0_0_0_0_11, 0x00, 0xef, 0xab, 0xef, 0xab] + [0xcc, 0xdd] * 4))
    assert h.window_size == 1024 and h.content_size == 0xddccddccddccddcc and h.dictionnary_id == 0xabefabef

def test_frame_missing_checksum():              # parse_no_checksum_error
    d = bytes([0x28, 0xB5, 0x2F, 0xFD, 0b01_1_0_0_1_00, 0x04, 0x00, 0x21, 0, 0, 0x10, 0x20, 0x30, 0x40, 0x42])
    _, _, err = R.decode_frames(d)
    assert (err.code, err.a, err.b) == (R.MissingChecksum, 4, 1)

def test_frame_window_too_big():                # parse_window_size_too_big_error
    d = bytes([0x28, 0xB5, 0x2F, 0xFD, 0b01_0_0_0_1_00, 0xff, 0x04, 0x05, 0x21, 0, 0, 0x10, 0x20, 0x30, 0x40, 0x42])
    _, _, err = R.decode_frames(d)
    assert (err.code, err.a, err.b) == (R.WindowSizeTooBig, 8 << 20, (1 << 41) + 7 * (1 << 38))

def test_window_descriptor():                   # frame.rs:281-309
    assert R.lib().rc_window_descriptor(0) == 1 << 10
    assert R.lib().rc_window_descriptor(0xff) == (1 << 41) + 7 * (1 << 38)
    assert R.lib().rc_window_descriptor(1) == (1 << 10) + 1024 // 8

def test_frame_iterator_two_frames():           # frame_iterator_tests::next_test
    d = bytes([0x53, 0x2a, 0x4d, 0x18, 0x03, 0, 0, 0, 0x10, 0x20, 0x30, 0x51, 0x2a, 0x4d, 0x18, 0x04, 0, 0, 0, 0x10, 0x20, 0x30, 0x40])
    out, frames, err = R.decode_frames(d)
    assert err is None and [out[f["out_off"]:f["out_off"] + f["out_len"]] for f in frames] == [bytes([0x10, 0x20, 0x30]), bytes([0x10, 0x20, 0x30, 0x40])]


# ---------------------------------------------------------------- tests/decoders/sequence.rs
def test_fuzzer_input_is_an_error_not_a_crash():   # fuzzer_panic_ok
    data = bytes([40, 181, 47, 253, 0, 10, 165, 0, 0, 85, 47, 0, 252, 59, 64, 44, 0, 51, 29, 44, 47, 10,
                  40, 0, 181, 181, 40, 181, 47, 253])
    _, _, err = R.decode_frames(data)
    assert err is not None and err.code == R.SequenceCodeMaxValueExceeded


# ---------------------------------------------------------------- resources/*.zst (SURVEY 8c fixture table)
@pytest.mark.parametrize("name,size,xxh", [
    ("welcome.zst", 126, 0x9f5d2e9e), ("romeo.txt.zst", 942, 0x51951482), ("moby-dick.txt.zst", 1276235, 0x688efa5c)])
def test_fixture_decodes_and_checksum(name, size, xxh):
    out = R.main_decode(fixture_bytes(name))
    assert len(out) == size and (R.xxh64(out) & 0xffffffff) == xxh

def test_fixture_moby_sha256():
    out = R.main_decode(fixture_bytes("moby-dick.txt.zst"))
    assert hashlib.sha256(out).hexdigest().startswith("61d5ab6a3910fab6")

def test_fixture_skippables():
    d = fixture_bytes("skippables.zst")
    assert R.main_decode(d) == b"" and len(R.main_decode(d, print_skippable=True)) == 4

def test_fixture_romeo3():
    one = R.main_decode(fixture_bytes("romeo.txt.zst"))
    assert R.main_decode(fixture_bytes("romeo3.txt.zst")) == one * 3

def test_fixture_welcome_structure():
    _, frames, err = R.decode_frames(fixture_bytes("welcome.zst"))
    assert err is None and [(f["kind"], f["out_len"]) for f in frames] == [(1, 48), (0, 126)]
    assert frames[0]["magic"] == 0x184D2A57 and frames[1]["n_blocks"] == 4

def test_mt_driver_matches_serial():
    d = fixture_bytes("romeo3.txt.zst") + fixture_bytes("welcome.zst")
    assert R.main_decode(d, threads=4) == R.main_decode(d)


# ---------------------------------------------------------------- reference quirks
def test_quirk_q2_empty_raw_block_rejected():   # frame libzstd emits for empty input
    d = bytes.fromhex("28b52ffd2400010000" + "99e9d851")
    _, _, err = R.decode_frames(d, quirks=True)
    assert err.code == R.EmptySliceError
    out, _, err = R.decode_frames(d, quirks=False)
    assert err is None and out == b""

def test_quirk_q2_empty_skippable_rejected():
    _, _, err = R.decode_frames(bytes([0x50, 0x2a, 0x4d, 0x18, 0, 0, 0, 0]))
    assert err.code == R.EmptySliceError

def test_xxh64_known_answers():
    assert R.xxh64(b"") == 0xEF46DB3751D8E999
    assert R.xxh64(b"a") == 0xD24EC4F1A98C6E5B
    assert R.xxh64(b"abc") == 0x44BC2CF5AD770999
