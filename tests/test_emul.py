"""CPU-only differential tests: the lane-serial device code (compiled with g++ by tests/emul) against
the CPU oracle.  Same functions the sm_100a kernels run one lane per block / per stream."""
import random

import pytest

import corpora
import emul_lib as E
import refcpu as R

Q, SKIP = 4, 1   # ZSB_REFERENCE_QUIRKS, ZSB_PRINT_SKIPPABLE


def as_states(cells):
    return [((c >> 16) & 0x3F, c >> 22, (c >> 8) & 0x1F) for c in cells]   # (output, baseline, bits_to_read)


# ---------------------------------------------------------------- FSE tables (fse.rs)
def test_fse_reference_vectors():
    rc, t = E.fse_parse([0x30, 0x6f, 0x9b, 0x03])
    assert rc == 0 and t["al"] == 5 and t["dist"] == [18, 6, 2, 2, 2, 1, 1] and as_states(t["cells"])[0xc] == (1, 0x18, 3)
    rc, t = E.fse_parse([0x21, 0x9d, 0x51, 0xcc, 0x18, 0x42, 0x44, 0x81, 0x8c, 0x94, 0xb4, 0x50, 0x1e])
    s = as_states(t["cells"])
    assert rc == 0 and t["al"] == 6 and s[0x3f] == (24, 0x10, 4) and s[0x2c] == (0, 0x34, 2) and t["consumed"] == 13


def random_distribution(r, al, max_sym):
    n = 1 << al
    nsym = r.randrange(2, max_sym + 1)
    dist = [0] * nsym
    remaining = n
    low = r.randrange(0, min(nsym, 8))
    for _ in range(low):
        s = r.randrange(nsym)
        if dist[s] == 0 and remaining > 1:
            dist[s] = -1; remaining -= 1
    while remaining:
        s = r.randrange(nsym)
        if dist[s] == -1:
            continue
        k = r.randrange(1, max(2, remaining // 2 + 1)) if r.random() < 0.3 else 1
        k = min(k, remaining)
        dist[s] += k; remaining -= k
    return dist


@pytest.mark.parametrize("stride", [1, 32])
def test_fse_build_equals_reference_grouping(stride):
    """closed form next/nb/base == the reference's per-symbol grouping (fse.rs:169-189) on random distributions"""
    r = random.Random(7)
    for it in range(300):
        al = r.randrange(5, 10)
        dist = random_distribution(r, al, r.choice([8, 29, 36, 53, 60]))
        want = R.fse_from_distribution(al, dist)
        rc, cells = E.fse_build(3, al, dist, stride)
        assert rc == 0 and as_states(cells) == want, (al, dist)


def test_fse_predefined_tables_and_xbits():
    LL = [4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1]
    rc, cells = E.fse_build(0, 6, LL)
    assert rc == 0 and as_states(cells) == R.fse_from_distribution(6, LL)
    ll_bits = [0] * 16 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
    assert all((c & 0x3F) == ll_bits[(c >> 16) & 0x3F] for c in cells)


def test_fse_errors():
    assert E.fse_parse([0x05, 0, 0, 0])[0] == R.LargeAccuracyLog
    assert E.fse_parse([0x00])[0] == R.NotEnoughBits
    rc, _ = E.fse_build(3, 5, [10, 10])          # 20 of 32 cells
    assert rc == R.CorruptedTable


# ---------------------------------------------------------------- Huffman (huffman.rs)
def test_huffman_direct_weights_vector():
    w = [0] * 65 + [1, 2]
    packed = [(w[i] << 4) + (w[i + 1] if i + 1 < len(w) else 0) for i in range(0, len(w), 2)]
    rc, t = E.huf_parse(bytes([127 + 67] + packed))
    assert rc == 0 and t["codes"] == {65: (2, 0), 67: (2, 1), 66: (1, 1)} and t["maxbits"] == 2


def test_huffman_descriptions_from_real_frames():
    """every Huffman tree description found in the corpora: same codes as the reference's tree"""
    import zstd_inspect as I
    blob = corpora.c4()[0] + corpora.fixture("moby-dick.txt.zst")
    n = 0
    for f in I.inspect(blob):
        for k in f.blocks:
            if k.type == "compressed" and k.lit_type == "compressed":
                hdr = 1 + (2 if k.lit_streams == 1 or (k.lit_regen < 1024 and k.lit_csize < 1024) else 3 if k.lit_regen < 16384 and k.lit_csize < 16384 else 4)
                desc = blob[k.src_off + hdr:k.src_off + hdr + k.lit_csize]
                want, consumed, weights = R.huffman_parse(desc)
                rc, t = E.huf_parse(desc)
                assert rc == 0 and t["codes"] == want and t["consumed"] == consumed and t["weights"] == weights
                n += 1
    assert n > 20


# ---------------------------------------------------------------- whole path
@pytest.mark.parametrize("name", corpora.FIXTURE_NAMES)
@pytest.mark.parametrize("skip", [0, SKIP])
def test_fixtures(name, skip):
    d = corpora.fixture(name)
    rc, out, frames, _ = E.decode(d, Q | skip)
    assert rc == 0 and all(f[0] == 0 for f in frames)
    assert out == R.main_decode(d, print_skippable=bool(skip))


def test_c4_all_modes():
    blob, exp, exp_skip, _ = corpora.c4()
    rc, out, frames, _ = E.decode(blob, Q)
    assert rc == 0 and out == exp == R.main_decode(blob)
    rc, out, frames, _ = E.decode(blob, SKIP)      # RFC mode accepts the same inputs
    assert rc == 0 and out == exp_skip


def test_c2_small_and_multiblock():
    blob, exp = corpora.c2_small(16)
    rc, out, frames, _ = E.decode(blob, Q)
    assert rc == 0 and out == exp
    blob, exp = corpora.c3_small(3 << 20)
    rc, out, frames, _ = E.decode(blob, Q)
    assert rc == 0 and out == exp


def test_rfc_only_inputs_follow_libzstd():
    """inputs the reference rejects: default mode decodes them like libzstd,
    ZSB_REFERENCE_QUIRKS rejects them like the reference"""
    for frame, plain in corpora.rfc_only():
        rc, out, frames, _ = E.decode(frame, 0)
        assert rc == 0 and all(f[0] == 0 for f in frames) and out == plain
        _, _, oerr = R.decode_frames(frame, quirks=True)
        rc, out, frames, _ = E.decode(frame, Q)
        got = rc or next((f[0] for f in frames if f[0]), 0)
        assert oerr is not None and got != 0


def test_mutations_never_crash_and_agree_when_both_accept():
    r = random.Random(11)
    n_ok = n_same_err = 0
    for name, d in corpora.mutation_sources().items():
        for _ in range(120):
            b = corpora.mutate(r, d)
            want, _, oerr = R.decode_frames(b, quirks=True)
            rc, out, frames, _ = E.decode(b, Q | SKIP)
            got = rc or next((f[0] for f in frames if f[0]), 0)
            if oerr is None and got == 0:
                assert out == want
                n_ok += 1
            elif oerr is not None and got == oerr.code:
                n_same_err += 1
            # a frame the reference accepts may be refused here only as RFC-invalid data
            if oerr is None and got:
                assert got in (100, 101, 102), (name, got)
    assert n_ok > 100 and n_same_err > 100


# ---------------------------------------------------------------- fast sequence path (zsb_seqfast.h)
def test_history_composition_is_associative():
    assert E.hist_assoc(12345, 200000) == 0


def test_fast_sequence_path_equals_careful_decoder():
    """every block decoded above also went through phase 1 + phase 2 of the fast path: identical records,
    sizes and repeat offsets, or a request for the careful decoder exactly where that one fails"""
    for name in corpora.FIXTURE_NAMES:
        E.decode(corpora.fixture(name), Q)
    E.decode(corpora.c4()[0], Q)
    E.decode(corpora.c2_small(16)[0], Q)
    E.decode(corpora.c3_small(3 << 20)[0], Q)
    r = random.Random(5)
    for name, d in corpora.mutation_sources().items():
        for _ in range(60):
            E.decode(corpora.mutate(r, d), Q | SKIP)
    ran, same, slow, diff = E.fast_stats()
    assert diff == 0 and same > 200 and slow > 5 and ran == same + slow, (ran, same, slow, diff)
    ran, same, slow, diff = E.huf_stats()          # fast Huffman stream decode (zsb_huf.h), one count per stream
    assert diff == 0 and same > 500 and slow > 5 and ran == same + slow, (ran, same, slow, diff)


def test_mutated_inputs_equal_the_oracle_exactly():
    """Error-path parity under ZSB_REFERENCE_QUIRKS: on every mutated input the CPU build of the device code + host walk gives what the
    oracle gives -- the same bytes, or the same error variant (the reference's FIRST error: section errors of a block are raised while
    the container is walked, frame.rs:210-223; literal streams are decoded until their bits run out, literals.rs:55,70-81).  The only
    tolerated difference: inputs on which the reference PANICS (oracle code 99: Huffman weights it cannot turn into a tree ...), where
    this library reports its own code >= 100."""
    import random
    import corpora
    import refcpu as R
    r = random.Random(21)
    n = n_panic = 0
    for name, d in corpora.mutation_sources().items():
        for _ in range(60):
            b = corpora.mutate(r, d)
            want, _, oerr = R.decode_frames(b, quirks=True)
            rc, out, frames, _ = E.decode(b, 4 | 1)
            got = rc or next((f[0] for f in frames if f[0]), 0)
            oc = oerr.code if oerr is not None else 0
            n += 1
            if oc == 99:
                n_panic += 1                    # whatever this library says (its own code >= 100, a later error, or the RFC's decoding)
                continue
            assert got == oc, (name, oc, got, b.hex()[:80])
            if oc == 0:
                assert out == want, name
    assert n == 540 and n_panic < 30


def test_mutated_frames_of_many_blocks_equal_the_oracle():
    """The same, on frames of 100-300 blocks (repeat-mode tables, treeless literals, raw and RLE blocks between compressed ones), where the ORDER
    of the reference's stages matters: all blocks are parsed before any is decoded (frame.rs:210-223), then Block::decode runs block after
    block, literals first (block.rs:83-85) -- a missing previous table in block k does not hide a corrupt bitstream in block k - 1.  Errors are
    compared in stream order (main.rs:42-53: frame i is decoded before frame i + 1 is parsed).  Listed exception: status 101, a block that
    regenerates more than 128 KiB (the reference has no limit; literal streams decoded to bit exhaustion can push a full block past it)."""
    import gen_corpus as G
    text = G.moby_text()
    rnd = random.Random(7)
    b = text[:150001] + bytes(rnd.getrandbits(8) for _ in range(20000)) + b"\0" * 30000 + text[300000:400003]
    srcs = [G.compress(text[:1000001], level=9, window_log=12), G.compress(b, level=3, window_log=10),
            G.compress(text[200000:1100007], level=9, window_log=13, content_size=False), corpora.fixture("moby-dick.txt.zst")]
    r = random.Random(77)
    n_same = n_big = 0
    for src in srcs:
        for _ in range(40):
            m = corpora.mutate(r, src)
            want, _, oerr = R.decode_frames(m, quirks=True)
            oc = oerr.code if oerr is not None else 0
            rc, out, frames, _ = E.decode(m, 4)
            st = [f[0] for f in frames]
            got = (next((s for s in st[:max(len(st) - 1, 0)] if s), 0) or rc) if rc else next((s for s in st if s), 0)
            if oc == 99:
                continue
            if got == 101:
                n_big += 1
                continue
            assert got == oc, (oc, got, len(m))
            if oc == 0:
                assert out == want
            n_same += 1
    assert n_same > 140 and n_big <= 4, (n_same, n_big)


def test_round1_fuzz_findings_on_the_cpu_build():
    """tests/golden/fuzz_fail_77_*.zst (see test_gpu_parity.py::test_round1_fuzz_findings_stay_fixed): the CPU build of the device code gives the
    oracle's first error (in stream order) on both"""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    for name in ("fuzz_fail_77_2417.zst", "fuzz_fail_77_3976.zst"):
        b = open(os.path.join(here, "golden", name), "rb").read()
        _, _, oerr = R.decode_frames(b, quirks=True)
        rc, out, frames, _ = E.decode(b, 4 | 1, cap=4 << 20)
        st = [f[0] for f in frames]
        got = (next((s for s in st[:max(len(st) - 1, 0)] if s), 0) or rc) if rc else next((s for s in st if s), 0)
        assert got == (oerr.code if oerr is not None else 0), name


def test_two_symbol_huffman_table_equals_single_lookups():
    """k_huf's table of two symbols per cell (zsb_huf.h): for random complete codes, every cell gives exactly what one-symbol lookups give --
    the first code under the ten bits, the second one if it fits in what is left, both children where the ten bits are the prefix of two
    11-bit codes -- and the table built from the cell starts (as the kernel does) equals the one derived from the finished one-symbol table"""
    r = random.Random(11)
    n_long = n_two = n_codes = 0
    for trial in range(300):
        mb = r.choice([4, 6, 8, 9, 10, 11, 11, 11])
        # a random complete code: split leaves of a full binary tree until there are enough symbols
        lens = [1, 1]
        want = r.randrange(2, min(120, 1 << mb) + 1)
        while len(lens) < want:
            open_ = [i for i, l in enumerate(lens) if l < mb]
            if not open_: break
            l = lens.pop(r.choice(open_)); lens += [l + 1, l + 1]
        if max(lens) != mb: continue
        syms = r.sample(range(256), len(lens))
        wts = [0] * (max(syms) + 1)
        for s_, l in zip(syms, lens): wts[s_] = mb + 1 - l
        # the last non-zero weight is implied (the description carries all but the last symbol)
        last = max(syms)
        rc, got_mb, lut, pa, pb = E.huf_pairs(wts[:last])
        assert rc == 0 and got_mb == mb, (trial, rc, got_mb, mb)
        assert pa == pb, trial
        n_codes += 1

        def one(bits, nbits):                    # one-symbol lookup on the top bits of a `nbits`-bit value, zero padded
            idx = (bits << mb >> nbits) if nbits <= mb else bits >> (nbits - mb)
            return lut[idx & ((1 << mb) - 1)]
        for x in range(1024):
            c = pa[x]
            s1, s2, l1, lng, cnt, lt = c & 0xFF, (c >> 8) & 0xFF, (c >> 16) & 15, (c >> 20) & 1, (c >> 21) & 3, c >> 24
            a_sym, a_len = one(x, 10)
            if a_len > 10:                       # ten bits do not determine the code: both 11-bit children
                assert mb == 11 and lng == 1 and cnt == 1 and l1 == 11 and lt == 11
                assert (s1, 11) == lut[2 * x] and (s2, 11) == lut[2 * x + 1]
                n_long += 1
                continue
            assert lng == 0 and s1 == a_sym and l1 == a_len
            rest_bits = 10 - a_len
            rest = x & ((1 << rest_bits) - 1)
            b_sym, b_len = one(rest, rest_bits) if rest_bits else (0, 99)
            if rest_bits and b_len <= rest_bits:
                assert cnt == 2 and s2 == b_sym and lt == a_len + b_len
                n_two += 1
            else:
                assert cnt == 1 and s2 == 0 and lt == a_len
    assert n_codes > 100 and n_long > 200 and n_two > 10000
