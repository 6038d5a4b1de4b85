"""Frames shard across ranks by contiguous frame ranges: the host-side plan and extraction through the
C ABI, and the world_size-2 flow over gloo on CPU (the oracle stands in for the per-rank GPU decode: there is no GPU
here and the library has no CPU path).  The real two-GPU run is in test_gpu_parity.py."""
import os
import socket
import sys

import pytest

import corpora
import refcpu as R
import zstd_decompressor_b200 as Z


def test_plan_covers_all_frames_and_balances_bytes():
    blob, exp = corpora.c2_small(64)
    sc = Z.Scan(blob, Z.REFERENCE_QUIRKS)
    for world in (1, 2, 3, 4, 8, 64, 100):
        first = Z.shard_plan(sc, world)
        assert first[0] == 0 and first[-1] == sc.n_frames and all(a <= b for a, b in zip(first, first[1:]))
        sizes = [sum(sc.frames[f].content_size for f in range(a, b)) for a, b in zip(first, first[1:])]
        assert sum(sizes) == len(exp)
        if world <= 8:
            assert max(sizes) - min(sizes) <= 131072            # equal frames: shards differ by at most one frame


def test_plan_with_uneven_frames_and_skippables():
    blob, exp, exp_skip, parts = corpora.c4()
    sc = Z.Scan(blob, 0)
    first = Z.shard_plan(sc, 4)
    assert first[0] == 0 and first[-1] == sc.n_frames
    w = [sc.frames[f].content_size if sc.frames[f].has_content_size else sc.frames[f].src_len for f in range(sc.n_frames)]
    loads = [sum(w[a:b]) for a, b in zip(first, first[1:])]
    assert max(loads) <= sum(w) / 4 + max(w)                    # no shard exceeds its fair share by more than one frame


def test_extract_rebases_descriptors():
    blob, exp = corpora.c2_small(16)
    sc = Z.Scan(blob, Z.REFERENCE_QUIRKS)
    sh = Z.Shard(sc, 5, 11)
    assert sh.n_frames == 6 and sh.src_off == sc.frames[5].src_off
    assert sh.src_len == sc.frames[10].src_off + sc.frames[10].src_len - sc.frames[5].src_off
    sub = blob[sh.src_off:sh.src_off + sh.src_len]
    sc2 = Z.Scan(sub, Z.REFERENCE_QUIRKS)                       # scanning the sub-buffer gives the same descriptors
    assert sc2.n_frames == 6 and sc2.n_blocks == sh.n_blocks
    for i in range(6):
        a, b = sh.frames[i], sc2.frames[i]
        assert (a.src_off, a.src_len, a.first_block, a.n_blocks, a.content_size, a.stored_checksum) == \
               (b.src_off, b.src_len, b.first_block, b.n_blocks, b.content_size, b.stored_checksum)
    for i in range(sh.n_blocks):
        a, b = sh.blocks[i], sc2.blocks[i]
        assert (a.src_off, a.size, a.frame, a.type, a.last) == (b.src_off, b.size, b.frame, b.type, b.last)
    empty = Z.Shard(sc, 3, 3)
    assert empty.n_frames == 0 and empty.src_len == 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob, exp = corpora.c2_small(12)
        blob = corpora.fixture("welcome.zst") + blob + corpora.fixture("romeo3.txt.zst")
        sc = Z.Scan(blob, Z.REFERENCE_QUIRKS)
        first = Z.shard_plan(sc, world)
        sh = Z.Shard(sc, first[rank], first[rank + 1])
        # the per-rank decode (the GPU path on a B200; here the oracle on the shard's own sub-buffer)
        local = R.main_decode(blob[sh.src_off:sh.src_off + sh.src_len])
        whole = Z.gather_outputs(local)
        q.put((rank, first, len(local), whole == R.main_decode(blob)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                   # same plan on every rank
    assert res[0][2] > 0 and res[1][2] > 0 and all(r[3] for r in res)
