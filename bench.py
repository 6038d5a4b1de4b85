#!/usr/bin/env python
"""bench.py -- decompressed GB/s of the B200-native Zstandard decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C1|C2|C3|C4|C5] [--frames F]

Workload (config.workload), generated on the fly by libzstd from the moby-dick fixture with fixed seeds (tools/gen_corpus.py):
  one GPU (default)     BASELINE config C2: 4096 independent 128 KiB text frames (zstd level 3, Huffman literals + FSE sequences,
                        content size + XXH64 checksum) -- the config the metric is quoted on;
  N > 1 GPUs (default)  BASELINE config C5: ONE corpus of 65 536 such frames (8 GiB decompressed, seed 5), rank r decodes the
                        contiguous frame range r of zsb_shard_plan (8 192 frames per GPU at N = 8): strong scaling, no collective
                        on the decode path, `value` = bytes of the whole corpus / max-over-ranks device time;
  --workload C1|C3|C4   the other BASELINE configs (single frames: replicas only); C3 also reports the decode without XXH64.

A step = one pass of the whole decode path (section parse, plan, Huffman literals, FSE sequences (+ the
careful re-decode of rejected blocks), plan, raw/RLE, sequence execution, XXH64) over the batch, checksums verified.
`value`: compressed input and output resident in HBM, CUDA-event time on the launching stream.
`e2e`  : the same through the C ABI with HOST buffers (zsb_scan_decode on pinned memory): the
         host walk, H2D of the compressed bytes and D2H of the output are inside the timed region.
`cpu_baseline` / `--impl reference`: the CPU oracle (oracle/refcpu.c, a C restatement of the Rust
reference, which cannot be built here) on the box's host cores, frames spread over all threads.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

FRAME_SIZE = 131072
METRIC = "decompressed GB/s"
C5_FRAMES = 65536


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "C1", "C2", "C3", "C4", "C5"],
                    help="BASELINE.json config; auto = C2 on one GPU (the config the metric is quoted on), C5 (one 8 GiB corpus sharded by frame) on several")
    ap.add_argument("--frames", type=int, default=0, help="frames of the C2 (default 4096 per GPU) / C5 (default 65536 in all) corpus")
    ap.add_argument("--c3-bytes", type=int, default=1 << 30, help="size of the single C3 frame")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def pick_workload(args, world):
    w = args.workload
    if w == "auto":
        w = "C2" if world == 1 else "C5"
    return w


def config_of(w, args, world):
    """The `config` object of the JSON line: identical for --impl ours and --impl reference (nothing in it is measured)."""
    if w == "C2":
        f = args.frames or 4096
        return {"workload": f"C2: {f} independent 128 KiB text frames per GPU (zstd -3, Huffman literals + FSE sequences, FCS + XXH64), seed 2 (+1000 per rank)",
                "frames_per_gpu": f, "frames_total": f * world, "frame_bytes": FRAME_SIZE, "decompressed_bytes_total": f * world * FRAME_SIZE,
                "l2": "inputs larger than L2 (compressed + output + scratch >> 126 MB), no flush needed", "checksum": "verified every step",
                "parallelism": f"every rank decodes its own {f} frames, {world} GPU(s), no collective", "scaling": "weak"}
    if w == "C5":
        f = args.frames or C5_FRAMES
        return {"workload": f"C5: one corpus of {f} independent 128 KiB text frames (seed 5, zstd -3, FCS + XXH64, {f * FRAME_SIZE / 2**30:g} GiB decompressed) sharded by frame "
                            f"over the GPUs with zsb_shard_plan",
                "frames_per_gpu": f // world, "frames_total": f, "frame_bytes": FRAME_SIZE, "decompressed_bytes_total": f * FRAME_SIZE,
                "l2": "inputs larger than L2, no flush needed", "checksum": "verified every step",
                "parallelism": f"contiguous frame ranges balanced on decompressed bytes, {world} GPU(s), no collective", "scaling": "strong"}
    if w == "C1":
        return {"workload": "C1: tests/fixtures/moby-dick.txt.zst, one frame of 10 compressed blocks (500 371 -> 1 276 235 bytes), replicas only",
                "frames_per_gpu": 1, "frames_total": world, "decompressed_bytes_total": 1276235 * world,
                "l2": "the 1.8 MB working set fits L2: 256 MB are written between steps to flush it", "checksum": "verified every step",
                "parallelism": f"{world} replica(s)", "scaling": "weak"}
    if w == "C3":
        return {"workload": f"C3: one frame of {args.c3_bytes / 2**30:g} GiB (shuffled text tiles, zstd -3, windowLog 23, 128 KiB blocks, matches up to 8 MiB back, XXH64), replicas only",
                "frames_per_gpu": 1, "frames_total": world, "decompressed_bytes_total": args.c3_bytes * world,
                "l2": "inputs larger than L2, no flush needed", "checksum": "verified every step (decode-only time reported beside it)",
                "parallelism": f"{world} replica(s): a single frame does not shard", "scaling": "weak"}
    return {"workload": "C4: mixed raw / RLE / compressed blocks, skippable frames, every literal and table mode, XXH64 (seed 4), replicas only",
            "frames_per_gpu": None, "frames_total": None, "decompressed_bytes_total": None,
            "l2": "working set fits L2: 256 MB are written between steps to flush it", "checksum": "verified every step",
            "parallelism": f"{world} replica(s)", "scaling": "weak"}


def csrc_digest():
    """SHA-256 over the CUDA/C++ sources: profiles/traffic.json records the digest of the build its ncu capture profiled."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "zstd-decompressor_b200", "csrc")
    for n in sorted(os.listdir(d)):
        with open(os.path.join(d, n), "rb") as f:
            h.update(n.encode()); h.update(f.read())
    return h.hexdigest()[:16]


def build_workload(w, args, rank, world):
    """-> (compressed bytes of this rank, expected plaintext of this rank)"""
    import gen_corpus as G
    if w == "C2":
        return G.make_c2(args.frames or 4096, seed=2 + 1000 * rank)
    if w == "C5":
        import ctypes as C
        import zstd_decompressor_b200 as Z
        f = args.frames or C5_FRAMES
        # the shard plan of the C ABI over the corpus' frame descriptors (sizes are declared: 128 KiB each)
        fr = (Z.ZsbFrame * f)()
        for i in range(f):
            fr[i].kind = 0; fr[i].has_content_size = 1; fr[i].content_size = FRAME_SIZE; fr[i].src_len = 55000
        first = (C.c_size_t * (world + 1))()
        assert Z.lib().zsb_shard_plan(fr, f, world, first) == 0
        return G.make_c2_range(f, 5, first[rank], first[rank + 1])
    if w == "C1":
        return G.make_c1()
    if w == "C3":
        return G.make_c3(total=args.c3_bytes)
    blob, exp_noskip, exp_skip, labels = G.make_c4()
    return blob, exp_noskip


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons during the timed region: NVML polled every ~5 ms from a thread (the device is found by
    its UUID, so CUDA_VISIBLE_DEVICES does not matter); `nvidia-smi -lms 100` if NVML cannot be used."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nv, self.stop_flag, self.source = index, [], None, None, False, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            u = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("" if u.startswith("GPU-") else "GPU-") + u)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

    def start(self):
        try:
            self.nv, self.h = self._nvml_handle()
            self.max_sm = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                c = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                m = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((time.time(), c, [n for n, bit in names if m & bit]))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                c, self.max_sm = float(f[1]), float(f[2])
            except ValueError:
                continue
            self.rows.append((time.time(), c, [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                                               if v.lower().startswith("active")]))

    def stop(self, t0, t1):
        if not self.source:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML and nvidia-smi unavailable)"], "samples": 0}
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        self.stop_flag = True
        pad = 0.0 if self.source == "nvml" else 0.15
        inside = [(c, r) for ts, c, r in self.rows if t0 - pad <= ts <= t1 + pad]
        if not inside:                               # a timed region shorter than the sampling period: the nearest samples
            inside = [(c, r) for ts, c, r in sorted(self.rows, key=lambda x: min(abs(x[0] - t0), abs(x[0] - t1)))[:3]]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_sm", None), "reasons": ["no clock sample"], "samples": 0, "source": self.source}
        reasons = sorted({n for _, r in inside for n in r})
        return {"sm_mhz": statistics.median(c for c, _ in inside), "sm_max_mhz": getattr(self, "max_sm", None), "reasons": reasons,
                "samples": len(inside), "source": self.source}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(blob, scan_frames, target_cpu_seconds, threads):
    """Decode a bounded sample of the workload's frames with the CPU oracle on `threads` threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refcpu as R
    # calibrate on a few frames, single thread
    n_cal = min(8, len(scan_frames))
    end = scan_frames[n_cal - 1][0] + scan_frames[n_cal - 1][1]
    t = time.perf_counter(); out = R.main_decode(blob[:end]); dt1 = time.perf_counter() - t
    rate1 = len(out) / dt1
    n = int(max(threads, min(len(scan_frames), target_cpu_seconds * rate1 / FRAME_SIZE)))
    n = max(n_cal, min(n, len(scan_frames)))
    end = scan_frames[n - 1][0] + scan_frames[n - 1][1]
    sample = blob[:end]
    t = time.perf_counter(); out = R.main_decode(sample, threads=threads); dt = time.perf_counter() - t
    return {"value": len(out) / dt / 1e9, "bytes": len(out), "seconds": dt, "frames": n, "single_thread_gbs": rate1 / 1e9, "sample_blob": sample}


def libzstd_rate(blob, scan_frames, n_frames, threads):
    """Context line: libzstd 1.5.5 (the system library, through ctypes, which releases the GIL) decoding the
    first n_frames frames, frames spread over `threads` threads.  Not the reference and not the target."""
    import concurrent.futures as cf
    import gen_corpus as G
    z = G.libzstd()
    frames = [(o, l) for o, l in scan_frames[:n_frames]]
    src = C.create_string_buffer(blob, len(blob))
    base = C.addressof(src)

    def work(part):
        out = C.create_string_buffer(FRAME_SIZE + 64)
        n = 0
        for o, l in part:
            r = z.ZSTD_decompress(out, FRAME_SIZE + 64, C.c_char_p(base + o), l)
            assert not z.ZSTD_isError(r)
            n += r
        return n
    parts = [frames[i::threads] for i in range(threads)]
    t = time.perf_counter()
    with cf.ThreadPoolExecutor(threads) as ex:
        total = sum(ex.map(work, parts))
    return total / (time.perf_counter() - t) / 1e9


def cpu_sample_blob(w, args):
    """A bounded sample of the workload for the CPU arm: (compressed bytes, [(frame offset, length)], what it is)"""
    import gen_corpus as G
    import zstd_inspect as I
    cores = os.cpu_count() or 1
    if w in ("C2", "C5"):
        n = min(args.frames or (4096 if w == "C2" else C5_FRAMES), 64 * cores)
        blob, _ = G.make_c2(n, seed=2 if w == "C2" else 5)
        what = f"the first {n} frames of the corpus"
    elif w == "C3":
        blob, _ = G.make_c3(total=min(args.c3_bytes, 48 << 20))
        what = "a 48 MiB frame of the same construction (a single frame decodes on one thread)"
    elif w == "C1":
        blob, _ = G.make_c1(); what = "the whole fixture"
    else:
        blob = G.make_c4()[0]; what = "the whole corpus"
    return blob, [(f.src_off, f.src_len) for f in I.inspect(blob)], what


def run_reference(args):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    w = pick_workload(args, world)
    cores = os.cpu_count() or 1
    blob, fr, what = cpu_sample_blob(w, args)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refcpu as R
    threads = cores if len(fr) > 1 else 1
    if len(fr) > 8:                                   # size one step at about 2 s of wall clock on all cores
        cal = cpu_oracle_rate(blob, fr, target_cpu_seconds=2.0 * cores, threads=cores)
        sample, nfr, single = cal["sample_blob"], cal["frames"], cal["single_thread_gbs"]
    else:
        sample, nfr, single = blob, len(fr), None
    for _ in range(args.warmup if len(fr) > 8 else min(args.warmup, 1)):
        R.main_decode(sample, threads=threads)
    steps = args.steps if len(fr) > 8 else max(1, min(args.steps, 3))
    t = time.perf_counter(); nbytes = 0
    for _ in range(steps):
        nbytes += len(R.main_decode(sample, threads=threads))
    dt = time.perf_counter() - t
    val = nbytes / dt / 1e9
    cfg = config_of(w, args, world)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port",
                             "sample": f"{what}: {nfr} frame(s) per step ({nbytes // steps} bytes out), oracle/refcpu.c (C restatement of the Rust reference, which cannot be built in "
                                       f"this image) frame-parallel on {threads} thread(s)" + (f"; single thread {single:.4f} GB/s" if single else "")},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import hashlib
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import zstd_decompressor_b200 as Z

    w = pick_workload(args, world)
    cfg = config_of(w, args, world)
    blob, expect = build_workload(w, args, rank, world)
    n_in, n_out = len(blob), len(expect)
    want_sha = hashlib.sha256(expect).digest()
    flags = Z.VERIFY_CHECKSUM | Z.REFERENCE_QUIRKS
    ctx = Z.Context(local)
    dec = Z.Decoder(ctx)
    # a dedicated (non-default) stream: zsb_ctx_set_stream(NULL) would select the context's own stream and
    # the events below would not bracket the kernels
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    small = (n_in + n_out) < (200 << 20)               # working set fits L2: flush between steps
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    # ---- resident arm: compressed bytes and output stay in HBM
    host_src = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
    d_src = torch.empty(n_in + 128, dtype=torch.uint8, device="cuda")
    d_src[:n_in].copy_(host_src); d_src[n_in:].zero_()
    d_dst = torch.empty(n_out + 64, dtype=torch.uint8, device="cuda")
    scan = Z.Scan(blob, flags)
    assert scan.status == 0
    frames_n = scan.n_frames
    n_checked = sum(1 for i in range(frames_n) if scan.frames[i].kind == 0 and scan.frames[i].has_checksum)

    def resident(fl, steps, warmup, profile):
        """K launches of the whole path with flags fl; -> (ms per step, per-kernel averages, launches averaged)"""
        dec.prepare(d_src.data_ptr(), n_in, scan, d_dst.data_ptr(), n_out, fl | Z.SRC_ON_DEVICE | Z.DST_ON_DEVICE)
        dec.launch(); res = dec.finish()                       # also sizes the scratch exactly
        err = res.first_error()
        assert err is None and res.total.value == n_out, f"decode failed: {err}"
        if fl & Z.VERIFY_CHECKSUM:
            assert sum(res.checksum_ok[i] for i in range(frames_n)) == n_checked, "stored XXH64 mismatch"
        got = d_dst[:n_out].cpu().numpy().tobytes()
        assert hashlib.sha256(got).digest() == want_sha, "GPU output differs from the plaintext"
        del got
        torch.cuda.synchronize()
        for _ in range(warmup):
            dec.launch()
        torch.cuda.synchronize()
        if profile:
            ctx.set_profile(True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        if flush_buf is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                dec.launch()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        else:                                           # L2 flushed before every step, each step timed on its own
            with torch.cuda.stream(stream):
                for _ in range(steps):
                    flush_buf.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream); dec.launch(); e1.record(stream)
                    torch.cuda.synchronize()
                    ms += e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
        kt, nl = (ctx.kernel_times_avg() if profile else ([], 0))
        if profile:
            ctx.set_profile(False)
        res = dec.finish()
        assert res.first_error() is None
        return ms / steps, kt, nl

    sampler = ClockSampler(local); sampler.start()
    t0 = time.time()
    ms_step, ktimes, nl = resident(flags, args.steps, args.warmup, True)
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    launches_per_step = ctx.last_launch_count()
    decode_only = None
    if w == "C3":                                       # the serial XXH64 of a single huge frame is reported apart from the decode
        ms_d, kt_d, _ = resident(Z.REFERENCE_QUIRKS, max(2, args.steps // 4), 1, True)
        dom_d = max(kt_d, key=lambda kv: kv[1]) if kt_d else ("none", 0.0)
        decode_only = {"value": n_out / (ms_d * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_d,
                       "what": "the same frame without ZSB_VERIFY_CHECKSUM (no k_xxh_one: XXH64 of one frame is a single dependent chain, 33.5 M rounds for 1 GiB)",
                       "kernels_ms": {k: round(v, 4) for k, v in kt_d},
                       "dominant_kernel": dom_d[0], "dominant_kernel_ms": dom_d[1],
                       "algorithmic_gbs_of_dominant_kernel": (n_in + n_out) / (dom_d[1] * 1e-3) / 1e9 if dom_d[1] > 0 else None}

    # ---- the walk over frames and blocks that precedes the decode (outside the timed region above: `scan` is made once): on the device for the
    #      resident buffer (zsb_scan_device, same descriptors), on the host for the host copy; wall clock per call, best of 5, reported beside `value`
    walk = None
    try:
        def best_ms(fn, reps=5):
            ts = []
            for _ in range(reps):
                torch.cuda.synchronize(); a = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - a) * 1e3)
            return min(ts), r
        dms, dscan = best_ms(lambda: Z.DeviceScan(ctx, d_src.data_ptr(), n_in, flags))
        hms, _ = best_ms(lambda: Z.Scan(blob, flags))
        assert (dscan.status, dscan.n_frames, dscan.n_blocks) == (scan.status, scan.n_frames, scan.n_blocks), "zsb_scan_device differs from zsb_scan"
        walk = {"device_ms": round(dms, 4), "host_ms": round(hms, 4), "frames": scan.n_frames, "blocks": scan.n_blocks,
                "what": "zsb_scan_device on the resident buffer / zsb_scan on the host copy; not inside ms_per_step (descriptors are made once), inside e2e (host walk, overlapped)"}
        del dscan
    except Z.ZsbError as e:
        walk = {"error": str(e)}

    # ---- end to end arm: host buffers through the C ABI (scan + H2D + kernels + D2H every step)
    host_dst = torch.empty(n_out + 64, dtype=torch.uint8).pin_memory()
    ctx2 = Z.Context(local)
    src_ptr, dst_ptr = host_src.data_ptr(), host_dst.data_ptr()

    def e2e_step():
        # one call: host walk, uploads, kernels, downloads (the walk overlaps the GPU work, zsb_scan_decode in include/zsb.h)
        sd = Z.ScanDecode(ctx2, (src_ptr, n_in), (dst_ptr, n_out), flags)
        assert sd.status == 0 and sd.total == n_out and sd.n_frames == frames_n and sd.first_error() is None
        return sd
    e2e_steps = args.e2e_steps if n_out < (2 << 30) else min(args.e2e_steps, 3)
    if e2e_steps > 0:
        e2e_step(); e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - te) * 1e3
    if e2e_steps > 0:
        assert hashlib.sha256(host_dst[:n_out].numpy().tobytes()).digest() == want_sha
    desc_bytes = scan.n_frames * C.sizeof(Z.ZsbFrame) + scan.n_blocks * C.sizeof(Z.ZsbBlock)

    # ---- reduce over ranks: max time, summed bytes
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])
        b = torch.tensor([n_out, n_in], dtype=torch.float64, device="cuda")
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        tot_out, tot_in = float(b[0]), float(b[1])
    else:
        tot_out, tot_in = float(n_out), float(n_in)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = tot_out / (ms_step * 1e-3) / 1e9
    e2e_val = tot_out / (e2e_ms / e2e_steps * 1e-3) / 1e9 if e2e_steps > 0 else None      # --e2e-steps 0: kernel profiling runs
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    b_alg = n_in + n_out                                   # every compressed byte read once, every output byte written once (this rank)
    dom = max(ktimes, key=lambda kv: kv[1]) if ktimes else ("none", 0.0)
    ach = b_alg / (dom[1] * 1e-3) / 1e9 if dom[1] > 0 else 0.0
    # DRAM bytes per launch of that kernel: only from an ncu capture of THIS build and THIS workload (tools/update_traffic.py writes the
    # digest of the sources next to the figures); otherwise null
    traffic, traffic_note = None, "no ncu capture of this build (profiles/traffic.json)"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("csrc_sha256") != csrc_digest():
            traffic_note = "profiles/traffic.json was captured on other sources than this build"
        elif tj.get("workload") != w or (w in ("C2", "C5") and tj.get("frames") != frames_n):
            traffic_note = f"profiles/traffic.json was captured on {tj.get('workload')} / {tj.get('frames')} frames"
        else:
            ks = tj.get("kernels", {})                      # (ncu names the template instance: "void k_seq_t<16, 128, 2, 0>" is k_seq)
            ent = ks.get(dom[0]) or next((v for k, v in ks.items() if (dom[0] + "_t<") in k), {})
            traffic = ent.get("total")
            traffic_note = f"ncu --set full capture {tj.get('captured')}, dram__bytes_read.sum + dram__bytes_write.sum of {dom[0]}"
    except Exception:
        pass
    measured = {"compressed_bytes_this_gpu": n_in, "decompressed_bytes_this_gpu": n_out, "frames_this_gpu": frames_n, "compressed_bytes_all_gpus": int(tot_in),
                "decompressed_bytes_all_gpus": int(tot_out)}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": cfg, "workload_measured": measured,
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": int(n_in + desc_bytes), "d2h_bytes_per_step": int(n_out),
                "ms_per_step": e2e_ms / e2e_steps if e2e_steps > 0 else None, "steps": e2e_steps,
                "path": "zsb_scan_decode on pinned host buffers (host walk, uploads, kernels and downloads of successive shards overlapped)"},
        "gpu_launches": launches_per_step * args.steps, "walk": walk,
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None, "traffic": traffic,
                     "traffic_source": traffic_note, "algorithmic_bytes_per_launch": b_alg, "kernel_ms": dom[1], "launches_averaged": nl, "peak_source": peak_src},
        "roofline_pipeline": {"achieved": b_alg / (ms_step * 1e-3) / 1e9, "frac": b_alg / (ms_step * 1e-3) / 1e9 / peak, "frac_of_8TBs_nominal": b_alg / (ms_step * 1e-3) / 8e12},
        "kernels_ms": {k: round(v, 4) for k, v in ktimes},
    }
    if decode_only:
        line["decode_only"] = decode_only
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        if w in ("C2", "C5"):
            fr = [(scan.frames[i].src_off, scan.frames[i].src_len) for i in range(scan.n_frames)]
            cb = cpu_oracle_rate(blob, fr, target_cpu_seconds=15.0, threads=cores)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "GB/s", "cores": cores, "kind": "port",
                                    "sample": f"first {cb['frames']} frames of the workload ({cb['bytes']} bytes out) by oracle/refcpu.c, frame-parallel on {cores} threads "
                                              f"({cb['seconds']:.2f} s wall); single thread {cb['single_thread_gbs']:.4f} GB/s"}
            try:   # context only: the production CPU decoder on the same frames
                line["cpu_baseline"]["context_libzstd_1_5_5"] = {"unit": "GB/s", "threads_1": round(libzstd_rate(blob, fr, 512, 1), 3),
                                                                   f"threads_{cores}": round(libzstd_rate(blob, fr, min(frames_n, 8192), cores), 3)}
            except Exception as e:
                line["cpu_baseline"]["context_libzstd_1_5_5"] = {"unavailable": str(e)[:80]}
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import refcpu as R
            sblob, sfr, what = cpu_sample_blob(w, args)
            thr = cores if len(sfr) > 1 else 1
            t = time.perf_counter(); out = R.main_decode(sblob, threads=thr); dt = time.perf_counter() - t
            line["cpu_baseline"] = {"value": len(out) / dt / 1e9, "unit": "GB/s", "cores": thr, "kind": "port",
                                    "sample": f"{what} ({len(out)} bytes out) by oracle/refcpu.c on {thr} thread(s), {dt:.2f} s wall"}
            try:
                import gen_corpus as G
                t = time.perf_counter(); G.libzstd_decompress(blob, n_out) if len(sfr) == 1 and w != "C4" else None; dz = time.perf_counter() - t
                if len(sfr) == 1 and w != "C4":
                    line["cpu_baseline"]["context_libzstd_1_5_5"] = {"unit": "GB/s", "threads_1": round(n_out / dz / 1e9, 3), "what": "the whole frame, one thread"}
            except Exception as e:
                line["cpu_baseline"]["context_libzstd_1_5_5"] = {"unavailable": str(e)[:80]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there) must not add to it: file
    descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved original."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    _claim_stdout()
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
