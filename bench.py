#!/usr/bin/env python
"""bench.py -- decompressed GB/s of the B200-native Zstandard decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F]

Workload (config.workload): BASELINE config C2 per GPU -- F = 4096 independent 128 KiB text frames
(zstd level 3, Huffman literals + FSE sequences, content size + XXH64 checksum), generated on the fly
by libzstd from the moby-dick fixture with fixed seeds (tools/gen_corpus.py).  With N > 1 GPUs every
rank decodes its own F-frame shard (frames shard by frame, no collective on the decode path): weak
scaling, `value` = bytes all ranks produced / max-over-ranks device time.

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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per GPU (BASELINE C2: 4096)")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(frames):
    return f"C2: {frames} independent 128 KiB text frames per GPU (zstd -3, Huffman literals + FSE sequences, FCS + XXH64)"


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import gen_corpus as G
    import zstd_inspect as I
    cores = os.cpu_count() or 1
    frames = min(args.frames, 64 * cores)
    blob, _ = G.make_c2(frames, seed=2)
    fr = [(f.src_off, f.src_len) for f in I.inspect(blob)]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refcpu as R
    # size one step at about 2 s of wall clock on all cores
    cal = cpu_oracle_rate(blob, fr, target_cpu_seconds=2.0 * cores, threads=cores)
    sample = cal["sample_blob"]
    for _ in range(args.warmup):
        R.main_decode(sample, threads=cores)
    t = time.perf_counter(); nbytes = 0
    for _ in range(args.steps):
        nbytes += len(R.main_decode(sample, threads=cores))
    dt = time.perf_counter() - t
    val = nbytes / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args.frames), "frames_per_gpu": args.frames},
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": f"{cal['frames']} of the workload's frames per step ({cal['bytes']} bytes out), oracle/refcpu.c frame-parallel on {cores} threads; "
                                       f"single thread {cal['single_thread_gbs']:.4f} GB/s; the Rust reference itself cannot be built in this image"},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
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
    import gen_corpus as G
    import zstd_decompressor_b200 as Z

    frames_n = args.frames
    blob, expect = G.make_c2(frames_n, seed=2 + 1000 * rank)
    n_in, n_out = len(blob), len(expect)
    flags = Z.VERIFY_CHECKSUM | Z.REFERENCE_QUIRKS
    ctx = Z.Context(local)
    dec = Z.Decoder(ctx)
    # a dedicated (non-default) stream: zsb_ctx_set_stream(NULL) would select the context's own stream and
    # the events below would not bracket the kernels
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    # ---- resident arm: compressed bytes and output stay in HBM
    host_src = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
    d_src = torch.empty(n_in + 64, dtype=torch.uint8, device="cuda")
    d_src[:n_in].copy_(host_src); d_src[n_in:].zero_()
    d_dst = torch.empty(n_out + 64, dtype=torch.uint8, device="cuda")
    scan = Z.Scan(blob, flags)
    assert scan.status == 0 and scan.n_frames == frames_n
    dec.prepare(d_src.data_ptr(), n_in, scan, d_dst.data_ptr(), n_out, flags | Z.SRC_ON_DEVICE | Z.DST_ON_DEVICE)
    dec.launch(); res = dec.finish()                       # also sizes the scratch exactly
    err = res.first_error()
    assert err is None and res.total.value == n_out, f"decode failed: {err}"
    assert all(res.checksum_ok[i] for i in range(frames_n)), "stored XXH64 mismatch"
    import hashlib
    got = d_dst[:n_out].cpu().numpy().tobytes()
    assert hashlib.sha256(got).digest() == hashlib.sha256(expect).digest(), "GPU output differs from the plaintext"
    del got
    launches_per_step = ctx.last_launch_count()

    torch.cuda.synchronize()
    for _ in range(args.warmup):
        dec.launch()
    torch.cuda.synchronize()
    ctx.set_profile(True)
    sampler = ClockSampler(local); sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        dec.launch()
    e1.record(stream)
    torch.cuda.synchronize()
    t1 = time.time()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1)
    ktimes, nl = ctx.kernel_times_avg()
    ctx.set_profile(False)
    res = dec.finish()
    assert res.first_error() is None

    # ---- end to end arm: host buffers through the C ABI (scan + H2D + kernels + D2H every step)
    host_dst = torch.empty(n_out + 64, dtype=torch.uint8).pin_memory()
    ctx2 = Z.Context(local)
    L = Z.lib()
    src_ptr, dst_ptr = host_src.data_ptr(), host_dst.data_ptr()

    def e2e_step():
        # one call: host walk, uploads, kernels, downloads (the walk overlaps the GPU work, zsb_scan_decode in include/zsb.h)
        sd = Z.ScanDecode(ctx2, (src_ptr, n_in), (dst_ptr, n_out), flags)
        assert sd.status == 0 and sd.total == n_out and sd.n_frames == frames_n and sd.first_error() is None
        return sd
    if args.e2e_steps > 0:
        e2e_step(); e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    te = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - te) * 1e3
    if args.e2e_steps > 0:
        assert hashlib.sha256(host_dst[:n_out].numpy().tobytes()).digest() == hashlib.sha256(expect).digest()
    desc_bytes = scan.n_frames * C.sizeof(Z.ZsbFrame) + scan.n_blocks * C.sizeof(Z.ZsbBlock)

    # ---- reduce over ranks: max time, summed bytes
    if world > 1:
        t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms = float(t[0]), float(t[1])
        b = torch.tensor([n_out, n_in], dtype=torch.float64, device="cuda")
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        tot_out, tot_in = float(b[0]), float(b[1])
    else:
        tot_out, tot_in = float(n_out), float(n_in)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = tot_out / (ms_step * 1e-3) / 1e9
    e2e_val = tot_out / (e2e_ms / args.e2e_steps * 1e-3) / 1e9 if args.e2e_steps > 0 else None      # --e2e-steps 0: kernel profiling runs
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    b_alg = n_in + n_out                                   # every compressed byte read once, every output byte written once
    dom = max(ktimes, key=lambda kv: kv[1]) if ktimes else ("none", 0.0)
    ach = b_alg / (dom[1] * 1e-3) / 1e9 if dom[1] > 0 else 0.0
    traffic = None                                         # DRAM bytes per launch of that kernel from the committed ncu --set full capture
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom[0], {}).get("total")
        if traffic is not None and frames_n != 4096:
            traffic = None                                 # captured on the 4 096-frame workload only
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(frames_n), "frames_per_gpu": frames_n, "compressed_bytes_per_gpu": n_in, "decompressed_bytes_per_gpu": n_out,
                   "l2": "inputs larger than L2 (compressed + output + scratch >> 126 MB), no flush needed", "checksum": "verified every step",
                   "parallelism": f"frames sharded over {world} GPU(s), no collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": int(n_in + desc_bytes), "d2h_bytes_per_step": int(n_out),
                "ms_per_step": e2e_ms / args.e2e_steps if args.e2e_steps > 0 else None, "steps": args.e2e_steps,
                "path": "zsb_scan_decode on pinned host buffers (host walk, uploads, kernels and downloads of successive shards overlapped)"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None, "traffic": traffic,
                     "algorithmic_bytes_per_launch": b_alg, "kernel_ms": dom[1], "launches_averaged": nl, "peak_source": peak_src},
        "roofline_pipeline": {"achieved": b_alg / (ms_step * 1e-3) / 1e9, "frac": b_alg / (ms_step * 1e-3) / 1e9 / peak, "frac_of_8TBs_nominal": b_alg / (ms_step * 1e-3) / 8e12},
        "kernels_ms": {k: round(v, 4) for k, v in ktimes},
    }
    if not args.no_cpu_baseline and world == 1:
        import zstd_inspect as I
        fr = [(f.src_off, f.src_len) for f in I.inspect(blob[:8 << 20])] if False else None
        fr = [(scan.frames[i].src_off, scan.frames[i].src_len) for i in range(scan.n_frames)]
        cores = os.cpu_count() or 1
        cb = cpu_oracle_rate(blob, fr, target_cpu_seconds=15.0, threads=cores)
        line["cpu_baseline"] = {"value": cb["value"], "unit": "GB/s", "cores": cores, "kind": "port",
                                "sample": f"first {cb['frames']} frames of the workload ({cb['bytes']} bytes out) by oracle/refcpu.c, frame-parallel on {cores} threads "
                                          f"({cb['seconds']:.2f} s wall); single thread {cb['single_thread_gbs']:.4f} GB/s"}
        try:   # context only: the production CPU decoder on the same frames
            line["cpu_baseline"]["context_libzstd_1_5_5"] = {"unit": "GB/s", "threads_1": round(libzstd_rate(blob, fr, 512, 1), 3),
                                                               f"threads_{cores}": round(libzstd_rate(blob, fr, frames_n, cores), 3)}
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
