"""Builds libzsb.so (CUDA kernels + C ABI) in-tree for sm_100a.  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libzsb.so")
SOURCES = ["zsb_kernels.cu", "zsb_host.cu", "zsb_multi.cu", "zsb_dscan.cu", "zsb_scan.cpp"]
HEADERS = ["zsb_common.h", "zsb_bits.h", "zsb_fse.h", "zsb_huf.h", "zsb_seq.h", "zsb_seqfast.h", "zsb_stream.h", "zsb_bulk.cuh", "zsb_parse.h", "zsb_kernels.h", "zsb_scan.h", "zsb_walk.h", "../../include/zsb.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the zsb CUDA library cannot be built")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


CLI_SRC = os.path.join(os.path.dirname(HERE), "cli", "zstd_decompressor.cpp")
CLI_OUT = os.path.join(os.path.dirname(HERE), "cli", "zstd-decompressor")


def build_cli(force=False):
    """The command line front end (cli/zstd_decompressor.cpp), linked against libzsb.so next to the package."""
    build()
    if not force and os.path.exists(CLI_OUT) and os.path.getmtime(CLI_OUT) > max(os.path.getmtime(CLI_SRC), os.path.getmtime(OUT)):
        return CLI_OUT
    cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-o", CLI_OUT, CLI_SRC, OUT, "-Wl,-rpath," + HERE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return CLI_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
