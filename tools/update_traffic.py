#!/usr/bin/env python
"""update_traffic.py -- regenerate profiles/traffic.json from an `ncu --set full` report of bench.py.

    python tools/update_traffic.py gpurun_out/prof.ncu-rep --workload C2 --frames 4096 [--captured "r02c, 1x B200"]

Per kernel (mean over the captured launches): dram__bytes_read.sum, dram__bytes_write.sum and their sum, plus the digest of the
sources the report was taken on (bench.csrc_digest): bench.py copies the dominant kernel's total into roofline.traffic only when
that digest equals the digest of the sources it runs, so a stale capture never passes for a measurement of the current build.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--captured", default="")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    acc = {}
    for r in rows[2:]:
        k = r[ik].split("(")[0]
        rd, wr = float(r[ir]) * scale[units[ir]], float(r[iw]) * scale[units[iw]]
        t = acc.setdefault(k, [0.0, 0.0, 0])
        t[0] += rd; t[1] += wr; t[2] += 1
    import bench
    js = {"_comment": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches) from an `ncu --set full --clock-control none` "
                      "capture of bench.py; written by tools/update_traffic.py; bench.py uses it only for the build whose digest is recorded here",
          "csrc_sha256": bench.csrc_digest(), "workload": a.workload, "frames": a.frames, "captured": a.captured,
          "kernels": {k: {"read": int(v[0] / v[2]), "write": int(v[1] / v[2]), "total": int((v[0] + v[1]) / v[2]), "launches": v[2]} for k, v in sorted(acc.items())}}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(js, f, indent=1)
    print(json.dumps(js["kernels"], indent=1))


if __name__ == "__main__":
    main()
