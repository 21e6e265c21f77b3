#!/usr/bin/env python
"""CLI-level comparison of `index` on one host: slow5tools-b200 index next to the reference's slow5tools index on the same
zlib+svb-zd BLOW5 file (written by our view, page cache warm).  The .idx files must be identical.

    python tools/bench_index.py [--reads 300000] [--samples 4096] [--dir /dev/shm]
"""
import argparse
import filecmp
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_view  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=300000)
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    from slow5tools_b200 import synth
    raw = os.path.join(a.dir, "s5b_idx_raw.blow5")
    bench_view.write_blow5(raw, synth.nanopore_signal(a.reads * a.samples, seed=42).numpy(), a.reads, a.samples)
    out = {"reads": a.reads, "samples": a.samples, "cores": os.cpu_count()}
    for method in ("zlib", "zstd", "none"):
        z = os.path.join(a.dir, "s5b_idx_%s.blow5" % method)
        subprocess.check_call([bench_view.CLI, "view", raw, "-c", method, "-s", "svb-zd" if method != "none" else "none", "-o", z],
                              stderr=subprocess.DEVNULL)
        z2 = z + ".copy.blow5"
        shutil.copy(z, z2)
        res = {"file_bytes": os.path.getsize(z)}
        for name, exe, f in (("reference", bench_view.REF, z2), ("ours", bench_view.CLI, z)):
            if not os.path.exists(exe):
                continue
            best = 1e9
            for _ in range(a.repeat):
                if os.path.exists(f + ".idx"):
                    os.remove(f + ".idx")
                t0 = time.perf_counter()
                subprocess.check_call([exe, "index", f], stderr=subprocess.DEVNULL)
                best = min(best, time.perf_counter() - t0)
            res[name + "_s"] = best
        if "reference_s" in res and "ours_s" in res:
            res["identical_idx"] = filecmp.cmp(z + ".idx", z2 + ".idx", shallow=False)
            res["speedup"] = res["reference_s"] / res["ours_s"]
        out[method] = res
        for p in (z, z2, z + ".idx", z2 + ".idx"):
            if os.path.exists(p):
                os.remove(p)
    os.remove(raw)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
