#!/usr/bin/env python
"""CLI-level comparison on one host: `slow5tools-b200 view` (GPU codec) next to the unmodified reference
`slow5tools view -t <ncores>` (oracle/_ref/slow5tools_ref) on the same synthetic BLOW5 file -- the comparison
BASELINE.json's north_star asks for.  Encode = none/none -> REC+SIG (default zlib+svb-zd), decode = REC+SIG -> none/none.

    python tools/bench_view.py [--reads 100000] [--samples 4096] [--dir /dev/shm]
"""
import argparse
import json
import os
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")


def write_blow5(path, sig, n_reads, n_samples, seed=1):
    """none/none BLOW5 0.2.0, no aux fields, one read group (layout: SURVEY 8a)."""
    text = ("@asic_id\t0\n@exp_start_time\t2026-01-01T00:00:00Z\n@flow_cell_id\tSYNTH\n@sample_frequency\t4000\n"
            "#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n"
            "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n").encode()
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        hdr = bytearray(68)
        hdr[0:6] = b"BLOW5\x01"
        hdr[6:9] = bytes([0, 2, 0])
        hdr[9] = 0
        hdr[10:14] = struct.pack("<I", 1)
        hdr[14] = 0
        hdr[64:68] = struct.pack("<I", len(text))
        f.write(hdr)
        f.write(text)
        fixed = struct.pack("<I4d", 0, 8192.0, 9.0, 1444.86, 4000.0)
        ids = rng.integers(0, 2**32, (n_reads, 4), dtype=np.uint64)
        for r in range(n_reads):
            rid = ("%08x-%04x-%04x-%04x-%012x" % (ids[r, 0], ids[r, 1] & 0xffff, ids[r, 2] & 0xffff, ids[r, 3] & 0xffff,
                                                     (int(ids[r, 0]) << 16 | int(ids[r, 1])) & 0xffffffffffff)).encode()
            body = struct.pack("<H", len(rid)) + rid + fixed + struct.pack("<Q", n_samples) + sig[r * n_samples:(r + 1) * n_samples].tobytes()
            f.write(struct.pack("<Q", len(body)))
            f.write(body)
        f.write(b"5WOLB")


def timed(cmd):
    t0 = time.perf_counter()
    subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--methods", default="zlib:svb-zd", help="comma separated REC:SIG output methods to time")
    ap.add_argument("--degrade", type=int, default=0, help="also time `degrade -b BITS` of the zlib:svb-zd file with both tools")
    a = ap.parse_args()
    from slow5tools_b200 import synth
    cores = os.cpu_count() or 1
    raw = os.path.join(a.dir, "s5b_raw.blow5")
    sig = synth.nanopore_signal(a.reads * a.samples, seed=42).numpy()
    write_blow5(raw, sig, a.reads, a.samples)
    out = {"reads": a.reads, "samples": a.samples, "cores": cores, "raw_bytes": os.path.getsize(raw), "methods": {}}
    z_ref, z_ours = os.path.join(a.dir, "s5b_ref.blow5"), os.path.join(a.dir, "s5b_ours.blow5")
    back = os.path.join(a.dir, "s5b_back.blow5")
    raw_bytes = open(raw, "rb").read()
    for combo in a.methods.split(","):
        rec_m, sig_m = combo.split(":")
        res = {}
        for name, exe, z in (("reference", REF, z_ref), ("ours", CLI, z_ours)):
            if not os.path.exists(exe):
                continue
            k = "4096" if name == "reference" else "20000"
            enc = min(timed([exe, "view", "-t", str(cores), "-K", k, raw, "-c", rec_m, "-s", sig_m, "-o", z]) for _ in range(a.repeat))
            dec = min(timed([exe, "view", "-t", str(cores), "-K", k, z, "-c", "none", "-s", "none", "-o", back])
                      for _ in range(a.repeat))
            same = open(back, "rb").read() == raw_bytes
            res[name] = {"encode_s": enc, "decode_s": dec, "encode_reads_per_s": a.reads / enc, "decode_reads_per_s": a.reads / dec,
                         "compressed_bytes": os.path.getsize(z), "roundtrip_identical": same}
        if "reference" in res and "ours" in res:
            # cross-check: the reference decodes OUR file to the original bytes
            subprocess.check_call([REF, "view", "-t", str(cores), z_ours, "-c", "none", "-s", "none", "-o", back], stderr=subprocess.DEVNULL)
            res["reference_reads_our_file"] = open(back, "rb").read() == raw_bytes
            res["size_vs_reference"] = res["ours"]["compressed_bytes"] / res["reference"]["compressed_bytes"]
            res["encode_speedup"] = res["reference"]["encode_s"] / res["ours"]["encode_s"]
            res["decode_speedup"] = res["reference"]["decode_s"] / res["ours"]["decode_s"]
        out["methods"][combo] = res
    if a.degrade and "zlib:svb-zd" in out["methods"] and os.path.exists(REF):
        # `degrade -b BITS` (lossy; default output zlib + ex-zd) of the reference-written zlib + svb-zd file, both tools; the two
        # outputs must decode to the same samples
        subprocess.check_call([REF, "view", "-t", str(cores), raw, "-o", z_ref], stderr=subprocess.DEVNULL)
        d_ref, d_ours = os.path.join(a.dir, "s5b_dg_ref.blow5"), os.path.join(a.dir, "s5b_dg_ours.blow5")
        t_ref = min(timed([REF, "degrade", "-t", str(cores), "-b", str(a.degrade), z_ref, "-o", d_ref]) for _ in range(a.repeat))
        t_ours = min(timed([CLI, "degrade", "-t", str(cores), "-b", str(a.degrade), z_ref, "-o", d_ours]) for _ in range(a.repeat))
        flat = []
        for f in (d_ref, d_ours):
            subprocess.check_call([CLI, "view", f, "-c", "none", "-s", "ex-zd", "-o", back], stderr=subprocess.DEVNULL)
            flat.append(open(back, "rb").read())
        out["degrade"] = {"bits": a.degrade, "reference_s": t_ref, "ours_s": t_ours, "speedup": t_ref / t_ours,
                          "reference_bytes": os.path.getsize(d_ref), "ours_bytes": os.path.getsize(d_ours),
                          "records_identical_once_unwrapped": flat[0] == flat[1]}
        for p in (d_ref, d_ours):
            os.remove(p)
    for p in (raw, z_ref, z_ours, back):
        if os.path.exists(p):
            os.remove(p)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
