"""dev: where does s5b_blow5_recode_host spend its time for zstd vs zlib input"""
import os, struct, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import slow5tools_b200 as s5
from slow5tools_b200 import synth
from slow5tools_b200._capi import METHOD
import bench_view
R, N = 20000, 4096
raw = "/dev/shm/rt_raw.blow5"
bench_view.write_blow5(raw, synth.nanopore_signal(R * N, seed=42).numpy(), R, N)
def records(path):
    b = open(path, "rb").read(); pos = 68 + struct.unpack_from("<I", b, 64)[0]; out = []
    while b[pos:pos+5] != b"5WOLB":
        sz = struct.unpack_from("<Q", b, pos)[0]; out.append(b[pos+8:pos+8+sz]); pos += 8 + sz
    return out
cdc = s5.Codec(0)
for name, m in (("zlib", METHOD.ZLIB), ("zstd", METHOD.ZSTD)):
    z = "/dev/shm/rt_%s.blow5" % name
    subprocess.check_call([bench_view.CLI, "view", raw, "-c", name, "-s", "svb-zd", "-o", z], stderr=subprocess.DEVNULL)
    recs = records(z)
    for rep in range(3):
        t0 = time.perf_counter(); rc, img = cdc.blow5_recode(m, METHOD.SVB_ZD, METHOD.NONE, METHOD.NONE, recs); t1 = time.perf_counter()
        print(name, "decode recode call %.1f ms (includes python staging) rc=%d out=%d" % ((t1 - t0) * 1e3, rc, len(img)))
    t0 = time.perf_counter(); rc, outs = cdc.depress_batch(m, recs); t1 = time.perf_counter()
    print(name, "depress_batch %.1f ms" % ((t1 - t0) * 1e3))
