"""dev: find records that do not survive raw -> zlib+ex-zd -> raw through the CLI"""
import os, struct, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from slow5tools_b200 import synth
import bench_view
CLI = bench_view.CLI
R, N = int(sys.argv[1]), 4096
d = "/dev/shm"
raw, z, back = d + "/dbg_raw.blow5", d + "/dbg_z.blow5", d + "/dbg_back.blow5"
sig = synth.nanopore_signal(R * N, seed=42).numpy()
bench_view.write_blow5(raw, sig, R, N)
for rec_m in sys.argv[2].split(","):
  for slow in ("", "1"):
    env = dict(os.environ)
    if slow: env["S5B_VIEW_SLOW_PATH"] = "1"
    subprocess.check_call([CLI, "view", "-K", "20000", raw, "-c", rec_m, "-s", "ex-zd", "-o", z], stderr=subprocess.DEVNULL)
    subprocess.check_call([CLI, "view", "-K", "20000", z, "-c", "none", "-s", "none", "-o", back], stderr=subprocess.DEVNULL, env=env)
    a, b = open(raw, "rb").read(), open(back, "rb").read()
    print(rec_m, "slow" if slow else "fast", "identical", a == b, len(a), len(b))
    if a != b:
        pos = 68 + struct.unpack_from("<I", a, 64)[0]
        bad = []
        pa = pb = pos
        r = 0
        while a[pa:pa+5] != b"5WOLB" and r < R:
            sa = struct.unpack_from("<Q", a, pa)[0]; sb = struct.unpack_from("<Q", b, pb)[0]
            ra, rb = a[pa+8:pa+8+sa], b[pb+8:pb+8+sb]
            if ra != rb:
                xa = np.frombuffer(ra[-N*2:], np.int16); xb = np.frombuffer(rb[-N*2:], np.int16) if sb == sa else None
                first = int(np.nonzero(xa != xb)[0][0]) if xb is not None else -1
                bad.append((r, sa, sb, first))
            pa += 8 + sa; pb += 8 + sb; r += 1
        print("bad records", len(bad), bad[:20])
        if bad:
            r0 = bad[0][0]
            xa = sig[r0*N:(r0+1)*N]
            print("first bad read", r0, "first diff at", bad[0][3], xa[max(0,bad[0][3]-4):bad[0][3]+6])
