"""dev: svbzd_decode_kernel timed by CUDA events on a stand-alone batch (compare with ncu's gpu__time_duration of the same launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import slow5tools_b200 as s5
from slow5tools_b200 import synth

R, N = int(sys.argv[1]) if len(sys.argv) > 1 else 250000, 4096
cdc = s5.Codec(0)
sig = synth.nanopore_signal(R * N, seed=42, device="cuda")
sig_off = torch.arange(R + 1, dtype=torch.int64, device="cuda") * N
ns = torch.full((R,), N, dtype=torch.int32, device="cuda")
slot = ((4 + N // 4 + 3 * N + 15) // 16) * 16
svb = torch.zeros(R * slot + 64, dtype=torch.uint8, device="cuda")
svb_off = torch.arange(R + 1, dtype=torch.int64, device="cuda") * slot
svb_len = torch.zeros(R, dtype=torch.int32, device="cuda")
st = torch.zeros(R, dtype=torch.int32, device="cuda")
cdc.svbzd_encode_dev(sig, sig_off, ns, svb, svb_off, svb_len, st)
torch.cuda.synchronize()
assert int(st.abs().sum()) == 0
back = torch.zeros_like(sig)
ns2 = torch.zeros_like(ns)
stream = cdc.recode_stream() if hasattr(cdc, "recode_stream") else None
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rep in range(3):
    times = []
    for it in range(10):
        torch.cuda.synchronize()
        ev[0].record()
        cdc.svbzd_decode_dev(svb, svb_off, svb_len, back, sig_off, ns2, st)
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]))
    print("decode ms per launch (events):", " ".join("%.3f" % t for t in times))
assert torch.equal(back, sig)
bytes_ = R * N * 2 + int(svb_len.sum())
print("bytes", bytes_, "best", min(times), "GB/s", bytes_ / min(times) / 1e6)
