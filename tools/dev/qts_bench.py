"""Times qts_round_kernel (slow5tools degrade's per-sample step) on a slab of BASELINE configs[2]'s size and prints its HBM
fraction: algorithmic bytes = 2 B read + 2 B written per sample.  usage: python tools/dev/qts_bench.py [reads] [samples]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from slow5tools_b200.codec import Codec  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n = reads * samples
cd = Codec(0)
x = torch.randint(0, 2048, (n,), dtype=torch.int16, device="cuda")
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6451.8)
for _ in range(3):
    cd.qts_round_dev(x, 3)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
a.record()
for _ in range(K):
    cd.qts_round_dev(x, 3)  # 8 GB slab: larger than L2, every launch streams from HBM
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
gbs = 4.0 * n / ms / 1e6
print(json.dumps({"kernel": "qts_round_kernel", "reads": reads, "samples_per_read": samples, "ms": ms, "achieved_gbs": gbs,
                  "peak_gbs": peak, "frac": gbs / peak, "bytes_per_sample": 4}))
