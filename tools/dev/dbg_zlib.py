import sys, zlib, numpy as np, torch
sys.path.insert(0, '.')
import slow5tools_b200 as s5
from slow5tools_b200 import synth
R, N = 100000, 4096
cdc = s5.Codec(0)
sig = synth.nanopore_signal(R * N, seed=42, device="cuda")
n = torch.full((R,), N, dtype=torch.int32, device="cuda")
soff = torch.arange(R + 1, dtype=torch.int64, device="cuda") * N
slot = int(s5.lib.s5b_svbzd_slot(N))
ooff = torch.arange(R + 1, dtype=torch.int64, device="cuda") * slot
svb = torch.zeros(R * slot + 16, dtype=torch.uint8, device="cuda")
svb_len = torch.zeros(R, dtype=torch.int32, device="cuda")
st = torch.ones(R, dtype=torch.int32, device="cuda")
cdc.svbzd_encode_dev(sig, soff, n, svb, ooff, svb_len, st)
zslot = int(s5.lib.s5b_zlib_bound(slot))
zoff = torch.arange(R + 1, dtype=torch.int64, device="cuda") * zslot
zbuf = torch.zeros(R * zslot + 16, dtype=torch.uint8, device="cuda")
zlen = torch.zeros(R, dtype=torch.int32, device="cuda")
zst = torch.ones(R, dtype=torch.int32, device="cuda")
split = torch.full((R,), 4 + (N + 3) // 4, dtype=torch.int32, device="cuda")
cdc.zlib_deflate_dev(svb, ooff, svb_len, zbuf, zoff, zlen, zst, split=split)
torch.cuda.synchronize()
bad = torch.nonzero(zst).flatten().cpu().numpy()
print("deflate bad:", len(bad), bad[:10], zst[bad[:10]].cpu().numpy() if len(bad) else "")
svb2 = torch.zeros_like(svb); svb2_len = torch.zeros_like(svb_len); ist = torch.ones(R, dtype=torch.int32, device="cuda")
cdc.zlib_inflate_dev(zbuf, zoff, zlen, svb2, ooff, svb2_len, ist)
torch.cuda.synchronize()
bad = torch.nonzero(ist).flatten().cpu().numpy()
print("inflate bad:", len(bad), bad[:10], ist[bad[:10]].cpu().numpy() if len(bad) else "")
neq = torch.nonzero(svb2_len != svb_len).flatten().cpu().numpy()
print("len mismatch:", len(neq), neq[:10])
zl = zlen.cpu().numpy(); sl = svb_len.cpu().numpy()
chk = list(bad[:5]) + list(neq[:5]) + [0, 1, 99999]
for i in chk:
    i = int(i)
    z = zbuf[i * zslot:i * zslot + int(zl[i])].cpu().numpy().tobytes()
    raw = svb[i * slot:i * slot + int(sl[i])].cpu().numpy().tobytes()
    try:
        d = zlib.decompress(z)
        print(i, "host zlib ok:", d == raw, len(z), len(raw))
    except Exception as e:
        print(i, "host zlib error", e, len(z), len(raw))
