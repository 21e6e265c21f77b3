"""dev: decode pass of a ragged batch (BASELINE configs[3] lengths: lognormal, clipped) with the device-resident transcoder --
stage times with / without the length-sorted inflate order and for several thread-kernel thresholds (set by the environment)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import slow5tools_b200 as s5
from slow5tools_b200 import synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
lens = synth.lognormal_lengths(R, seed=5).astype(np.int64)
cdc = s5.Codec(0)
total = int(lens.sum())
sig = synth.nanopore_signal(total, seed=9, device="cuda")
# uncompressed records, variable length: 82 bytes of fixed fields + samples
head = synth.blow5_records(sig[:R * 8], R, 8, seed=3)[:, :synth.REC_HEAD].contiguous()
rl = synth.REC_HEAD + 2 * lens
off = np.zeros(R + 1, np.int64); off[1:] = np.cumsum(rl)
raw = torch.empty(int(off[-1]) + 64, dtype=torch.uint8, device="cuda")
sb = sig.view(torch.uint8)
soff = np.zeros(R + 1, np.int64); soff[1:] = np.cumsum(2 * lens)
for r in range(R):
    o = int(off[r])
    h = head[r].clone()
    h[-8:] = torch.tensor(list(int(lens[r]).to_bytes(8, "little")), dtype=torch.uint8, device="cuda")
    raw[o:o + synth.REC_HEAD] = h
    raw[o + synth.REC_HEAD:o + int(rl[r])] = sb[int(soff[r]):int(soff[r + 1])]
M_NONE, M_ZLIB, M_SVB = 0, 1, 2
enc = torch.zeros(int(off[-1]) * 3 // 4 + R * 600, dtype=torch.uint8, device="cuda")
eoff = torch.zeros(R + 1, dtype=torch.int64, device="cuda")
res = torch.zeros(2, dtype=torch.int64, device="cuda")
cdc.blow5_recode_dev(M_NONE, M_NONE, M_ZLIB, M_SVB, raw, int(off[-1]), off[:-1].astype(np.uint64), rl.astype(np.uint32), enc, res, eoff)
cdc.sync()
r_ = res.cpu().numpy(); assert int(r_[1]) == 0, int(r_[1])
eo = eoff.cpu().numpy().view(np.uint64)
zo, zl = eo[:-1] + np.uint64(8), (eo[1:] - eo[:-1] - np.uint64(8)).astype(np.uint32)
back = torch.zeros(int(off[-1]) + 8 * R + 64, dtype=torch.uint8, device="cuda")
res2 = torch.zeros(2, dtype=torch.int64, device="cuda")
for _ in range(2):
    cdc.blow5_recode_dev(M_ZLIB, M_SVB, M_NONE, M_NONE, enc, int(r_[0]), zo, zl, back, res2, None)
cdc.sync()
cdc.stage_timing(True); cdc.stage_report(reset=True)
K = 3
for _ in range(K):
    cdc.blow5_recode_dev(M_ZLIB, M_SVB, M_NONE, M_NONE, enc, int(r_[0]), zo, zl, back, res2, None)
cdc.sync()
st = {k: round(v[0] / K, 2) for k, v in cdc.stage_report(reset=True).items() if v[0] > 0}
r2 = res2.cpu().numpy(); assert int(r2[1]) == 0
ok = all(torch.equal(back[int(off[r]) + 8 * (r + 1):int(off[r]) + 8 * (r + 1) + int(rl[r])], raw[int(off[r]):int(off[r + 1])]) for r in range(0, R, max(1, R // 200)))
print(json.dumps({"reads": R, "raw_GB": round(total * 2 / 1e9, 2), "zlib_mean_bytes": int(zl.mean()), "order": os.environ.get("S5B_INFLATE_ORDER", "1"),
                  "thread_max": os.environ.get("S5B_INFLATE_THREAD_MAX", "16384"), "roundtrip_ok": ok, "decode_stage_ms": st}))
