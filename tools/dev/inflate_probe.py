"""dev: which of the ragged test records fail to inflate on the device"""
import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import slow5tools_b200 as s5
from recode_helpers import make_records, walk_image, M_NONE, M_ZLIB, M_SVB_ZD
from test_recode_gpu import host_recode, LENS

cdc = s5.Codec(0)
recs, _ = make_records(LENS, seed=13)
rc, img, off = host_recode(cdc, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs, gap=8)
assert rc == 0
stored = walk_image(img)
print("streams", len(stored))
for mis in (0, 1, 3):
    n = len(stored)
    lens = np.array([len(z) for z in stored], np.uint32)
    ioff = np.zeros(n + 1, np.uint64)
    pos = mis
    for i, z in enumerate(stored):
        ioff[i] = pos
        pos += len(z) + mis
    ioff[-1] = pos
    din = np.zeros((pos + 64 + 15) // 16 * 16, np.uint8)
    for z, o in zip(stored, ioff):
        din[int(o):int(o) + len(z)] = np.frombuffer(z, np.uint8)
    raw = [zlib.decompress(z) for z in stored]
    ooff = np.zeros(n + 1, np.uint64)
    ooff[1:] = np.cumsum([len(r) + 64 for r in raw])
    out = torch.zeros(int(ooff[-1]) + 64, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    st = torch.full((n,), 99, dtype=torch.int32, device="cuda")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cdc.zlib_inflate_dev(d(din), d(ioff.view(np.int64)), d(lens.view(np.int32)), out, d(ooff.view(np.int64)), out_len, st)
    torch.cuda.synchronize()
    sth, lh, oh = st.cpu().numpy(), out_len.cpu().numpy(), out.cpu().numpy()
    bad = [i for i in range(n) if sth[i] != 0 or oh[int(ooff[i]):int(ooff[i]) + lh[i]].tobytes() != raw[i]]
    print("misalign", mis, "bad", [(i, LENS[i], int(sth[i]), len(stored[i])) for i in bad])
