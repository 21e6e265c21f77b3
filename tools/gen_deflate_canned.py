#!/usr/bin/env python
"""Generates slow5tools_b200/csrc/deflate_canned.inc: the fixed ("canned") dynamic-Huffman code the deflate encoder uses for the
part of a BLOW5 record in front of the svb-zd data bytes (record header fields, svb-zd sample count, control bytes).

That part looks the same in every record of every file: a few dozen ASCII / little-endian header bytes and ~N/4 control bytes
that are almost all zero.  Building a Huffman code per record for it (histogram, sort, tree, canonical codes, run-length coded
header) cost about as much as coding the 4x larger data part; a code fixed ahead of time costs nothing at run time and loses
next to nothing in size.  The block is still a DYNAMIC block (RFC 1951 3.2.7) -- its header is simply the same bits every
time -- so any inflate reads it, and a block that would not shrink is still written stored.

Frequencies: records of the SURVEY 8d signal model at three noise levels (1 %, 2.5 % and 8 % two-byte deltas) plus the real
reads of tests/golden/fixtures, tokenised exactly like the kernels do (literals + distance-1 run matches of 3..32 bytes inside
32-byte strips), with a floor so that every literal and every match length that the tokeniser can produce has a code.

    python tools/gen_deflate_canned.py          (rewrites the .inc and checks it with zlib)
"""
import heapq
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


def len_code(m):
    if m <= 10:
        return m - 3, 0, 0
    if m <= 18:
        return 8 + ((m - 11) >> 1), 1, (m - 11) & 1
    return 12 + ((m - 19) >> 2), 2, (m - 19) & 3


def tokens(block, prev0):
    """[(kind, value)]: kind 1 literal byte, kind 2 match length (distance 1); strips of 32 bytes like strip_token()"""
    out = []
    for t0 in range(0, len(block), 32):
        s = block[t0:t0 + 32]
        prev = [prev0] + list(s[:-1])
        eq = [int(b) == int(p) for b, p in zip(s, prev)]
        i = 0
        while i < len(s):
            if eq[i]:
                j = i
                while j + 1 < len(s) and eq[j + 1]:
                    j += 1
                m = j - i + 1
                if m >= 3:
                    out.append((2, m))
                else:
                    out.extend((1, int(b)) for b in s[i:j + 1])
                i = j + 1
            else:
                out.append((1, int(s[i])))
                i += 1
        prev0 = int(s[-1])
    return out


def svb_keys(sig):
    d = np.diff(np.concatenate([[0], sig.astype(np.int64)]))
    z = (d << 1) ^ (d >> 63)
    c = (z > 0xFF).astype(np.uint8) + (z > 0xFFFF).astype(np.uint8)
    c = np.concatenate([c, np.zeros((-len(c)) % 4, np.uint8)]).reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def synthetic_fronts(n_reads, noise, seed):
    """the bytes in front of the svb-zd data of synthetic records: fixed fields, len_raw_signal, u32 N, control bytes"""
    from slow5tools_b200 import synth
    import torch
    N = 4096
    sig = synth.nanopore_signal(n_reads * N, seed=seed, noise_sigma=noise).numpy().reshape(n_reads, N)
    rec = synth.blow5_records(torch.from_numpy(sig.reshape(-1)), n_reads, N, seed=seed).numpy()
    out = []
    for r in range(n_reads):
        keys = svb_keys(sig[r])
        head = bytearray(rec[r, :synth.REC_HEAD].tobytes())
        head[-8:] = struct.pack("<Q", 4 + len(keys) + N + int(np.unpackbits(keys).sum()))   # stored length: about right
        out.append(bytes(head) + struct.pack("<I", N) + keys.tobytes())
    return out


def real_fronts():
    """the same part of the real reads in the fixtures (uncompressed BLOW5: parse, svb-zd control bytes from the samples)"""
    out = []
    p = os.path.join(ROOT, "tests", "golden", "fixtures", "exp_1_lossless.blow5")
    if not os.path.exists(p):
        return out
    b = open(p, "rb").read()
    hsize = struct.unpack_from("<I", b, 64)[0]
    at = 68 + hsize
    while at + 8 <= len(b) and b[at:at + 5] != b"5WOLB":
        size = struct.unpack_from("<Q", b, at)[0]
        rec = b[at + 8:at + 8 + size]
        at += 8 + size
        idl = struct.unpack_from("<H", rec, 0)[0]
        o = 2 + idl + 4 + 32
        n = struct.unpack_from("<Q", rec, o)[0]
        sig = np.frombuffer(rec, "<i2", n, o + 8)
        keys = svb_keys(sig)
        head = bytearray(rec[:o + 8])
        out.append(bytes(head) + struct.pack("<I", n) + keys.tobytes())
    return out


def huffman_lengths(freq, limit):
    """code lengths of the symbols with freq > 0 (length-limited the way zlib's gen_bitlen repairs an over-long code)"""
    syms = [s for s, f in enumerate(freq) if f > 0]
    heap = [(freq[s], i, (s,)) for i, s in enumerate(syms)]
    heapq.heapify(heap)
    depth = {s: 0 for s in syms}
    tick = len(heap)
    while len(heap) > 1:
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        for s in a[2] + b[2]:
            depth[s] += 1
        heapq.heappush(heap, (a[0] + b[0], tick, a[2] + b[2]))
        tick += 1
    if max(depth.values()) > limit:
        cnt = [0] * (limit + 1)
        for s in syms:
            cnt[min(depth[s], limit)] += 1
        excess = sum(c << (limit - l) for l, c in enumerate(cnt) if l) - (1 << limit)
        while excess > 0:
            bits = limit - 1
            while cnt[bits] == 0:
                bits -= 1
            cnt[bits] -= 1
            cnt[bits + 1] += 2
            cnt[limit] -= 1
            excess -= 1
        order = sorted(syms, key=lambda s: (freq[s], -s))   # rarest first get the longest codes
        i = 0
        for bits in range(limit, 0, -1):
            for _ in range(cnt[bits]):
                depth[order[i]] = bits
                i += 1
    lens = [0] * len(freq)
    for s in syms:
        lens[s] = depth[s]
    return lens


def canonical(lens):
    """canonical codes, bit-reversed (ready to be written LSB first)"""
    maxl = max(lens)
    cnt = [0] * (maxl + 2)
    for l in lens:
        if l:
            cnt[l] += 1
    nxt, code = [0] * (maxl + 2), 0
    for l in range(1, maxl + 1):
        code = (code + cnt[l - 1]) << 1
        nxt[l] = code
    out = [0] * len(lens)
    for s, l in enumerate(lens):
        if l:
            c = nxt[l]
            nxt[l] += 1
            out[s] = int(format(c, "0%db" % l)[::-1], 2)
    return out


class Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, bits, nbits):
        self.v |= bits << self.n
        self.n += nbits

    def bytes(self):
        return self.v.to_bytes((self.n + 7) // 8, "little")


def header_bits(lens, hlit):
    """HLIT, HDIST, HCLEN, the code-length code and the run-length coded lengths (everything after BFINAL/BTYPE)"""
    seq = lens[:hlit] + [1]          # one distance code (distance 1) of one bit
    sym = []                         # (symbol, extra value)
    i = 0
    while i < len(seq):
        v = seq[i]
        j = i
        while j + 1 < len(seq) and seq[j + 1] == v:
            j += 1
        run = j - i + 1
        if v == 0:
            while run >= 11:
                k = min(run, 138)
                sym.append((18, k - 11))
                run -= k
            if run >= 3:
                sym.append((17, run - 3))
                run = 0
            sym.extend([(0, 0)] * run)
        else:
            sym.append((v, 0))
            run -= 1
            while run >= 3:
                k = min(run, 6)
                sym.append((16, k - 3))
                run -= k
            sym.extend([(v, 0)] * run)
        i = j + 1
    clfreq = [0] * 19
    for s, _ in sym:
        clfreq[s] += 1
    cllen = huffman_lengths(clfreq, 7)
    clcode = canonical(cllen)
    hclen = 19
    while hclen > 4 and cllen[CL_ORDER[hclen - 1]] == 0:
        hclen -= 1
    b = Bits()
    b.put(hlit - 257, 5)
    b.put(0, 5)
    b.put(hclen - 4, 4)
    for q in range(hclen):
        b.put(cllen[CL_ORDER[q]], 3)
    for s, x in sym:
        b.put(clcode[s], cllen[s])
        if s == 16:
            b.put(x, 2)
        elif s == 17:
            b.put(x, 3)
        elif s == 18:
            b.put(x, 7)
    return b


def main():
    fronts = real_fronts() * 20
    for noise, seed in ((7.0, 11), (9.0, 12), (14.0, 13)):
        fronts += synthetic_fronts(60, noise, seed)
    freq = [0.0] * 286
    for f in fronts:
        for kind, v in tokens(np.frombuffer(f, np.uint8), 0x200):
            freq[v if kind == 1 else 257 + len_code(v)[0]] += 1
        freq[256] += 1
    total = sum(freq)
    hlit = 257 + len_code(32)[0] + 1
    floor = total / 3000.0
    for s in range(hlit):
        freq[s] = freq[s] + floor if s != 256 else freq[s]
    lens = huffman_lengths([int(round(f * 16)) for f in freq[:hlit]] + [0] * (286 - hlit), 15) + [0, 0]
    assert max(lens) <= 15 and all(lens[s] for s in range(hlit))
    assert sum(1 << (15 - l) for l in lens if l) == 1 << 15, "the code must be complete"
    codes = canonical(lens)
    hdr = header_bits(lens, hlit)

    # ---- check with zlib: header + tokens of every front, as one final block of a raw deflate stream
    for f in fronts[::7] + [bytes(range(256)) * 3, b"", b"\0" * 1000]:
        b = Bits()
        b.put(1 | (2 << 1), 3)
        b.put(hdr.v, hdr.n)
        for kind, v in tokens(np.frombuffer(f, np.uint8), 0x200):
            if kind == 1:
                b.put(codes[v], lens[v])
            else:
                s, xb, xv = len_code(v)
                b.put(codes[257 + s], lens[257 + s])
                b.put(xv, xb)
                b.put(0, 1)
        b.put(codes[256], lens[256])
        assert zlib.decompress(b.bytes(), -15) == f
    sizes = []
    for f in fronts:
        bits = 3 + hdr.n + lens[256]
        for kind, v in tokens(np.frombuffer(f, np.uint8), 0x200):
            bits += lens[v] if kind == 1 else lens[257 + len_code(v)[0]] + len_code(v)[1] + 1
        # the same tokens under a code built for this block alone (what the encoder did before)
        own = [0] * 286
        for kind, v in tokens(np.frombuffer(f, np.uint8), 0x200):
            own[v if kind == 1 else 257 + len_code(v)[0]] += 1
        own[256] = 1
        ol = huffman_lengths(own, 15)
        oh = 286
        while oh > 257 and ol[oh - 1] == 0:
            oh -= 1
        obits = 3 + header_bits(ol, oh).n + sum(own[q] * ol[q] for q in range(286))
        obits += sum(own[257 + q] * (0 if q < 8 else 1 if q < 12 else 2) + own[257 + q] for q in range(16))
        sizes.append((bits / 8, len(zlib.compress(f, 6)) - 6, len(f), obits / 8))
    a = np.array(sizes)
    print("front part: %.0f bytes on average -> canned %.1f bytes, own code per block %.1f bytes, zlib -6 %.1f bytes; "
          "header %d bits, longest code %d" % (a[:, 2].mean(), a[:, 0].mean(), a[:, 3].mean(), a[:, 1].mean(), hdr.n, max(lens)))
    syn = a[-60:]
    print("  (last synthetic set alone: %.0f -> canned %.1f, own %.1f, zlib %.1f)" % (syn[:, 2].mean(), syn[:, 0].mean(), syn[:, 3].mean(), syn[:, 1].mean()))

    words = list(struct.unpack("<%dI" % ((hdr.n + 31) // 32), hdr.bytes().ljust(((hdr.n + 31) // 32) * 4, b"\0")))
    tab = [codes[s] | (lens[s] << 16) for s in range(288)]
    with open(os.path.join(ROOT, "slow5tools_b200", "csrc", "deflate_canned.inc"), "w") as f:
        f.write("// generated by tools/gen_deflate_canned.py -- do not edit (included inside namespace s5b::defl)\n")
        f.write("// code (bit-reversed, ready to be written LSB first) | length << 16 of literal/length symbol s\n")
        f.write("__device__ const uint32_t g_canned_tab[288] = {\n")
        for i in range(0, 288, 8):
            f.write("    " + ", ".join("0x%05x" % t for t in tab[i:i + 8]) + ",\n")
        f.write("};\n// HLIT, HDIST, HCLEN, code-length code, run-length coded lengths: the bits after BFINAL / BTYPE\n")
        f.write("constexpr uint32_t CANNED_HDR_BITS = %d;\nconstexpr int CANNED_HDR_WORDS = %d;\n" % (hdr.n, len(words)))
        f.write("constexpr int CANNED_MAX_LEN = %d;  // longest code (the emit kernel sizes its room for a block from it)\n" % max(lens))
        f.write("__device__ const uint32_t g_canned_hdr[CANNED_HDR_WORDS] = {\n")
        for i in range(0, len(words), 6):
            f.write("    " + ", ".join("0x%08x" % w for w in words[i:i + 6]) + ",\n")
        f.write("};\n")
    # ---- the decoder's side: inflate_thread_kernel recognises the header and loads these instead of building them
    # (mirrors build_code() in inflate_thread_kernels.cu)
    cnt = [0] * 16
    for l in lens[:hlit]:
        cnt[l] += l > 0
    base, lim, start = [0] * 16, [0] * 16, [0] * 16
    code = at = 0
    for l in range(1, 16):
        c = cnt[l]
        base[l] = at - code
        lim[l] = min((code + c) << (15 - l), 0xffff)
        start[l] = at
        at += c
        code = (code + c) << 1
    nxt = list(start)
    srt = [0] * 288
    hi_start = list(start)
    for sym in range(hlit):
        l = lens[sym]
        if l:
            srt[nxt[l]] = sym
            nxt[l] += 1
            if sym < 256:
                hi_start[l] = nxt[l]
    for l in range(1, 16):
        if cnt[l] == 0:
            hi_start[l] = start[l]
    with open(os.path.join(ROOT, "slow5tools_b200", "csrc", "inflate_canned.inc"), "w") as f:
        f.write("// generated by tools/gen_deflate_canned.py -- do not edit (included inside namespace s5b::inft)\n")
        f.write("// the header bits of the encoder's canned block (after BFINAL / BTYPE) and the decode tables build_code() would\n")
        f.write("// make from them: limit << 4 | length per code length, (base & 0xffff) | first non-literal index << 16 per code\n")
        f.write("// length, the symbols sorted by (length, symbol), their low bytes packed four to a word\n")
        f.write("constexpr uint32_t CANNED_HDR_BITS = %d;\nconstexpr int CANNED_HDR_WORDS = %d;\n" % (hdr.n, len(words)))
        f.write("__device__ const uint32_t g_canned_hdr[CANNED_HDR_WORDS] = {\n")
        for i in range(0, len(words), 6):
            f.write("    " + ", ".join("0x%08x" % w for w in words[i:i + 6]) + ",\n")
        f.write("};\n__device__ const uint32_t g_canned_L[16] = {" + ", ".join("0x%05x" % ((lim[l] << 4) | l) for l in range(16)) + "};\n")
        f.write("__device__ const uint32_t g_canned_lbn[16] = {" + ", ".join("0x%08x" % ((base[l] & 0xffff) | (hi_start[l] << 16)) for l in range(16)) + "};\n")
        f.write("__device__ const uint16_t g_canned_sorted[288] = {\n")
        for i in range(0, 288, 16):
            f.write("    " + ", ".join("%d" % v for v in srt[i:i + 16]) + ",\n")
        f.write("};\n__device__ const uint32_t g_canned_sorted8[72] = {\n")
        w8 = [(srt[4 * i] & 255) | ((srt[4 * i + 1] & 255) << 8) | ((srt[4 * i + 2] & 255) << 16) | ((srt[4 * i + 3] & 255) << 24) for i in range(72)]
        for i in range(0, 72, 8):
            f.write("    " + ", ".join("0x%08x" % v for v in w8[i:i + 8]) + ",\n")
        f.write("};\n")


if __name__ == "__main__":
    main()
