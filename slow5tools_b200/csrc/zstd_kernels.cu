// zstd_kernels.cu -- sm_100a Zstandard frame decoder, one warp per record frame.
//
// Replaces ptr_depress_zstd (slow5lib/src/slow5_press.c:1205-1230) for whole batches: every record is one
// independent frame that must carry its content size (:1206-1211), so output slots are exact.  The serial
// pieces (frame / block / table parsing, FSE state machines) are the host+device functions of zstd_core.h
// executed by lane 0 (or by lanes 0..3 for the four Huffman literal streams, which the format made
// independent precisely so they can be decoded in parallel); the warp cooperates on staging the frame in
// shared memory and on every bulk copy (raw / RLE blocks, literal runs, matches).
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "zstd_core.h"
#include "../../include/slow5b200.h"

namespace s5b {

namespace {

using namespace s5bz;

constexpr int ZD_WARPS = 4;
constexpr int ZD_IN_STAGE = 6144;   // frames up to this size are decoded out of shared memory
constexpr int ZD_LIT_STAGE = 6144;  // literal buffers up to this size live in shared memory

struct __align__(16) ZdWarpSmem {
    Tables t;
    uint8_t in[ZD_IN_STAGE];
    uint8_t lit[ZD_LIT_STAGE];
};

__device__ __forceinline__ void warp_copy_bytes(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}
// overlap-safe match copy: dst[k] = dst[k - offset] for k in [0, n), sources older than this call only
__device__ __forceinline__ void warp_match_copy(uint8_t *dst, uint64_t offset, uint32_t n, int lane) {
    if (offset >= n) {
        for (uint32_t i = lane; i < n; i += 32) dst[i] = *(dst + i - offset);
    } else {
        for (uint32_t i = lane; i < n; i += 32) dst[i] = *(dst - offset + (i % offset));
    }
}

// ---- window decode of the four Huffman literal streams (all 32 lanes, 8 per stream) ------------------------
// A zstd Huffman stream is read from its last bit towards its first: with x = number of unread bits, the next
// symbol is huf[bits [x - mb, x)] (bits below the stream start read as zero) and x drops by that symbol's length;
// a valid stream ends at exactly x == 0 after `regen` symbols (huf_decode_stream in zstd_core.h).  Each stream is
// cut into 8 chunks of x; a lane decodes the symbols that start inside its chunk, first speculatively from the
// chunk's top bit -- Huffman codes self-synchronise, so the run soon falls onto true symbol boundaries -- and again
// from its upper neighbour's end position until all boundaries agree (usually one extra round); an 8-lane prefix
// scan over the symbol counts then places every lane's output and a last pass writes it.
struct HufLane {
    const uint32_t *w4;  // 4-byte aligned address at or below the stream's first byte (shared or global memory)
    uint32_t delta;      // grid bit position of stream bit 0 (0, 8, 16 or 24)
    uint64_t acc;        // unread bits, left-aligned
    int cnt;             // valid bits in acc
    int j;               // next (lower) grid word to load
    __device__ __forceinline__ uint32_t word(int k) const {
        if (k < 0) return 0u;
        uint32_t w = w4[k];
        if (k == 0) w &= ~((1u << delta) - 1u);  // bytes below the stream start are not part of it
        return w;
    }
    // position the reader so that stream bit x - 1 is the next bit (x >= 1)
    __device__ __forceinline__ void seek(int x) {
        const int g = x + (int)delta;
        j = (g - 1) >> 5;
        const int c0 = g - 32 * j;  // 1..32 bits of word j lie below g
        acc = (uint64_t)word(j) << (64 - c0);
        cnt = c0;
        --j;
        refill();
    }
    __device__ __forceinline__ void refill() {
        if (cnt <= 32) {
            acc |= (uint64_t)word(j) << (32 - cnt);
            cnt += 32;
            --j;
        }
    }
};

// Symbols that start at x in (limit, start]; returns the end position (<= limit, may be negative for a damaged
// stream) and the symbol count.  WRITE: symbols go to out[0 ..).
template <bool WRITE>
__device__ __forceinline__ void huf_run(const uint16_t *huf, const int mb, HufLane &b, const int start, const int limit,
                                        int &end, uint32_t &count, uint8_t *out) {
    int x = start;
    uint32_t n = 0;
    if (x > limit) b.seek(x);
    while (x > limit) {
        b.refill();
        const uint32_t e = huf[(uint32_t)(b.acc >> (64 - mb))];
        const int nb = (int)(e & 15u);
        if (WRITE) out[n] = (uint8_t)(e >> 4);
        ++n;
        b.acc <<= nb;
        b.cnt -= nb;
        x -= nb;
    }
    end = x;
    count = n;
}

// Returns Z_OK, or Z_ERR_CORRUPT when the streams do not end exactly at their first bit after exactly the
// regenerated sizes (the caller then lets the serial decoders give the verdict).
__device__ int huf_decode_4x_window(const Tables &t, const uint8_t *const sp[4], const uint32_t sl[4], uint8_t *lit,
                                    const uint32_t q, const uint32_t regen, const int lane) {
    const int g = lane >> 3, k = lane & 7;  // stream, chunk
    const uint8_t *src = sp[0];
    uint32_t len = sl[0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (g == i) {
            src = sp[i];
            len = sl[i];
        }
    const uint32_t want = g < 3 ? q : regen - 3 * q;
    bool ok = len != 0 && len < (1u << 24);
    const uint32_t lastb = ok ? src[len - 1] : 1u;
    ok = ok && lastb != 0;
    const int T = ok ? (int)(len * 8 - (8 - (uint32_t)highest_bit(lastb))) : 0;  // data bits below the end mark
    HufLane b;
    const uint32_t sk = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    b.w4 = reinterpret_cast<const uint32_t *>(src - sk);
    b.delta = sk * 8u;
    const int mb = t.huf_bits;
    const int C = max(32, (T + 7) >> 3);
    int start = max(T - k * C, 0);
    const int limit = k == 7 ? 0 : max(T - (k + 1) * C, 0);
    int end;
    uint32_t cnt;
    huf_run<false>(t.huf, mb, b, start, limit, end, cnt, nullptr);
    for (int round = 1; round <= 8; ++round) {
        const int pe = __shfl_up_sync(FULL, end, 1, 8);
        const bool need = k > 0 && pe != start;
        if (!__any_sync(FULL, need)) break;
        if (need) {
            start = pe;
            huf_run<false>(t.huf, mb, b, start, limit, end, cnt, nullptr);
        }
    }
    // agreement, exact end, exact symbol count -- per stream, then for the whole warp
    const int pe = __shfl_up_sync(FULL, end, 1, 8);
    bool good = ok && (k == 0 || pe == start) && (k < 7 || end == 0);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, d, 8);
        if (k >= d) incl += v;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 7, 8);
    good = good && total == want;
    if (!__all_sync(FULL, good)) return Z_ERR_CORRUPT;
    huf_run<true>(t.huf, mb, b, start, limit, end, cnt, lit + (uint32_t)g * q + (incl - cnt));
    __syncwarp();
    return Z_OK;
}

// warp version of s5bz::decode_literals: tree by lane 0, the four streams by the window decoder
__device__ int decode_literals_warp(Tables &t, const uint8_t *src, uint32_t len, uint8_t *lit, uint32_t lit_cap,
                                    const uint8_t **lit_ptr, uint32_t *lit_len, int lane) {
    if (len < 1) return Z_ERR_CORRUPT;
    const uint32_t b0 = src[0];
    const uint32_t type = b0 & 3, sf = (b0 >> 2) & 3;
    uint32_t regen, comp = 0, hdr, streams = 1;
    if (type < 2) {
        if ((sf & 1) == 0) {
            regen = b0 >> 3;
            hdr = 1;
        } else if (sf == 1) {
            if (len < 2) return Z_ERR_CORRUPT;
            regen = (b0 >> 4) | ((uint32_t)src[1] << 4);
            hdr = 2;
        } else {
            if (len < 3) return Z_ERR_CORRUPT;
            regen = (b0 >> 4) | ((uint32_t)src[1] << 4) | ((uint32_t)src[2] << 12);
            hdr = 3;
        }
        if (type == 0) {
            if (hdr + regen > len) return Z_ERR_CORRUPT;
            *lit_ptr = src + hdr;
            *lit_len = regen;
            return (int)(hdr + regen);
        }
        if (hdr + 1 > len || regen > lit_cap) return Z_ERR_CORRUPT;
        const uint8_t v = src[hdr];
        for (uint32_t i = lane; i < regen; i += 32) lit[i] = v;
        __syncwarp();
        *lit_ptr = lit;
        *lit_len = regen;
        return (int)(hdr + 1);
    }
    if (sf == 0 || sf == 1) {
        if (len < 3) return Z_ERR_CORRUPT;
        const uint32_t v = b0 | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16);
        regen = (v >> 4) & 0x3FF;
        comp = (v >> 14) & 0x3FF;
        hdr = 3;
        streams = sf == 0 ? 1 : 4;
    } else if (sf == 2) {
        if (len < 4) return Z_ERR_CORRUPT;
        const uint32_t v = rd32(src);
        regen = (v >> 4) & 0x3FFF;
        comp = v >> 18;
        hdr = 4;
        streams = 4;
    } else {
        if (len < 5) return Z_ERR_CORRUPT;
        const uint64_t v = (uint64_t)rd32(src) | ((uint64_t)src[4] << 32);
        regen = (uint32_t)(v >> 4) & 0x3FFFF;
        comp = (uint32_t)(v >> 22) & 0x3FFFF;
        hdr = 5;
        streams = 4;
    }
    if (hdr + comp > len || regen > lit_cap) return Z_ERR_CORRUPT;
    const uint8_t *p = src + hdr;
    uint32_t left = comp;
    if (type == 2) {
        int used = 0;
        if (lane == 0) used = huf_read_tree(t, p, left);
        used = __shfl_sync(FULL, used, 0);
        __syncwarp();
        if (used < 0) return used;
        p += used;
        left -= used;
    } else if (t.huf_bits == 0) {
        return Z_ERR_CORRUPT;
    }
    int rc = Z_OK;
    if (streams == 1) {
        if (lane == 0) rc = huf_decode_stream(t, p, left, lit, regen);
    } else {
        if (left < 6) return Z_ERR_CORRUPT;
        const uint32_t s1 = p[0] | (p[1] << 8), s2 = p[2] | (p[3] << 8), s3 = p[4] | (p[5] << 8);
        if (6ull + s1 + s2 + s3 > left) return Z_ERR_CORRUPT;
        const uint32_t s4 = left - 6 - s1 - s2 - s3;
        const uint32_t q = (regen + 3) / 4;
        if (3ull * q > regen) return Z_ERR_CORRUPT;
        const uint8_t *b = p + 6;
        const uint8_t *const sp[4] = {b, b + s1, b + s1 + s2, b + s1 + s2 + s3};
        const uint32_t sl[4] = {s1, s2, s3, s4};
        if (huf_decode_4x_window(t, sp, sl, lit, q, regen, lane) != Z_OK) {
            // damaged (or pathological) streams: the serial decoders own the verdict
            if (lane == 0) rc = huf_decode_stream(t, b, s1, lit, q);
            else if (lane == 1) rc = huf_decode_stream(t, b + s1, s2, lit + q, q);
            else if (lane == 2) rc = huf_decode_stream(t, b + s1 + s2, s3, lit + 2 * q, q);
            else if (lane == 3) rc = huf_decode_stream(t, b + s1 + s2 + s3, s4, lit + 3 * q, regen - 3 * q);
        }
    }
    // any lane's failure fails the block
    rc = __any_sync(FULL, rc != Z_OK) ? Z_ERR_CORRUPT : Z_OK;
    __syncwarp();
    if (rc != Z_OK) return rc;
    *lit_ptr = lit;
    *lit_len = regen;
    return (int)(hdr + comp);
}

// warp version of s5bz::decode_sequences: lane 0 walks the FSE state machine, the warp copies
__device__ int decode_sequences_warp(Tables &t, FrameState &fs, const uint8_t *src, uint32_t len, const uint8_t *lit,
                                     uint32_t lit_len, uint8_t *dst, uint64_t dst_cap, uint64_t *dst_pos, int lane) {
    if (len < 1) return Z_ERR_CORRUPT;
    uint32_t pos = 0;
    uint32_t nseq = src[pos++];
    if (nseq >= 128) {
        if (nseq == 255) {
            if (len < 3) return Z_ERR_CORRUPT;
            nseq = src[1] + ((uint32_t)src[2] << 8) + 0x7F00;
            pos = 3;
        } else {
            if (len < 2) return Z_ERR_CORRUPT;
            nseq = ((nseq - 128) << 8) + src[1];
            pos = 2;
        }
    }
    uint64_t out = *dst_pos;
    uint32_t lp = 0;
    if (nseq) {
        int rc = Z_OK;
        BackBits bb;
        uint32_t sl = 0, so = 0, sm = 0;
        if (lane == 0) {
            if (pos >= len) {
                rc = Z_ERR_CORRUPT;
            } else {
                const uint32_t modes = src[pos++];
                for (int w = 0; w < 3 && rc == Z_OK; ++w) {
                    const int used = seq_table(t, w, (modes >> (6 - 2 * w)) & 3, src + pos, len - pos);
                    if (used < 0) rc = used;
                    else pos += used;
                }
                if (rc == Z_OK && (pos >= len || !bb.init(src + pos, len - pos))) rc = Z_ERR_CORRUPT;
                if (rc == Z_OK) {
                    sl = bb.read(t.ll_al);
                    so = bb.read(t.of_al);
                    sm = bb.read(t.ml_al);
                    if (bb.off < 0) rc = Z_ERR_CORRUPT;
                }
            }
        }
        rc = __shfl_sync(FULL, rc, 0);
        __syncwarp();
        if (rc != Z_OK) return rc;
        for (uint32_t i = 0; i < nseq; ++i) {
            uint32_t llen = 0, mlen = 0;
            uint64_t offset = 0;
            if (lane == 0) {
                const int of_code = t.of[so].sym, ml_code = t.ml[sm].sym, ll_code = t.ll[sl].sym;
                if (of_code > 31 || ml_code > 52 || ll_code > 35) {
                    rc = Z_ERR_CORRUPT;
                } else {
                    const uint64_t ofv = (1ull << of_code) + bb.read(of_code);
                    mlen = ml_base_of(ml_code) + bb.read(ml_bits_of(ml_code));
                    llen = ll_base_of(ll_code) + bb.read(ll_bits_of(ll_code));
                    if (i + 1 < nseq) {
                        sl = t.ll[sl].base + bb.read(t.ll[sl].nbits);
                        sm = t.ml[sm].base + bb.read(t.ml[sm].nbits);
                        so = t.of[so].base + bb.read(t.of[so].nbits);
                    }
                    if (bb.off < 0) rc = Z_ERR_CORRUPT;
                    if (ofv > 3) {
                        offset = ofv - 3;
                        fs.rep[2] = fs.rep[1];
                        fs.rep[1] = fs.rep[0];
                        fs.rep[0] = offset;
                    } else {
                        uint32_t idx = (uint32_t)ofv - 1;
                        if (llen == 0) idx++;
                        if (idx == 0) {
                            offset = fs.rep[0];
                        } else {
                            offset = idx < 3 ? fs.rep[idx] : fs.rep[0] - 1;
                            if (idx > 1) fs.rep[2] = fs.rep[1];
                            fs.rep[1] = fs.rep[0];
                            fs.rep[0] = offset;
                        }
                    }
                    if (i + 1 == nseq && bb.off != 0) rc = Z_ERR_CORRUPT;
                }
            }
            rc = __shfl_sync(FULL, rc, 0);
            if (rc != Z_OK) return rc;
            llen = __shfl_sync(FULL, llen, 0);
            mlen = __shfl_sync(FULL, mlen, 0);
            offset = __shfl_sync(FULL, offset, 0);
            if (llen > lit_len - lp) return Z_ERR_CORRUPT;
            if (out + llen + mlen > dst_cap) return Z_ERR_NOSPACE;
            warp_copy_bytes(dst + out, lit + lp, llen, lane);
            out += llen;
            lp += llen;
            if (offset == 0 || offset > out) return Z_ERR_CORRUPT;
            __syncwarp();
            warp_match_copy(dst + out, offset, mlen, lane);
            out += mlen;
            __syncwarp();
        }
    } else if (pos != len) {
        return Z_ERR_CORRUPT;
    }
    const uint32_t rest = lit_len - lp;
    if (out + rest > dst_cap) return Z_ERR_NOSPACE;
    warp_copy_bytes(dst + out, lit + lp, rest, lane);
    out += rest;
    __syncwarp();
    *dst_pos = out;
    return Z_OK;
}

}  // namespace

__global__ void __launch_bounds__(ZD_WARPS * 32) zstd_decode_kernel(const InflateArgs a, uint8_t *lit_scratch,
                                                                    uint32_t lit_scratch_per_warp) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    ZdWarpSmem &ws = reinterpret_cast<ZdWarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    uint8_t *glit = lit_scratch + (uint64_t)(blockIdx.x * ZD_WARPS + (threadIdx.x >> 5)) * lit_scratch_per_warp;

    for (;;) {
        unsigned long long r = 0;
        if (lane == 0) r = atomicAdd(a.work_counter, 1ULL);
        r = __shfl_sync(FULL, r, 0);
        if (r >= a.n_reads) break;
        const uint64_t ioff = a.in_off[r];
        const uint32_t ilen = a.in_len[r];
        const uint64_t ooff = a.out_off[r];
        const uint64_t ocap = a.out_off[r + 1] - ooff;
        int rc = Z_OK;
        uint64_t out = 0;
        if (ioff + ilen > a.in_capacity) {
            rc = S5B_ERR_ARG;
        } else {
            const uint8_t *src = a.in + ioff;
            if (ilen <= ZD_IN_STAGE) {  // stage the frame: the serial decoders then read shared memory
                warp_copy_bytes(ws.in, src, ilen, lane);
                __syncwarp();
                src = ws.in;
            }
            uint8_t *dst = a.out + ooff;
            FrameInfo fi;
            rc = parse_frame_header(src, ilen, fi);
            if (rc == Z_OK && !fi.has_content_size) rc = Z_ERR_CORRUPT;  // slow5_press.c:1206-1211
            if (rc == Z_OK && fi.content_size > ocap) {
                out = fi.content_size;
                rc = Z_ERR_NOSPACE;
            }
            if (rc == Z_OK) {
                if (lane == 0) {
                    ws.t.huf_bits = 0;
                    ws.t.ll_al = ws.t.of_al = ws.t.ml_al = -1;
                }
                __syncwarp();
                FrameState fs;
                fs.rep[0] = 1;
                fs.rep[1] = 4;
                fs.rep[2] = 8;
                uint64_t pos = fi.header_bytes;
                for (;;) {
                    if (pos + 3 > ilen) {
                        rc = Z_ERR_CORRUPT;
                        break;
                    }
                    const uint32_t bh = src[pos] | ((uint32_t)src[pos + 1] << 8) | ((uint32_t)src[pos + 2] << 16);
                    pos += 3;
                    const bool last = bh & 1;
                    const uint32_t type = (bh >> 1) & 3, bsize = bh >> 3;
                    if (type == 0) {
                        if (pos + bsize > ilen || out + bsize > fi.content_size) {
                            rc = Z_ERR_CORRUPT;
                            break;
                        }
                        warp_copy_bytes(dst + out, src + pos, bsize, lane);
                        out += bsize;
                        pos += bsize;
                    } else if (type == 1) {
                        if (pos + 1 > ilen || out + bsize > fi.content_size) {
                            rc = Z_ERR_CORRUPT;
                            break;
                        }
                        const uint8_t v = src[pos];
                        for (uint32_t k = lane; k < bsize; k += 32) dst[out + k] = v;
                        out += bsize;
                        pos += 1;
                    } else if (type == 2) {
                        if (pos + bsize > ilen || bsize > (128u << 10)) {
                            rc = Z_ERR_CORRUPT;
                            break;
                        }
                        // literal buffer: shared memory when the block's literals fit, else this warp's global scratch
                        const uint8_t *lp;
                        uint32_t ll;
                        // the regenerated size is not known before the header is parsed: pick by an upper bound
                        const bool small = fi.content_size <= ZD_LIT_STAGE;
                        const int used = decode_literals_warp(ws.t, src + pos, bsize, small ? ws.lit : glit,
                                                              small ? ZD_LIT_STAGE : lit_scratch_per_warp, &lp, &ll, lane);
                        if (used < 0) {
                            rc = used;
                            break;
                        }
                        // each lane keeps its own copy of the repeat offsets; only lane 0's advances
                        rc = decode_sequences_warp(ws.t, fs, src + pos + used, bsize - used, lp, ll, dst, fi.content_size, &out, lane);
                        if (rc != Z_OK) {
                            if (rc == Z_ERR_NOSPACE) rc = Z_ERR_CORRUPT;
                            break;
                        }
                        pos += bsize;
                    } else {
                        rc = Z_ERR_CORRUPT;
                        break;
                    }
                    __syncwarp();
                    if (last) break;
                }
                if (rc == Z_OK && out != fi.content_size) rc = Z_ERR_CORRUPT;
                if (rc == Z_OK && fi.checksum) {
                    if (pos + 4 > ilen) {
                        rc = Z_ERR_CORRUPT;
                    } else {
                        __threadfence_block();
                        int bad = 0;
                        if (lane == 0) bad = (uint32_t)xxh64(dst, out) != rd32(src + pos);
                        if (__shfl_sync(FULL, bad, 0)) rc = Z_ERR_CORRUPT;
                        pos += 4;
                    }
                }
                if (rc == Z_OK && pos != ilen) rc = Z_ERR_CORRUPT;
            }
        }
        __syncwarp();
        if (lane == 0) {
            a.status[r] = rc;
            a.out_len[r] = (rc == Z_OK || rc == Z_ERR_NOSPACE) ? (uint32_t)out : 0u;
        }
    }
}

int zstd_decode_blocks_per_sm() {
    int n = 0;
    if (cudaFuncSetAttribute(zstd_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(ZdWarpSmem) * ZD_WARPS)) != cudaSuccess)
        return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, zstd_decode_kernel, ZD_WARPS * 32,
                                                      sizeof(ZdWarpSmem) * ZD_WARPS) != cudaSuccess)
        return 0;
    return n;
}

// content size of every frame, read from its header on the device (the pipelined transcoder sizes its slots from these
// without a trip to the host)
__global__ void zstd_sizes_kernel(const uint8_t *in, const uint64_t *off, const uint32_t *len, uint64_t n, uint32_t mul,
                                  uint32_t add, uint32_t *size, int32_t *status) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    s5bz::FrameInfo fi;
    int32_t st = S5B_OK;
    uint32_t sz = 0;
    if (s5bz::parse_frame_header(in + off[r], len[r], fi) != s5bz::Z_OK || !fi.has_content_size) {
        st = S5B_ERR_PRESS;
    } else if (fi.content_size > (uint64_t)len[r] * mul + add) {
        st = S5B_ERR_NOSPACE;
    } else {
        sz = (uint32_t)fi.content_size;
    }
    size[r] = sz;
    status[r] = st;
}
cudaError_t launch_zstd_sizes(const uint8_t *in, const uint64_t *off, const uint32_t *len, uint64_t n, uint32_t mul,
                              uint32_t add, uint32_t *size, int32_t *status, cudaStream_t st) {
    if (!n) return cudaSuccess;
    zstd_sizes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, off, len, n, mul, add, size, status);
    return cudaGetLastError();
}

size_t zstd_decode_scratch_bytes(int num_sms, int blocks_per_sm) {
    return (size_t)num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1) * ZD_WARPS * (128u << 10);
}

cudaError_t launch_zstd_decode(const InflateArgs &a, int num_sms, int blocks_per_sm, void *lit_scratch, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t want = (a.n_reads + ZD_WARPS - 1) / ZD_WARPS;
    uint64_t cap = (uint64_t)num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (!grid) grid = 1;
    zstd_decode_kernel<<<grid, ZD_WARPS * 32, sizeof(ZdWarpSmem) * ZD_WARPS, st>>>(a, static_cast<uint8_t *>(lit_scratch), 128u << 10);
    return cudaGetLastError();
}

}  // namespace s5b
