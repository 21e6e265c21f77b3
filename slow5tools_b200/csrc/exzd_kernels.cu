// exzd_kernels.cu -- sm_100a kernels for the "ex-zd" signal codec (QTS shift, 16-bit zigzag-delta, one byte per
// value with an exception list for values above 255), one warp per read.
//
// Replaces, for whole batches, the reference CPU routines
//   ptr_compress_ex_zd / _v0        slow5lib/src/slow5_press.c:1778 / :1721-1776
//   ex_zd_press_16 / ex_press       :1596-1628 / :1263-1424   (find_qts / do_qts :1675-1711, zigdelta_16_u16 :1573-1594)
//   ptr_depress_ex_zd / _v0         :1824-1848 / :1787-1822
//   ex_zd_depress_16 / ex_depress   :1646-1673 / :1441-1561   (unzigdelta_u16_16 :1635-1645, do_rev_qts_inplace :1713)
// Output bytes are identical to the reference's (tests/test_exzd_gpu.py checks against the oracle and the
// compiled reference).
//
// Stream layout (little endian):
//   u8 version=0 | u64 nin | u8 q | u16 zd[0] | u32 nex |
//   nex > 1 : u32 lenP, svb(pos[0], pos[i]-pos[i-1]-1 ..) | u32 lenE, svb(zd-256 ..)      nex == 1 : u32 pos, u32 zd-256
//   one byte per non-exception value of zd[1..nin), in order                (svb = plain StreamVByte: keys, then data)
//
// Encode makes three passes over the read (the second and third come out of L2): (A) OR of all samples -> q;
// (B) exception count and the byte sizes of the two svb sections, which fix every section's offset; (C) the
// emitting pass -- exception positions / values are written straight to their final place (data bytes by the
// owning lane, key bits OR-ed into a sliding shared-memory window), the byte stream is assembled in shared memory
// and leaves with bulk shared->global copies.  The reference works in a buffer of count + 1024 bytes and aborts
// (SLOW5_ASSERT) when the stream does not fit (:1728); such reads get S5B_ERR_PRESS here.
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "../../include/slow5b200.h"

namespace s5b {
namespace {

// inclusive warp scan; shfl.up's predicate says whether the source lane exists, so each step is SHFL + one predicated add
__device__ __forceinline__ uint32_t scan_incl(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .u32 t;\n\t"
            "shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t"
            "@p add.u32 %0, %0, t;\n\t}"
            : "+r"(v)
            : "r"(d));
    }
    return v;
}
__device__ __forceinline__ uint64_t next_read(unsigned long long *counter, int lane) {
    unsigned long long r = 0;
    if (lane == 0) r = atomicAdd(counter, 1ULL);
    return __shfl_sync(FULL, r, 0);
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// 16-bit zigzag of a delta that wraps mod 2^16 (zigzag_one_16 takes an int16_t, slow5_press.c:1567-1570)
__device__ __forceinline__ uint32_t zz16(int d) {
    const int x = (int)(short)d;
    return (uint32_t)((x + x) ^ (x >> 15)) & 0xFFFFu;
}
__device__ __forceinline__ int unzz16(uint32_t z) { return (int)(z >> 1) ^ -(int)(z & 1u); }
// StreamVByte "1234" byte count of a value (streamvbyte_encode.c:31-54)
__device__ __forceinline__ uint32_t svb_bytes(uint32_t v) { return 1u + (v > 0xFFu) + (v > 0xFFFFu) + (v > 0xFFFFFFu); }

constexpr int XE_WARPS = 8;
constexpr int XE_DB = 16 + 4 * 256 + 16;  // byte stream of four iterations (+ carried partial segment)
constexpr int XE_KWIN = 128;              // key-window bytes per svb section (512 exceptions)

struct __align__(128) XeWarpSmem {
    uint8_t dbuf[XE_DB];
    uint32_t kwin[2][XE_KWIN / 4];  // [0] positions, [1] values
    uint32_t list[256];             // this iteration's exceptions, in order: (value index << 16) | zd
};

// the lane's 8 samples of iteration `base` (values i = base + 8*lane + k), shifted right by q, plus the sample before them
struct Samples {
    int x[8];
    int prev;
};
__device__ __forceinline__ uint4 load_raw(const int16_t *sig, uint32_t it, uint32_t n, int lane) {
    const uint32_t i0 = it * 256 + 8 * lane;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (i0 < n) w = __ldg(reinterpret_cast<const uint4 *>(sig + i0));  // slots are padded to 8 samples
    return w;
}
__device__ __forceinline__ Samples shift_samples(const uint4 w, uint32_t q, int lane, int &carry) {
    Samples s;
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.x[2 * k] = (int)(short)(ww[k] & 0xFFFFu) >> q;  // do_qts: arithmetic shift of the int16 sample (:1700-1711)
        s.x[2 * k + 1] = (int)ww[k] >> (16 + q);
    }
    const int up = __shfl_up_sync(FULL, s.x[7], 1);
    s.prev = lane ? up : carry;
    carry = __shfl_sync(FULL, s.x[7], 31);
    return s;
}

// Size pass of the encoder: exception count and the data bytes of the position / value sections for a given q, plus
// zd[0] and the OR of all samples.  Loop-free per lane: only a lane's first exception can be more than 255 values after
// its predecessor, and a value-section entry takes two bytes exactly when zd > 511.
struct Sizes {
    uint32_t nex, pdt, edt, zd0, orv;
};
__device__ __forceinline__ Sizes size_pass(const int16_t *sig, const uint32_t n, const uint32_t iters, const uint32_t q,
                                           const int lane, const uint32_t lt) {
    uint32_t cnt = 0, pd = 0, ed = 0, zd0 = 0, orv = 0;
    int carry = 0;
    int lastpos = -1;  // warp-uniform: position (index into zd[1..]) of the last exception so far
    uint4 w = load_raw(sig, 0, n, lane);
    for (uint32_t it = 0; it < iters; ++it) {
        const uint4 wn = it + 1 < iters ? load_raw(sig, it + 1, n, lane) : make_uint4(0, 0, 0, 0);
        const uint32_t base = it * 256;
        const uint32_t i0 = base + 8 * lane;
        if (i0 + 8 <= n) {
            orv |= w.x | w.y | w.z | w.w;
        } else {
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (i0 + k < n) orv |= (ww[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
        }
        const Samples s = shift_samples(w, q, lane, carry);
        w = wn;
        int prev = s.prev;
        uint32_t f = 0, g = 0, z0 = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t z = zz16(s.x[k] - prev);
            prev = s.x[k];
            if (k == 0) z0 = z;
            const uint32_t i = i0 + k;
            const bool valid = i >= 1 && i < n;
            if (valid && z > 255u) f |= 1u << k;
            if (valid && z > 511u) g |= 1u << k;
        }
        if (it == 0) zd0 = __shfl_sync(FULL, z0, 0);
        const uint32_t B = __ballot_sync(FULL, f != 0);
        if (B) {
            const int p0 = (int)i0 - 1;  // position of the lane's value k is p0 + k
            const int own_last = p0 + (31 - __clz((int)(f | 1u)));
            const uint32_t lower = B & lt;
            const int src = lower ? 31 - __clz((int)lower) : 0;
            const int got = __shfl_sync(FULL, own_last, src);
            const int pp = lower ? got : lastpos;
            if (f) {
                const uint32_t c = __popc(f);
                const int first = p0 + (__ffs((int)f) - 1);
                cnt += c;
                pd += c - 1 + svb_bytes((uint32_t)(first - pp - 1));
                ed += c + __popc(g);
            }
            lastpos = __shfl_sync(FULL, own_last, 31 - __clz((int)B));
        }
    }
    Sizes r;
    r.nex = __reduce_add_sync(FULL, cnt);
    r.pdt = __reduce_add_sync(FULL, pd);
    r.edt = __reduce_add_sync(FULL, ed);
    r.zd0 = zd0;
    r.orv = __reduce_or_sync(FULL, (orv | (orv >> 16)) & 0xFFFFu);
    return r;
}

// ---- byte-stream output through shared memory (same scheme as the svb-zd encoder's data stream)
struct OutStream {
    uint8_t *gbase;  // 16-byte aligned global address of buf[0]
    uint8_t *buf;
    uint32_t pos;   // bytes appended
    uint32_t head;  // first valid byte of segment 0
};
__device__ __forceinline__ void out_drain(OutStream &d, const int lane) {
    const uint32_t nseg = d.pos >> 4;
    uint32_t first = 0;
    if (d.head && nseg) {
        if (lane >= (int)d.head && lane < 16) d.gbase[lane] = d.buf[lane];
        d.head = 0;
        first = 1;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        if (nseg > first) bulk_s2g(d.gbase + first * 16, smem_u32(d.buf) + first * 16, (nseg - first) * 16);
        bulk_commit();
        bulk_wait_read<0>();
    }
    __syncwarp();
    const uint32_t rem = d.pos & 15u;
    uint8_t t = 0;
    if (lane < (int)rem) t = d.buf[nseg * 16 + lane];
    __syncwarp();
    if (lane < (int)rem) d.buf[lane] = t;
    d.gbase += nseg * 16;
    d.pos = rem;
    __syncwarp();
}

// Flushes the complete key bytes of both svb sections' windows to global memory and slides the windows (they share
// the exception rank, so they move together).  rank: exceptions emitted so far; wbase: rank of the windows' first
// 2-bit slot (a multiple of 4).  The last flush also writes the partial key byte (zero padded, as the reference's).
__device__ __forceinline__ void keys_flush(XeWarpSmem &ws, uint8_t *keys_p, uint8_t *keys_e, uint32_t &wbase,
                                           const uint32_t rank, const bool final, const int lane) {
    __syncwarp();
    const uint8_t *kp = reinterpret_cast<const uint8_t *>(ws.kwin[0]);
    const uint8_t *ke = reinterpret_cast<const uint8_t *>(ws.kwin[1]);
    const uint32_t nfull = (rank - wbase) >> 2;
    const uint32_t nout = final ? ((rank - wbase + 3) >> 2) : nfull;
    for (uint32_t i = lane; i < nout; i += 32) {
        keys_p[(wbase >> 2) + i] = kp[i];
        keys_e[(wbase >> 2) + i] = ke[i];
    }
    uint32_t carry_p = 0, carry_e = 0;
    if (!final && nfull < XE_KWIN) {
        carry_p = kp[nfull];
        carry_e = ke[nfull];
    }
    __syncwarp();
    for (uint32_t i = lane; i < XE_KWIN / 4; i += 32) ws.kwin[0][i] = ws.kwin[1][i] = 0;
    __syncwarp();
    if (lane == 0) {
        ws.kwin[0][0] = carry_p;
        ws.kwin[1][0] = carry_e;
    }
    wbase += nfull * 4;
    __syncwarp();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(XE_WARPS * 32, 4) exzd_encode_kernel(const SvbEncodeArgs a) {
    __shared__ XeWarpSmem smem[XE_WARPS];
    const int lane = threadIdx.x & 31;
    XeWarpSmem &ws = smem[threadIdx.x >> 5];
    const uint32_t lt = lanemask_lt();

    for (;;) {
        const uint64_t r = next_read(a.work_counter, lane);
        if (r >= a.n_reads) break;
        const uint32_t n = a.n_samples[r];
        const uint64_t soff = a.sig_off[r];
        const uint64_t scap = a.sig_off[r + 1] - soff;
        const uint64_t ooff = a.svb_off[r];
        const uint64_t ocap = a.svb_off[r + 1] - ooff;
        int32_t st = S5B_OK;
        if ((soff & 7) || scap < n || n == 0) st = S5B_ERR_ARG;  // an empty read is undefined in the reference (:1612)
        else if (ocap < 16) st = S5B_ERR_NOSPACE;
        if (st != S5B_OK) {
            if (lane == 0) {
                a.status[r] = st;
                a.svb_len[r] = 0;
            }
            continue;
        }
        const int16_t *sig = a.sig + soff;
        uint8_t *dst = a.svb + ooff;
        const uint32_t iters = (n + 255) >> 8;

        // ---- passes A + B: q = shared low zero bits, at most 5 (find_qts, :1675-1698), and the section sizes.  The size
        // pass runs speculatively with q = 0 (one odd sample makes it so) while OR-ing the samples; only reads whose
        // samples share low zero bits -- lossy "degrade"d data -- pay a second size pass with their real q
        Sizes sz = size_pass(sig, n, iters, 0, lane, lt);
        const uint32_t q = (sz.orv & 31u) ? (uint32_t)(__ffs((int)(sz.orv & 31u)) - 1) : 5u;
        if (q) sz = size_pass(sig, n, iters, q, lane, lt);
        const uint32_t nex = sz.nex, pdt = sz.pdt, edt = sz.edt, zd0 = sz.zd0;
        const uint32_t nkeys = (nex + 3) >> 2;
        // section offsets
        uint64_t o_keysP = 16, o_dataP = 16, o_keysE = 16, o_dataE = 16, o_bytes = 16;
        if (nex > 1) {
            o_keysP = 16 + 4;
            o_dataP = o_keysP + nkeys;
            o_keysE = o_dataP + pdt + 4;
            o_dataE = o_keysE + nkeys;
            o_bytes = o_dataE + edt;
        } else if (nex == 1) {
            o_bytes = 16 + 8;
        }
        const uint64_t total = o_bytes + (uint64_t)(n - 1 - nex);
        // the reference's working buffer is count + 1024 bytes and it aborts when the stream outgrows it (:1728)
        if (total > 2ull * n + 1024ull || total > ocap) {
            if (lane == 0) {
                a.status[r] = total > 2ull * n + 1024ull ? S5B_ERR_PRESS : S5B_ERR_NOSPACE;
                a.svb_len[r] = 0;
            }
            continue;
        }
        // ---- header
        if (lane == 0) {
            dst[0] = 0;
            const uint64_t nin = n;
            for (int b = 0; b < 8; ++b) dst[1 + b] = (uint8_t)(nin >> (8 * b));
            dst[9] = (uint8_t)q;
            dst[10] = (uint8_t)zd0;
            dst[11] = (uint8_t)(zd0 >> 8);
            for (int b = 0; b < 4; ++b) dst[12 + b] = (uint8_t)(nex >> (8 * b));
            if (nex > 1) {
                const uint32_t lenP = nkeys + pdt, lenE = nkeys + edt;
                for (int b = 0; b < 4; ++b) {
                    dst[16 + b] = (uint8_t)(lenP >> (8 * b));
                    dst[o_dataP + pdt + b] = (uint8_t)(lenE >> (8 * b));
                }
            }
        }
        // ---- pass C: emit
        for (uint32_t i = lane; i < XE_KWIN / 4; i += 32) ws.kwin[0][i] = ws.kwin[1][i] = 0;
        __syncwarp();
        OutStream os;
        os.buf = ws.dbuf;
        os.head = os.pos = (uint32_t)(reinterpret_cast<uintptr_t>(dst + o_bytes) & 15u);
        os.gbase = dst + o_bytes - os.head;
        {
            int carry = 0;
            int lastpos = -1;
            uint32_t R = 0, PD = 0, ED = 0;  // exceptions / section data bytes emitted so far (warp-uniform)
            uint32_t wbase = 0;              // first rank covered by the key windows
            for (uint32_t it = 0; it < iters; ++it) {
                const uint32_t base = it * 256;
                const Samples s = shift_samples(load_raw(sig, it, n, lane), q, lane, carry);
                int prev = s.prev;
                uint32_t f = 0, vmask = 0;
                uint32_t z[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    z[k] = zz16(s.x[k] - prev);
                    prev = s.x[k];
                    const uint32_t i = base + 8 * lane + k;
                    if (i >= 1 && i < n) {
                        vmask |= 1u << k;
                        if (z[k] > 255u) f |= 1u << k;
                    }
                }
                const uint32_t B = __ballot_sync(FULL, f != 0);
                uint32_t excl_cnt = 0, it_cnt = 0;
                if (B) {
                    // compact the iteration's exceptions, in order, into a shared list: (value index << 16) | zd
                    const uint32_t c = __popc(f);
                    const uint32_t incl = scan_incl(c);
                    excl_cnt = incl - c;
                    it_cnt = __shfl_sync(FULL, incl, 31);
                    {
                        uint32_t rk = excl_cnt;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (f & (1u << k)) ws.list[rk] = ((8u * lane + k) << 16) | z[k];
                            rk += (f >> k) & 1u;
                        }
                    }
                    // make room in the key windows for this iteration's (at most 256) exceptions
                    if (nex > 1 && R - wbase > 256u) keys_flush(ws, dst + o_keysP, dst + o_keysE, wbase, R, false, lane);
                    __syncwarp();
                    // one exception per lane: delta to the previous position, svb sizes, offsets by a packed scan
                    for (uint32_t j0 = 0; j0 < it_cnt; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        const bool act = j < it_cnt;
                        const uint32_t e = act ? ws.list[j] : 0u;
                        const int pos = (int)(base + (e >> 16)) - 1;
                        int pp = lastpos;
                        if (act && j) pp = (int)(base + (ws.list[j - 1] >> 16)) - 1;
                        const uint32_t dv = (uint32_t)(pos - pp - 1);
                        const uint32_t ev = (e & 0xFFFFu) - 256u;
                        const uint32_t bd = act ? svb_bytes(dv) : 0u, be = act ? 1u + (ev > 0xFFu) : 0u;
                        const uint32_t packed = bd | (be << 16);
                        const uint32_t inc2 = scan_incl(packed);
                        const uint32_t tot = __shfl_sync(FULL, inc2, 31);
                        if (act) {
                            if (nex > 1) {
                                uint8_t *gp = dst + o_dataP + PD + ((inc2 - packed) & 0xFFFFu);
                                uint8_t *ge = dst + o_dataE + ED + ((inc2 - packed) >> 16);
                                gp[0] = (uint8_t)dv;
                                if (bd > 1) gp[1] = (uint8_t)(dv >> 8);
                                if (bd > 2) gp[2] = (uint8_t)(dv >> 16);
                                if (bd > 3) gp[3] = (uint8_t)(dv >> 24);
                                ge[0] = (uint8_t)ev;
                                if (be > 1) ge[1] = (uint8_t)(ev >> 8);
                                const uint32_t slot = R + j - wbase;  // 2-bit slot inside the windows
                                atomicOr(&ws.kwin[0][slot >> 4], (bd - 1) << (2 * (slot & 15u)));
                                atomicOr(&ws.kwin[1][slot >> 4], (be - 1) << (2 * (slot & 15u)));
                            } else {  // a single exception is stored raw (:1405-1411)
                                for (int b2 = 0; b2 < 4; ++b2) {
                                    dst[16 + b2] = (uint8_t)(dv >> (8 * b2));
                                    dst[20 + b2] = (uint8_t)(ev >> (8 * b2));
                                }
                            }
                        }
                        PD += tot & 0xFFFFu;
                        ED += tot >> 16;
                    }
                    lastpos = (int)(base + (ws.list[it_cnt - 1] >> 16)) - 1;
                    R += it_cnt;
                    __syncwarp();
                }
                // byte stream: the lane's non-exception values, in order
                {
                    const uint32_t lo = base ? base : 1u;
                    const uint32_t hi = min(base + 256u, n);
                    const uint32_t i0 = min(max(base + 8u * lane, lo), hi);
                    uint8_t *p = os.buf + os.pos + (i0 - lo) - excl_cnt;
                    const uint32_t keep = vmask & ~f;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (keep & (1u << k)) *p++ = (uint8_t)z[k];
                    os.pos += (hi > lo ? hi - lo : 0u) - it_cnt;
                }
                __syncwarp();
                if ((it & 3u) == 3u) out_drain(os, lane);
            }
            if (nex > 1) keys_flush(ws, dst + o_keysP, dst + o_keysE, wbase, R, true, lane);
        }
        out_drain(os, lane);
        if (lane >= (int)os.head && lane < (int)os.pos) os.gbase[lane] = os.buf[lane];
        if (lane == 0) {
            a.svb_len[r] = (uint32_t)total;
            a.status[r] = S5B_OK;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
constexpr int XD_WARPS = 8;
constexpr int XD_QCAP = 512;  // exception queue entries (power of two, >= 256 in one iteration + one batch of 32)

struct __align__(16) XdWarpSmem {
    uint32_t qpos[XD_QCAP];  // position (index into zd[1..]) of queued exceptions, ring indexed by exception number
    uint16_t qval[XD_QCAP];  // their zd values
    uint32_t bitmap[8];      // exception flags of the current 256-value iteration
    uint16_t ztmp[256];      // exception values of the current iteration, by value index
};

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

__global__ void __launch_bounds__(XD_WARPS * 32) exzd_decode_kernel(const SvbDecodeArgs a) {
    __shared__ XdWarpSmem smem[XD_WARPS];
    const int lane = threadIdx.x & 31;
    XdWarpSmem &ws = smem[threadIdx.x >> 5];

    for (;;) {
        const uint64_t r = next_read(a.work_counter, lane);
        if (r >= a.n_reads) break;
        const uint64_t ioff = a.svb_off[r];
        const uint32_t ilen = a.svb_len[r];
        const uint8_t *p = a.svb + ioff;
        const uint64_t soff = a.sig_off[r];
        const uint64_t scap = a.sig_off[r + 1] - soff;
        int32_t st = S5B_OK;
        uint64_t nin = 0;
        uint32_t q = 0, zd0 = 0, nex = 0;
        if (ilen < 16 || ioff + ilen > a.svb_capacity || (soff & 7)) {
            st = S5B_ERR_ARG;
        } else if (p[0] != 0) {
            st = S5B_ERR_PRESS;  // unsupported ex-zd version (:1838-1842)
        } else {
            nin = (uint64_t)ld_u32(p + 1) | ((uint64_t)ld_u32(p + 5) << 32);
            q = p[9];
            zd0 = (uint32_t)p[10] | ((uint32_t)p[11] << 8);
            nex = ld_u32(p + 12);
            if (nin == 0 || nin > 0xFFFFFFFFull) st = S5B_ERR_ARG;
            else if (q > 5 || (uint64_t)nex > nin - 1) st = S5B_ERR_PRESS;
            else if (scap < nin) st = S5B_ERR_NOSPACE;
        }
        // ---- sections
        const uint32_t n = (uint32_t)nin;
        const uint32_t m = n - 1;  // values in the exception-coded part
        const uint32_t nkeys = (nex + 3) >> 2;
        const uint8_t *keysP = p, *dataP = p, *keysE = p, *dataE = p, *bytes = p;
        uint32_t dP = 0, dE = 0, pos1 = 0, ex1 = 0;
        if (st == S5B_OK) {
            uint64_t off = 16;
            if (nex > 1) {
                for (int sct = 0; sct < 2 && st == S5B_OK; ++sct) {
                    if (off + 4 > ilen) {
                        st = S5B_ERR_PRESS;
                        break;
                    }
                    const uint32_t len = ld_u32(p + off);
                    if (len < nkeys || off + 4 + len > ilen) {
                        st = S5B_ERR_PRESS;
                        break;
                    }
                    if (sct == 0) {
                        keysP = p + off + 4;
                        dataP = keysP + nkeys;
                        dP = len - nkeys;
                    } else {
                        keysE = p + off + 4;
                        dataE = keysE + nkeys;
                        dE = len - nkeys;
                    }
                    off += 4 + (uint64_t)len;
                }
            } else if (nex == 1) {
                if (off + 8 > ilen) st = S5B_ERR_PRESS;
                else {
                    pos1 = ld_u32(p + off);
                    ex1 = ld_u32(p + off + 4);
                    off += 8;
                    if (pos1 >= m) st = S5B_ERR_PRESS;
                }
            }
            if (st == S5B_OK) {
                bytes = p + off;
                if ((uint64_t)ilen - off != (uint64_t)(m - nex)) st = S5B_ERR_PRESS;  // one byte per remaining value
            }
        }
        if (st != S5B_OK) {
            if (lane == 0) {
                a.status[r] = st;
                a.n_samples[r] = (uint32_t)nin;
            }
            continue;
        }
        int16_t *out = a.sig + soff;

        // exception queue state (warp-uniform)
        uint32_t decoded = 0, consumed = 0, pdo = 0, edo = 0;
        uint64_t lastpos = ~0ull;  // "-1": pos = lastpos + (delta + 1)
        bool ok = true;
        if (nex == 1) {
            if (lane == 0) {
                ws.qpos[0] = pos1;
                ws.qval[0] = (uint16_t)(ex1 + 256u);
            }
            decoded = 1;
            lastpos = pos1;
            __syncwarp();
        }
        uint32_t bytes_done = 0;
        uint32_t acc = 0;  // running sum (mod 2^16 is all that survives the int16 store)
        const uint32_t iters = (n + 255) >> 8;
        for (uint32_t it = 0; it < iters && ok; ++it) {
            const uint32_t base = it * 256;
            // ---- refill: every exception of this iteration (value index pos + 1 < base + 256) must be queued
            while (ok && decoded < nex && (lastpos == ~0ull || lastpos + 1 < (uint64_t)base + 256)) {
                const uint32_t e = decoded + lane;
                const bool act = e < nex;
                uint32_t bp = 0, be = 0;
                if (act) {
                    bp = 1 + ((keysP[e >> 2] >> (2 * (e & 3))) & 3u);
                    be = 1 + ((keysE[e >> 2] >> (2 * (e & 3))) & 3u);
                }
                const uint32_t packed = bp | (be << 16);
                const uint32_t incl = scan_incl(packed);
                const uint32_t tot = __shfl_sync(FULL, incl, 31);
                const uint32_t op = pdo + ((incl - packed) & 0xFFFFu), oe = edo + ((incl - packed) >> 16);
                if (pdo + (tot & 0xFFFFu) > dP || edo + (tot >> 16) > dE) {  // svb section shorter than its keys say
                    ok = false;
                    break;
                }
                uint32_t dv = 0, ev = 0;
                for (uint32_t b = 0; b < bp; ++b) dv |= (uint32_t)dataP[op + b] << (8 * b);
                for (uint32_t b = 0; b < be; ++b) ev |= (uint32_t)dataE[oe + b] << (8 * b);
                // undelta (:1427-1438): pos[i] = pos[i-1] + delta + 1, 64-bit so an absurd delta cannot wrap around
                uint64_t step = act ? (uint64_t)dv + 1 : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint64_t t = __shfl_up_sync(FULL, step, d);
                    if (lane >= d) step += t;
                }
                const uint64_t pos = lastpos + step;
                if (__any_sync(FULL, act && pos >= (uint64_t)m)) {
                    ok = false;
                    break;
                }
                if (act) {
                    ws.qpos[e & (XD_QCAP - 1)] = (uint32_t)pos;
                    ws.qval[e & (XD_QCAP - 1)] = (uint16_t)(ev + 256u);
                }
                const uint32_t nb = min(32u, nex - decoded);
                lastpos = __shfl_sync(FULL, pos, nb - 1);
                decoded += nb;
                pdo += tot & 0xFFFFu;
                edo += tot >> 16;
                __syncwarp();
            }
            if (!ok) break;
            // ---- this iteration's exceptions: flags + values by value index
            if (lane < 8) ws.bitmap[lane] = 0;
            __syncwarp();
            uint32_t it_cnt = 0;
            for (;;) {
                const uint32_t e = consumed + lane;
                bool inr = false;
                if (e < decoded) {
                    const uint32_t ps = ws.qpos[e & (XD_QCAP - 1)];
                    if ((uint64_t)ps + 1 < (uint64_t)base + 256) {
                        inr = true;
                        const uint32_t idx = ps + 1 - base;
                        atomicOr(&ws.bitmap[idx >> 5], 1u << (idx & 31u));
                        ws.ztmp[idx] = ws.qval[e & (XD_QCAP - 1)];
                    }
                }
                const uint32_t c = __popc(__ballot_sync(FULL, inr));
                consumed += c;
                it_cnt += c;
                if (c < 32) break;
            }
            __syncwarp();
            const uint32_t f = it_cnt ? (ws.bitmap[lane >> 2] >> (8 * (lane & 3))) & 0xFFu : 0u;
            uint32_t excl_cnt = 0;
            if (it_cnt) {
                const uint32_t c = __popc(f);
                excl_cnt = scan_incl(c) - c;
            }
            const uint32_t lo = base ? base : 1u;
            const uint32_t hi = min(base + 256u, n);
            const uint32_t i0 = min(max(base + 8u * lane, lo), hi);
            const uint8_t *bp = bytes + bytes_done + (i0 - lo) - excl_cnt;
            // the lane's (at most 8) stream bytes: three aligned word loads realigned into a 64-bit window; the last
            // lanes of a slab, where the words would run past it, read byte by byte
            uint64_t win = 0;
            {
                const uint8_t *wa = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(bp) & ~uintptr_t(3));
                if (wa + 12 <= a.svb + a.svb_capacity) {
                    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(wa);
                    const uint32_t w0 = __ldg(w32), w1 = __ldg(w32 + 1), w2 = __ldg(w32 + 2);
                    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(bp) & 3u) * 8;
                    win = (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
                } else {
                    const uint32_t need = __popc(~f & 0xFFu);
                    for (uint32_t t = 0; t < need && bp + t < p + ilen; ++t) win |= (uint64_t)bp[t] << (8 * t);
                }
            }
            // ---- values, zigzag decode, running sum (unzigdelta_u16_16 :1635-1645), QTS shift back (:1713-1718)
            uint32_t sum[8];
            uint32_t run = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t i = base + 8 * lane + k;
                uint32_t z = 0;
                if (i < n) {
                    if (i == 0) z = zd0;
                    else if (f & (1u << k)) z = ws.ztmp[8 * lane + k];
                    else {
                        z = (uint32_t)win & 0xFFu;
                        win >>= 8;
                    }
                }
                run += (uint32_t)unzz16(z);
                sum[k] = run;
            }
            const uint32_t incl_sum = scan_incl(run);
            const uint32_t basev = acc + incl_sum - run;
            acc += __shfl_sync(FULL, incl_sum, 31);
            const uint32_t i8 = base + 8 * lane;
            if (i8 + 8 <= n) {
                uint32_t o16[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o16[k] = ((sum[k] + basev) << q) & 0xFFFFu;
                uint4 w;
                w.x = o16[0] | (o16[1] << 16);
                w.y = o16[2] | (o16[3] << 16);
                w.z = o16[4] | (o16[5] << 16);
                w.w = o16[6] | (o16[7] << 16);
                *reinterpret_cast<uint4 *>(out + i8) = w;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (i8 + k < n) out[i8 + k] = (int16_t)(uint16_t)(((sum[k] + basev) << q) & 0xFFFFu);
            }
            bytes_done += (hi > lo ? hi - lo : 0u) - it_cnt;
            {   // the byte stream is consumed front to back, at most 256 bytes per iteration: pull the lines two iterations
                // ahead towards the SM so the dependent loads above do not wait on DRAM
                const uint8_t *pf = bytes + bytes_done + 256 + 16 * lane;
                if (pf < p + ilen) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
            }
            __syncwarp();
        }
        // every exception used, both svb sections consumed exactly (the reference compares the decoder's byte count with
        // the stored length, :1492-1500 / :1521-1529)
        if (ok && (consumed != nex || (nex > 1 && (pdo != dP || edo != dE)))) ok = false;
        if (lane == 0) {
            a.n_samples[r] = n;
            a.status[r] = ok ? S5B_OK : S5B_ERR_PRESS;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static unsigned exzd_grid(uint64_t n_reads, int warps, int num_sms, int blocks_per_sm) {
    uint64_t want = (n_reads + warps - 1) / warps;
    uint64_t cap = (uint64_t)num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
    uint64_t g = want < cap ? want : cap;
    return (unsigned)(g ? g : 1);
}
int exzd_encode_blocks_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, exzd_encode_kernel, XE_WARPS * 32, 0) != cudaSuccess) return 0;
    return n;
}
int exzd_decode_blocks_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, exzd_decode_kernel, XD_WARPS * 32, 0) != cudaSuccess) return 0;
    return n;
}
uint64_t exzd_bound(uint32_t n_samples) { return 2ull * n_samples + 1024ull; }  // the reference's working buffer (:1728)
cudaError_t launch_exzd_encode(const SvbEncodeArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    exzd_encode_kernel<<<exzd_grid(a.n_reads, XE_WARPS, num_sms, blocks_per_sm), XE_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_exzd_decode(const SvbDecodeArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    exzd_decode_kernel<<<exzd_grid(a.n_reads, XD_WARPS, num_sms, blocks_per_sm), XD_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace s5b
