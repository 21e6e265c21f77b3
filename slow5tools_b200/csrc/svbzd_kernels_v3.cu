// svbzd_kernels_v3.cu -- the round-1 "v3" svb-zd kernels, kept for A/B timing against the current ones
// (S5B_SVBZD_LEGACY=1 selects them at context creation).  sm_100a kernels for the svb-zd signal codec (StreamVByte "1234" coding of
// zigzag-delta values, u32 sample-count header), one warp per read.
//
// Replaces, for whole batches, the reference CPU routines
//   ptr_compress_svb_zd / ptr_compress_svb      slow5lib/src/slow5_press.c:1082-1115 / :1062-1079
//   ptr_depress_svb_zd  / ptr_depress_svb       slow5lib/src/slow5_press.c:1143-1173 / :1118-1140
//   __slow5_zigzag_delta_encode / _decode       thirdparty/streamvbyte/src/streamvbyte_zigzag.c:15-40
//   __slow5_streamvbyte_encode / _decode        thirdparty/streamvbyte/src/streamvbyte_{en,de}code.c
// Output bytes are identical to the reference's (tests/test_svbzd_gpu.py checks against the oracle).
//
// Data movement (both kernels are HBM-streaming, integer-only, no tensor cores):
//   * each warp owns one read at a time (dynamic work counter), loops over 256-sample iterations
//     (8 samples / lane, one 128-bit shared-memory load per lane);
//   * encode: the int16 signal is staged HBM->smem by 1-D bulk async copies (TMA engine, UBLKCP)
//     into a 2-stage per-warp pipeline guarded by mbarriers; a warp prefix scan over the per-lane
//     byte counts gives every lane its data offset; key and data bytes are assembled in smem and
//     leave as 128-bit coalesced stores;
//   * decode: the variable-length data stream is staged by bulk async copies into a 4 x 1 KiB
//     per-warp ring; a warp prefix scan over the control-byte lengths resolves the per-lane data
//     offsets, a second scan rebuilds the running sum; samples leave as 128-bit coalesced stores.
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "../../include/slow5b200.h"

namespace s5b {
namespace v3 {

// ------------------------------------------------------------------------------------------------
// common helpers
// ------------------------------------------------------------------------------------------------
// inclusive warp scan; shfl.up's predicate says whether the source lane exists, so each step is
// SHFL + one predicated add
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
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

// sign-extended low half of a packed int16 pair in one PRMT (selector nibble 9 = byte 1's sign replicated;
// the __byte_perm intrinsic masks that bit away, so this is inline PTX)
__device__ __forceinline__ int sext_lo16(uint32_t w) {
    int r;
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(w));
    return r;
}

// next read index for this warp (uniform across lanes)
__device__ __forceinline__ uint64_t next_work(unsigned long long *counter, int lane) {
    unsigned long long r = 0;
    if (lane == 0) r = atomicAdd(counter, 1ULL);
    return __shfl_sync(FULL, r, 0);
}

// ------------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------------
constexpr int ENC_WARPS = 8;
constexpr int ENC_CH_SAMPLES = 1024;  // samples per bulk-copy chunk (4 iterations of 256)
constexpr int ENC_CH_BYTES = ENC_CH_SAMPLES * 2;
constexpr int ENC_STAGES = 2;
constexpr int ENC_DBUF = 16 + 3 * ENC_CH_SAMPLES + 16;  // data bytes of one chunk (+ carried partial segment)

struct __align__(128) EncWarpSmem {
    uint8_t in[ENC_STAGES][ENC_CH_BYTES];  // staged signal
    uint8_t dbuf[ENC_DBUF];                // dbuf[0] <-> 16-byte aligned global address
    uint8_t kbuf[ENC_CH_SAMPLES / 4];      // key bytes of one chunk, natural index
    unsigned long long bar[ENC_STAGES];
};

// Per-warp output state of the data stream (all members warp-uniform).  Data bytes of a chunk are
// appended linearly to dbuf; complete 16-byte segments leave after every iteration as 128-bit stores;
// at the end of a chunk the (< 16 byte) remainder moves to the front.
struct EncData {
    uint8_t *gbase;  // 16-byte aligned global address of dbuf[0]
    uint32_t pos;    // bytes appended (index into dbuf)
    uint32_t fseg;   // 16-byte segments of dbuf already stored
    uint32_t head;   // first valid byte of segment 0 (stream start not 16-byte aligned), else 0
};

__device__ __forceinline__ void enc_flush_segments(EncData &d, const uint8_t *dbuf, const int lane) {
    // an iteration appends at most 768 bytes -> at most 49 new complete segments -> two store rounds
    const uint32_t wseg = d.pos >> 4;
    if (wseg > d.fseg) {
        uint32_t seg = d.fseg + lane;
        if (d.head) {  // ragged stream start: segment 0 leaves as byte stores
            if (lane >= (int)d.head && lane < 16) d.gbase[lane] = dbuf[lane];
            d.head = 0;
            seg += 1;
        }
        const uint4 *s = reinterpret_cast<const uint4 *>(dbuf);
        uint4 *g = reinterpret_cast<uint4 *>(d.gbase);
        if (seg < wseg) g[seg] = s[seg];
        if (seg + 32 < wseg) g[seg + 32] = s[seg + 32];
        d.fseg = wseg;
    }
}

// Second half of an iteration: codes, keys, prefix scan, byte assembly.  WIDE (3-byte codes present in
// the warp, |delta| >= 32768) is a separate instantiation so the common path carries 1-bit codes only.
template <bool PARTIAL, bool WIDE>
__device__ __forceinline__ void enc_emit(const uint32_t (&z)[8], const int lane, const int nvalid, EncData &d,
                                         uint8_t *dbuf, uint16_t *kslot) {
    // svb code = bytes - 1 (streamvbyte_encode.c:31-54; code 3 is unreachable from int16 input: z < 2^17)
    uint32_t t[8], c[8];
    uint32_t key = 0, csum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        t[j] = z[j] >> 8;
        c[j] = WIDE ? (z[j] > 0xFFu) + (z[j] > 0xFFFFu) : (z[j] + 0xFF00u) >> 16;
        csum += c[j];
        key |= c[j] << (2 * j);  // 2 bits per value, value i -> byte i/4, shift 2*(i%4) (streamvbyte_encode.c:56-79)
    }
    uint32_t lane_len = 8 + csum;
    if (PARTIAL) lane_len = min(max(nvalid - lane * 8, 0), 8) + csum;  // invalid samples: z = 0, c = 0, no byte
    const uint32_t incl = warp_incl_scan(lane_len);  // prefix scan over per-lane byte counts -> data offsets
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    uint8_t *p = dbuf + d.pos + (incl - lane_len);
    if (!PARTIAL) {
        // Both bytes are stored unconditionally except for the lane's last value: a spurious high byte
        // lands where the same lane's next value puts its low byte afterwards (program order), so it
        // never survives.  The last value must not spill into the next lane's first byte.
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            p[0] = (uint8_t)z[j];
            if (j < 7 || c[j]) p[1] = (uint8_t)t[j];
            if (WIDE && c[j] == 2) p[2] = (uint8_t)(z[j] >> 16);
            p += 1 + c[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (lane * 8 + j < nvalid) {
                p[0] = (uint8_t)z[j];
                if (c[j]) p[1] = (uint8_t)t[j];
                if (WIDE && c[j] == 2) p[2] = (uint8_t)(z[j] >> 16);
                p += 1 + c[j];
            }
        }
    }
    kslot[lane] = (uint16_t)key;  // invalid samples carry code 0, so padding bits are zero
    d.pos += total;
}

// One 256-sample iteration: lane owns samples 8*lane .. 8*lane+7 of the iteration.
template <bool PARTIAL>
__device__ __forceinline__ void enc_iteration(const uint4 w, int &carry, const int lane, const int nvalid,
                                              EncData &d, uint8_t *dbuf, uint16_t *kslot) {
    // widen (slow5_press.c:1095-1097): one PRMT (sign-replicating) / one arithmetic shift per sample
    int x[8];
    x[0] = sext_lo16(w.x);
    x[1] = (int)w.x >> 16;
    x[2] = sext_lo16(w.y);
    x[3] = (int)w.y >> 16;
    x[4] = sext_lo16(w.z);
    x[5] = (int)w.z >> 16;
    x[6] = sext_lo16(w.w);
    x[7] = (int)w.w >> 16;
    const int up = __shfl_up_sync(FULL, (int)w.w, 1);
    int prev = lane ? (up >> 16) : carry;
    carry = __shfl_sync(FULL, (int)w.w, 31) >> 16;

    uint32_t z[8], zor = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int dd = x[j] - prev;  // zigzag-delta, streamvbyte_zigzag.c:4-6,15-20
        prev = x[j];
        z[j] = ((uint32_t)dd << 1) ^ (uint32_t)(dd >> 31);
        if (PARTIAL && lane * 8 + j >= nvalid) z[j] = 0;
        zor |= z[j];
    }
    if (__any_sync(FULL, zor > 0xFFFFu)) {
        enc_emit<PARTIAL, true>(z, lane, nvalid, d, dbuf, kslot);
    } else {
        enc_emit<PARTIAL, false>(z, lane, nvalid, d, dbuf, kslot);
    }
    __syncwarp();
    enc_flush_segments(d, dbuf, lane);
}

__global__ void __launch_bounds__(ENC_WARPS * 32) svbzd_encode_kernel(const SvbEncodeArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    EncWarpSmem *smem = reinterpret_cast<EncWarpSmem *>(smem_raw);
    const int lane = threadIdx.x & 31;
    EncWarpSmem &ws = smem[threadIdx.x >> 5];
    const uint32_t bar0 = smem_u32(&ws.bar[0]);
    const uint32_t in0 = smem_u32(&ws.in[0][0]);
    if (lane == 0) {
        for (int s = 0; s < ENC_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t q = 0;  // chunks consumed by this warp so far (stage = q & 1, parity = (q >> 1) & 1)

    for (;;) {
        const uint64_t r = next_work(a.work_counter, lane);
        if (r >= a.n_reads) break;
        const uint32_t n = a.n_samples[r];
        const uint64_t soff = a.sig_off[r];
        const uint64_t scap = a.sig_off[r + 1] - soff;
        const uint64_t ooff = a.svb_off[r];
        const uint64_t ocap = a.svb_off[r + 1] - ooff;
        const uint32_t nkeys = (n + 3) >> 2;
        int32_t st = S5B_OK;
        if ((soff & 7) || scap < n) st = S5B_ERR_ARG;
        else if (ocap < 4ull + nkeys + 3ull * n) st = S5B_ERR_NOSPACE;
        if (st != S5B_OK) {
            if (lane == 0) {
                a.status[r] = st;
                a.svb_len[r] = 0;
            }
            continue;
        }
        const uint8_t *src = reinterpret_cast<const uint8_t *>(a.sig + soff);
        uint8_t *dst = a.svb + ooff;
        // bytes that may be bulk-copied for this read: whole 16-byte granules inside the read's slot
        // (the slot is a multiple of 8 samples except possibly the last one of the slab)
        const uint64_t slot_bytes16 = (scap * 2) & ~15ull;
        const uint64_t n_bytes = (uint64_t)n * 2;

        if (lane < 4) dst[lane] = (uint8_t)(n >> (8 * lane));  // u32 LE header, slow5_press.c:1074
        uint8_t *kdst = dst + 4;
        uint8_t *ddst = kdst + nkeys;
        EncData d;
        d.head = d.pos = (uint32_t)(reinterpret_cast<uintptr_t>(ddst) & 15u);
        d.gbase = ddst - d.head;
        d.fseg = 0;

        const uint32_t nchunks = (n + ENC_CH_SAMPLES - 1) / ENC_CH_SAMPLES;
        auto issue = [&](uint32_t k, uint32_t qq) {
            // chunk k of this read -> stage qq & 1
            const uint64_t b0 = (uint64_t)k * ENC_CH_BYTES;
            uint64_t want = n_bytes - b0;
            if (want > ENC_CH_BYTES) want = ENC_CH_BYTES;
            want = (want + 15) & ~15ull;
            uint64_t can = slot_bytes16 > b0 ? slot_bytes16 - b0 : 0;
            const uint32_t bytes = (uint32_t)(want < can ? want : can);
            const uint32_t stage = qq & 1;
            if (lane == 0) {
                if (bytes) {
                    mbar_arrive_expect_tx(bar0 + 8 * stage, bytes);
                    bulk_g2s(in0 + stage * ENC_CH_BYTES, src + b0, bytes, bar0 + 8 * stage);
                } else {
                    // nothing bulk-copyable (a < 8-sample tail in the last slot): complete the phase by hand
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * stage) : "memory");
                }
            }
            return bytes;
        };
        int carry = 0;  // prev = 0 for the first sample, slow5_press.c:1106
        uint32_t bytes_cur = nchunks ? issue(0, q) : 0;
        for (uint32_t k = 0; k < nchunks; ++k) {
            uint32_t bytes_next = 0;
            if (k + 1 < nchunks) bytes_next = issue(k + 1, q + 1);
            const uint32_t stage = q & 1;
            mbar_wait(bar0 + 8 * stage, (q >> 1) & 1);
            const uint32_t chunk_samples = min((uint32_t)ENC_CH_SAMPLES, n - k * ENC_CH_SAMPLES);
            if (bytes_cur < chunk_samples * 2) {
                // ragged end of the slab: the bulk copy could not take the last (< 16 byte) piece
                const int16_t *g = a.sig + soff + (uint64_t)k * ENC_CH_SAMPLES;
                int16_t *s = reinterpret_cast<int16_t *>(ws.in[stage]);
                for (uint32_t i = bytes_cur / 2 + lane; i < chunk_samples; i += 32) s[i] = g[i];
                __syncwarp();
            }
            const uint4 *in4 = reinterpret_cast<const uint4 *>(ws.in[stage]);
            uint16_t *k16 = reinterpret_cast<uint16_t *>(ws.kbuf);
            const uint32_t full_iters = chunk_samples >> 8;
            for (uint32_t it = 0; it < full_iters; ++it)
                enc_iteration<false>(in4[it * 32 + lane], carry, lane, 256, d, ws.dbuf, k16 + it * 32);
            const int tail = chunk_samples & 255;
            if (tail) enc_iteration<true>(in4[full_iters * 32 + lane], carry, lane, tail, d, ws.dbuf, k16 + full_iters * 32);
            // ---- end of chunk: key bytes out, carry the partial data segment to the front
            {
                const uint32_t nk = (chunk_samples + 3) >> 2;
                uint8_t *kg = kdst + (uint64_t)k * (ENC_CH_SAMPLES / 4);
                if ((reinterpret_cast<uintptr_t>(kg) & 3u) == 0) {
                    const uint32_t nw = nk >> 2;
                    const uint32_t *ks = reinterpret_cast<const uint32_t *>(ws.kbuf);
                    uint32_t *kg32 = reinterpret_cast<uint32_t *>(kg);
                    for (uint32_t i = lane; i < nw; i += 32) kg32[i] = ks[i];
                    const uint32_t i = (nw << 2) + lane;
                    if (i < nk) kg[i] = ws.kbuf[i];
                } else {
                    for (uint32_t i = lane; i < nk; i += 32) kg[i] = ws.kbuf[i];
                }
            }
            if (d.fseg) {
                const uint32_t rem = d.pos & 15u;
                uint8_t t = 0;
                if (lane < (int)rem) t = ws.dbuf[d.fseg * 16 + lane];
                __syncwarp();
                if (lane < (int)rem) ws.dbuf[lane] = t;
                d.gbase += d.fseg * 16;
                d.pos = rem;
                d.fseg = 0;
            }
            ++q;
            bytes_cur = bytes_next;
            __syncwarp();  // stage, kbuf and dbuf front are free / visible before the next chunk touches them
        }
        // remaining (< 16) data bytes of the stream leave as byte stores
        if (lane >= (int)d.head && lane < (int)d.pos) d.gbase[lane] = ws.dbuf[lane];
        const uint64_t data_bytes = (uint64_t)((d.gbase + d.pos) - ddst);
        if (lane == 0) {
            a.svb_len[r] = (uint32_t)(4 + nkeys + data_bytes);
            a.status[r] = S5B_OK;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
constexpr int DEC_WARPS = 8;
constexpr int DEC_BLK = 1024;  // bytes per bulk copy
constexpr int DEC_NB = 4;      // ring blocks per warp
constexpr int DEC_RING = DEC_BLK * DEC_NB;

struct __align__(128) DecWarpSmem {
    uint8_t ring[DEC_RING];
    unsigned long long bar[DEC_NB];
};

// zigzag decode (streamvbyte_zigzag.c:23-25): (v >> 1) ^ -(v & 1) == (v >> 1) - (v & 1) * v, which is one
// instruction shorter (LOP, SHF, IMAD)
__device__ __forceinline__ uint32_t zz_dec(uint32_t v) { return (v >> 1) - (v & 1u) * v; }

// Per-read decode state (warp-uniform unless noted).
struct DecState {
    const uint8_t *data16;  // 16-byte aligned global address of ring position 0
    uint64_t lim;           // bulk-copyable bytes from data16
    uint32_t nblk;          // ring blocks the stream spans
    uint32_t issued, waited;
    uint32_t skew;          // data stream starts at ring position skew
    uint32_t D;             // data bytes the stream must consume exactly (slow5_press.c:1130-1136)
    uint32_t pos;           // data bytes consumed
    uint32_t acc;           // running sum, prev = 0 (slow5_press.c:1162)
    uint32_t nkeys;
    uint32_t kk;            // per lane: the two control bytes of the coming iteration
};

__device__ __forceinline__ void dec_issue_block(DecState &s, uint32_t bar0, uint32_t ring0, const int lane) {
    const uint32_t slot = s.issued % DEC_NB;
    const uint64_t b0 = (uint64_t)s.issued * DEC_BLK;
    uint64_t bytes = s.lim - b0;
    if (bytes > DEC_BLK) bytes = DEC_BLK;
    if (lane == 0) {
        mbar_arrive_expect_tx(bar0 + 8 * slot, (uint32_t)bytes);
        bulk_g2s(ring0 + slot * DEC_BLK, s.data16 + b0, (uint32_t)bytes, bar0 + 8 * slot);
    }
    ++s.issued;
}

// One 256-sample iteration; lane owns values 8*lane .. 8*lane+7.  Returns false when the control bytes
// claim more data than the stream holds.
template <bool PARTIAL>
__device__ __forceinline__ bool dec_iteration(DecState &s, const uint8_t *ring, const uint8_t *knext, const bool guard,
                                              const uint32_t next_ki, int16_t *o, const int nvalid, const int lane,
                                              uint32_t &phase_bits, const uint32_t bar0, const uint32_t ring0) {
    const uint32_t k_now = s.kk;
    if (!PARTIAL) {  // prefetch the next iteration's control bytes (the last iteration has no successor)
        if (!guard) {
            s.kk = (uint32_t)__ldg(knext) | ((uint32_t)__ldg(knext + 1) << 8);
        } else {
            s.kk = 0;
            if (next_ki < s.nkeys) s.kk = __ldg(knext);
            if (next_ki + 1 < s.nkeys) s.kk |= (uint32_t)__ldg(knext + 1) << 8;
        }
    }
    uint32_t c[8];
    uint32_t csum = 0, cor = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        c[j] = (k_now >> (2 * j)) & 3u;
        if (PARTIAL && lane * 8 + j >= nvalid) c[j] = 0;
        csum += c[j];
        cor |= c[j];
    }
    uint32_t lane_len = 8 + csum;
    if (PARTIAL) lane_len = min(max(nvalid - lane * 8, 0), 8) + csum;
    // prefix scan over the control-byte lengths -> per-lane data offsets
    const uint32_t incl = warp_incl_scan(lane_len);
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    if (s.pos + total > s.D) return false;  // stream claims more data than it holds: never gather past it
    const uint32_t need = (s.skew + s.pos + total + DEC_BLK - 1) / DEC_BLK;
    while (s.waited < need) {
        const uint32_t slot = s.waited % DEC_NB;
        mbar_wait(bar0 + 8 * slot, (phase_bits >> slot) & 1u);
        phase_bits ^= 1u << slot;
        ++s.waited;
    }
    const uint32_t ri = (s.skew + s.pos + incl - lane_len) & (DEC_RING - 1);
    const bool wrap = __any_sync(FULL, ri + lane_len > DEC_RING);
    const bool wide = __any_sync(FULL, (cor & 2u) != 0);
    uint32_t v[8];
    if (!PARTIAL && !wrap && !wide) {
        // common path: 1- and 2-byte values, no ring wrap inside any lane's run
        const uint8_t *q = ring + ri;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = q[0];
            if (c[j]) v[j] |= (uint32_t)q[1] << 8;
            q += 1 + c[j];
        }
    } else {
        uint32_t p = ri;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = 0;
            if (!PARTIAL || lane * 8 + j < nvalid) {
                v[j] = ring[p & (DEC_RING - 1)];
                if (c[j] >= 1) v[j] |= (uint32_t)ring[(p + 1) & (DEC_RING - 1)] << 8;
                if (c[j] >= 2) v[j] |= (uint32_t)ring[(p + 2) & (DEC_RING - 1)] << 16;
                if (c[j] == 3) v[j] |= (uint32_t)ring[(p + 3) & (DEC_RING - 1)] << 24;
                p += 1 + c[j];
            }
        }
    }
    // zigzag decode + running sum (streamvbyte_zigzag.c:23-25,34-40); mod 2^32 arithmetic, the truncating
    // int16 store keeps the low 16 bits
    uint32_t sum[8];
    uint32_t run = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        run += zz_dec(v[j]);
        sum[j] = run;
    }
    const uint32_t incl_sum = warp_incl_scan(run);
    const uint32_t base = s.acc + incl_sum - run;
    s.acc += __shfl_sync(FULL, incl_sum, 31);
    if (!PARTIAL) {
        uint4 w;
        w.x = __byte_perm(sum[0] + base, sum[1] + base, 0x5410);
        w.y = __byte_perm(sum[2] + base, sum[3] + base, 0x5410);
        w.z = __byte_perm(sum[4] + base, sum[5] + base, 0x5410);
        w.w = __byte_perm(sum[6] + base, sum[7] + base, 0x5410);
        *reinterpret_cast<uint4 *>(o) = w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (lane * 8 + j < nvalid) o[j] = (int16_t)(uint16_t)(sum[j] + base);
    }
    s.pos += total;
    __syncwarp();  // all lanes have finished reading the ring before blocks are recycled
    const uint32_t done_blocks = (s.skew + s.pos) / DEC_BLK;
    while (s.issued < s.nblk && s.issued < done_blocks + DEC_NB) dec_issue_block(s, bar0, ring0, lane);
    return true;
}

__global__ void __launch_bounds__(DEC_WARPS * 32, 5) svbzd_decode_kernel(const SvbDecodeArgs a) {
    __shared__ DecWarpSmem smem[DEC_WARPS];
    const int lane = threadIdx.x & 31;
    DecWarpSmem &ws = smem[threadIdx.x >> 5];
    const uint32_t bar0 = smem_u32(&ws.bar[0]);
    const uint32_t ring0 = smem_u32(&ws.ring[0]);
    if (lane == 0) {
        for (int s = 0; s < DEC_NB; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t phase_bits = 0;  // per ring slot: parity of the next completion to wait for

    for (;;) {
        const uint64_t r = next_work(a.work_counter, lane);
        if (r >= a.n_reads) break;
        const uint64_t ioff = a.svb_off[r];
        const uint32_t ilen = a.svb_len[r];
        const uint8_t *p = a.svb + ioff;
        int32_t st = S5B_OK;
        uint32_t n = 0;
        if (ilen < 4 || ioff + ilen > a.svb_capacity) {
            st = S5B_ERR_ARG;
        } else {
            n = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);  // :1120
        }
        const uint64_t soff = a.sig_off[r];
        const uint64_t scap = a.sig_off[r + 1] - soff;
        const uint32_t nkeys = (uint32_t)(((uint64_t)n + 3) >> 2);
        if (st == S5B_OK) {
            if (soff & 7) st = S5B_ERR_ARG;
            else if (4ull + nkeys > ilen) st = S5B_ERR_PRESS;  // keys alone overrun the stream
            else if (scap < n) st = S5B_ERR_NOSPACE;
        }
        if (st != S5B_OK) {
            if (lane == 0) {
                a.status[r] = st;
                a.n_samples[r] = n;
            }
            continue;
        }
        const uint8_t *keys = p + 4;
        const uint8_t *data = keys + nkeys;
        DecState s;
        s.D = ilen - 4 - nkeys;
        s.nkeys = nkeys;
        s.skew = (uint32_t)(reinterpret_cast<uintptr_t>(data) & 15u);
        s.data16 = data - s.skew;
        // bytes of [data16, ...) that may be bulk-copied: up to the 16-byte granule covering the stream end,
        // never past the slab
        s.lim = ((uint64_t)s.skew + s.D + 15) & ~15ull;
        {
            const uint64_t room = a.svb_capacity - (uint64_t)(s.data16 - a.svb);
            if (s.lim > room) s.lim = room & ~15ull;
        }
        s.nblk = s.D ? (uint32_t)((s.lim + DEC_BLK - 1) / DEC_BLK) : 0;
        s.issued = s.waited = 0;
        s.pos = 0;
        s.acc = 0;
        while (s.issued < s.nblk && s.issued < DEC_NB) dec_issue_block(s, bar0, ring0, lane);

        int16_t *out = a.sig + soff + lane * 8;
        // control bytes of the first iteration
        s.kk = 0;
        {
            const uint32_t ki = 2 * lane;
            if (ki < nkeys) s.kk = __ldg(keys + ki);
            if (ki + 1 < nkeys) s.kk |= (uint32_t)__ldg(keys + ki + 1) << 8;
        }
        const uint32_t full_iters = n >> 8;
        const int tail = (int)(n & 255u);
        const uint8_t *kp = keys + 2 * lane + 64;  // this lane's control bytes of the NEXT iteration
        bool ok = true;
        for (uint32_t it = 0; it < full_iters; ++it) {
            // unguarded prefetch only when the next iteration is a full one as well
            const bool guard = it + 2 > full_iters;
            ok = dec_iteration<false>(s, ws.ring, kp, guard, (it + 1) * 64 + 2 * lane, out, 256, lane, phase_bits, bar0,
                                      ring0);
            if (!ok) break;
            kp += 64;
            out += 256;
        }
        if (ok && tail) ok = dec_iteration<true>(s, ws.ring, kp, true, 0, out, tail, lane, phase_bits, bar0, ring0);
        // drain copies that were issued but never needed (only possible for a malformed stream)
        while (s.waited < s.issued) {
            const uint32_t slot = s.waited % DEC_NB;
            mbar_wait(bar0 + 8 * slot, (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            ++s.waited;
        }
        __syncwarp();
        if (lane == 0) {
            a.n_samples[r] = n;
            a.status[r] = (!ok || s.pos != s.D) ? S5B_ERR_PRESS : S5B_OK;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int svbzd_encode_blocks_per_sm() {
    int n = 0;
    if (cudaFuncSetAttribute(svbzd_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(EncWarpSmem) * ENC_WARPS)) != cudaSuccess)
        return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, svbzd_encode_kernel, ENC_WARPS * 32,
                                                      sizeof(EncWarpSmem) * ENC_WARPS) != cudaSuccess)
        return 0;
    return n;
}
int svbzd_decode_blocks_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, svbzd_decode_kernel, DEC_WARPS * 32, 0) != cudaSuccess)
        return 0;
    return n;
}
cudaError_t launch_svbzd_encode(const SvbEncodeArgs &a, unsigned grid, cudaStream_t st) {
    svbzd_encode_kernel<<<grid, ENC_WARPS * 32, sizeof(EncWarpSmem) * ENC_WARPS, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_svbzd_decode(const SvbDecodeArgs &a, unsigned grid, cudaStream_t st) {
    svbzd_decode_kernel<<<grid, DEC_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}
int enc_warps() { return ENC_WARPS; }
int dec_warps() { return DEC_WARPS; }

}  // namespace v3
}  // namespace s5b
