// svbzd_kernels.cu -- sm_100a kernels for the svb-zd signal codec (StreamVByte "1234" coding of
// zigzag-delta values, u32 sample-count header), one warp per read.
//
// Replaces, for whole batches, the reference CPU routines
//   ptr_compress_svb_zd / ptr_compress_svb      slow5lib/src/slow5_press.c:1082-1115 / :1062-1079
//   ptr_depress_svb_zd  / ptr_depress_svb       slow5lib/src/slow5_press.c:1143-1173 / :1118-1140
//   __slow5_zigzag_delta_encode / _decode       thirdparty/streamvbyte/src/streamvbyte_zigzag.c:15-40
//   __slow5_streamvbyte_encode / _decode        thirdparty/streamvbyte/src/streamvbyte_{en,de}code.c
// Output bytes are identical to the reference's (tests/test_svbzd_gpu.py checks against the oracle).
//
// Data movement (both kernels are HBM-streaming, integer-only, no tensor cores; "v4"):
//   * each warp owns one read at a time (dynamic work counter) and loops over 256-sample iterations;
//   * encode: the int16 signal is staged HBM->smem by 1-D bulk async copies (TMA engine, UBLKCP.S.G)
//     into a 2-stage per-warp pipeline guarded by mbarriers; a lane owns two quads of values
//     (4*lane.. and 128+4*lane..) so that the byte scatter into shared memory is nearly free of bank
//     conflicts; one warp prefix scan over both quads' byte counts gives every lane its data offsets;
//     the data stream leaves once per 1024-sample chunk as one bulk shared->global copy (UBLKCP.G.S),
//     the key bytes as two word stores per lane; svbzd_encode_bytes_kernel is the same body for samples at any byte
//     alignment (inside packed records): staging starts at the 16-byte granule below, lanes realign with funnel shifts;
//   * decode: the variable-length data stream is staged by bulk async copies into a 4 x 1 KiB
//     per-warp ring (mirrored at its end so word loads never wrap); a warp prefix scan over the
//     control-byte lengths resolves the per-lane data offsets; values are widened pairwise with PRMT,
//     zigzag-decoded and prefix-summed two 16-bit lanes per register; a second scan carries the running
//     sum across lanes; samples leave as 128-bit coalesced stores.
#include <cstdlib>
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "../../include/slow5b200.h"

namespace s5b {

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
// Lane mapping of one 256-sample iteration: lane owns values 4*lane .. 4*lane+3 ("quad A") and
// 128+4*lane .. 128+4*lane+3 ("quad B").  Neighbouring lanes of one byte-store instruction are then
// ~4.1 bytes apart (a warp's 32 addresses span ~33 words: no shared-memory bank conflicts), where
// 8 consecutive values per lane put them ~8.3 bytes apart (66 words, 2 wavefronts per store).
#ifndef S5B_ENC_MIN_CTAS
#define S5B_ENC_MIN_CTAS 4
#endif
#ifndef S5B_DEC_MIN_CTAS
#define S5B_DEC_MIN_CTAS 5
#endif
#ifndef S5B_ENC_NDBUF
#define S5B_ENC_NDBUF 1  // 1: single data buffer, every drain waits for its own bulk copy (smaller footprint, more CTAs)
#endif
constexpr int ENC_WARPS = 8;
constexpr int ENC_CH_SAMPLES = 1024;  // samples per bulk-copy chunk (4 iterations of 256)
constexpr int ENC_CH_BYTES = ENC_CH_SAMPLES * 2;
constexpr int ENC_STAGES = 2;
// Data bytes of one chunk between two drains.  Iteration i of a chunk starts with at most 15 + 512*i
// bytes buffered (an iteration without 3-byte codes appends <= 512; one with them drains right away),
// and appends at most 768: 15 + 3*512 + 768 = 2319.
constexpr int ENC_DB = 2320;

constexpr int ENC_STAGE_BYTES = ENC_CH_BYTES + 32;  // a chunk of a source that starts up to 15 bytes into its 16-byte granule, and the
                                                    // word behind it that a realigning load touches
struct __align__(128) EncWarpSmem {
    uint8_t in[ENC_STAGES][ENC_STAGE_BYTES];  // staged signal
    uint8_t dbuf[S5B_ENC_NDBUF][ENC_DB];   // data-stream double buffer; dbuf[x][0] <-> 16-byte aligned global address
    uint8_t kbuf[ENC_CH_SAMPLES / 4];      // key bytes of one chunk, natural index
    unsigned long long bar[ENC_STAGES];
};
static_assert(ENC_DB % 16 == 0, "both halves of the double buffer must stay 16-byte aligned");

// Per-warp output state of the data stream (all members warp-uniform).  Data bytes are appended linearly
// to the current half of the double buffer; a drain sends its complete 16-byte segments to global memory
// with one bulk shared->global copy (async proxy, no LDS/STG by the lanes) and moves the (< 16 byte)
// remainder to the front of the other half, which becomes current.
struct EncData {
    uint8_t *gbase;  // 16-byte aligned global address of buf[0]
    uint8_t *buf;    // current half
    uint8_t *oth;    // other half
    uint32_t pos;    // bytes appended (index into buf)
    uint32_t head;   // first valid byte of segment 0 (stream start not 16-byte aligned), else 0
};

__device__ __forceinline__ void enc_drain(EncData &d, const int lane) {
    const uint32_t nseg = d.pos >> 4;
    uint32_t first = 0;
    if (d.head && nseg) {  // ragged stream start: segment 0 leaves as byte stores, never touching the bytes before it
        if (lane >= (int)d.head && lane < 16) d.gbase[lane] = d.buf[lane];
        d.head = 0;
        first = 1;
    }
    fence_proxy_async_smem();  // every lane's byte stores become visible to the async proxy
    __syncwarp();
    if (lane == 0) {
        if (nseg > first) bulk_s2g(d.gbase + first * 16, smem_u32(d.buf) + first * 16, (nseg - first) * 16);
        bulk_commit();       // (possibly empty) group, so "all but the newest" below always covers the other half
        bulk_wait_read<S5B_ENC_NDBUF - 1>();  // the previous drain has finished reading the other half
    }
    __syncwarp();
    const uint32_t rem = d.pos & 15u;
#if S5B_ENC_NDBUF == 2
    if (lane < (int)rem) d.oth[lane] = d.buf[nseg * 16 + lane];
    uint8_t *t = d.buf;
    d.buf = d.oth;
    d.oth = t;
#else
    uint8_t t = 0;
    if (lane < (int)rem) t = d.buf[nseg * 16 + lane];
    __syncwarp();
    if (lane < (int)rem) d.buf[lane] = t;
#endif
    d.gbase += nseg * 16;
    d.pos = rem;
}

// min(t, 1) kept opaque so the 1-bit code stays an integer (the compiler otherwise turns it into a predicate and
// every later use into a compare/select pair)
__device__ __forceinline__ uint32_t min1(uint32_t t) {
    uint32_t c;
    asm("min.u32 %0, %1, 1;" : "=r"(c) : "r"(t));
    return c;
}

// zigzag, streamvbyte_zigzag.c:4-6
__device__ __forceinline__ uint32_t zz_enc(int dd) { return ((uint32_t)dd << 1) ^ (uint32_t)(dd >> 31); }

// Second half of an iteration: codes, keys, prefix scan, byte assembly.  WIDE (3-byte codes present in
// the warp, |delta| >= 32768) is a separate instantiation so the common path carries 1-bit codes only.
template <bool PARTIAL, bool WIDE>
__device__ __forceinline__ void enc_emit(const uint32_t (&za)[4], const uint32_t (&zb)[4], const int lane,
                                         const int nvalid, EncData &d, uint8_t *kslot) {
    // svb code = bytes - 1 (streamvbyte_encode.c:31-54; code 3 is unreachable from int16 input: z < 2^17)
    uint32_t ta[4], tb[4], ca[4], cb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ta[j] = za[j] >> 8;
        tb[j] = zb[j] >> 8;
        ca[j] = WIDE ? (za[j] > 0xFFu) + (za[j] > 0xFFFFu) : min1(ta[j]);
        cb[j] = WIDE ? (zb[j] > 0xFFu) + (zb[j] > 0xFFFFu) : min1(tb[j]);
    }
    // 2 bits per value, value i -> key byte i/4, shift 2*(i%4) (streamvbyte_encode.c:56-79)
    const uint32_t key_a = ((ca[3] * 4 + ca[2]) * 4 + ca[1]) * 4 + ca[0];
    const uint32_t key_b = ((cb[3] * 4 + cb[2]) * 4 + cb[1]) * 4 + cb[0];
    uint32_t len_a = 4 + ca[0] + ca[1] + ca[2] + ca[3];
    uint32_t len_b = 4 + cb[0] + cb[1] + cb[2] + cb[3];
    if (PARTIAL) {  // invalid samples: z = 0, c = 0, no byte
        len_a += (uint32_t)min(max(nvalid - lane * 4, 0), 4) - 4;
        len_b += (uint32_t)min(max(nvalid - 128 - lane * 4, 0), 4) - 4;
    }
    // one prefix scan over both quads' byte counts (16 bits each; a warp appends <= 32*12 per quad row)
    const uint32_t lens = len_a | (len_b << 16);
    const uint32_t incl = warp_incl_scan(lens);
    const uint32_t tot = __shfl_sync(FULL, incl, 31);
    const uint32_t excl = incl - lens;
    const uint32_t tot_a = tot & 0xFFFFu;
    uint8_t *pa = d.buf + d.pos + (excl & 0xFFFFu);
    uint8_t *pb = d.buf + d.pos + tot_a + (excl >> 16);
    if (!PARTIAL) {
        // Both bytes are stored unconditionally except for the quad's last value: a spurious high byte
        // lands where the same lane's next value puts its low byte afterwards (program order), so it
        // never survives.  The last value must not spill into the next lane's first byte.
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            pa[0] = (uint8_t)za[j];
            if (j < 3 || ca[j]) pa[1] = (uint8_t)ta[j];
            if (WIDE && ca[j] == 2) pa[2] = (uint8_t)(za[j] >> 16);
            pa += 1 + ca[j];
            pb[0] = (uint8_t)zb[j];
            if (j < 3 || cb[j]) pb[1] = (uint8_t)tb[j];
            if (WIDE && cb[j] == 2) pb[2] = (uint8_t)(zb[j] >> 16);
            pb += 1 + cb[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (lane * 4 + j < nvalid) {
                pa[0] = (uint8_t)za[j];
                if (ca[j]) pa[1] = (uint8_t)ta[j];
                if (WIDE && ca[j] == 2) pa[2] = (uint8_t)(za[j] >> 16);
                pa += 1 + ca[j];
            }
            if (128 + lane * 4 + j < nvalid) {
                pb[0] = (uint8_t)zb[j];
                if (cb[j]) pb[1] = (uint8_t)tb[j];
                if (WIDE && cb[j] == 2) pb[2] = (uint8_t)(zb[j] >> 16);
                pb += 1 + cb[j];
            }
        }
    }
    kslot[lane] = (uint8_t)key_a;  // invalid samples carry code 0, so padding bits are zero
    kslot[32 + lane] = (uint8_t)key_b;
    d.pos += tot_a + (tot >> 16);
}

// One 256-sample iteration.  wa / wb: the lane's quads (4 packed int16 each); carry (meaningful in lane 0):
// the sample before this iteration's first one; rot_src = (lane + 31) & 31.
template <bool PARTIAL>
__device__ __forceinline__ void enc_iteration(const uint2 wa, const uint2 wb, int &carry, const int lane,
                                              const int rot_src, const int nvalid, EncData &d, uint8_t *kslot) {
    // widen (slow5_press.c:1095-1097): one PRMT (sign-replicating) / one arithmetic shift per sample
    int xa[4], xb[4];
    xa[0] = sext_lo16(wa.x);
    xa[1] = (int)wa.x >> 16;
    xa[2] = sext_lo16(wa.y);
    xa[3] = (int)wa.y >> 16;
    xb[0] = sext_lo16(wb.x);
    xb[1] = (int)wb.x >> 16;
    xb[2] = sext_lo16(wb.y);
    xb[3] = (int)wb.y >> 16;
    // previous samples: one rotate-by-one shuffle of (last of quad A, last of quad B).  Lane 0 receives lane
    // 31's pair: its quad-A last sample precedes lane 0's quad B, its quad-B last sample is the next carry.
    const uint32_t lasts = __byte_perm(wa.y, wb.y, 0x7632);
    const uint32_t rot = __shfl_sync(FULL, lasts, rot_src);
    const int r_lo = sext_lo16(rot), r_hi = (int)rot >> 16;
    int pa = lane ? r_lo : carry;
    int pb = lane ? r_hi : r_lo;
    carry = r_hi;

    uint32_t za[4], zb[4], zor = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // zigzag-delta, streamvbyte_zigzag.c:15-20
        za[j] = zz_enc(xa[j] - pa);
        pa = xa[j];
        zb[j] = zz_enc(xb[j] - pb);
        pb = xb[j];
        if (PARTIAL && lane * 4 + j >= nvalid) za[j] = 0;
        if (PARTIAL && 128 + lane * 4 + j >= nvalid) zb[j] = 0;
        zor |= za[j] | zb[j];
    }
    if (__any_sync(FULL, zor > 0xFFFFu)) {
        enc_emit<PARTIAL, true>(za, zb, lane, nvalid, d, kslot);
        __syncwarp();
        enc_drain(d, lane);  // keeps the buffered bytes inside the ENC_DB envelope (rare path)
    } else {
        enc_emit<PARTIAL, false>(za, zb, lane, nvalid, d, kslot);
    }
}

// the lane's quad idx (8 bytes) of a staged chunk whose first sample sits `mis` bytes into the stage (BYTE_SRC), else aligned
template <bool BYTE_SRC>
__device__ __forceinline__ uint2 enc_load_quad(const uint8_t *stage, uint32_t mis, uint32_t idx) {
    if (!BYTE_SRC) return reinterpret_cast<const uint2 *>(stage)[idx];
    const uint32_t bo = mis + idx * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(stage) + (bo >> 2);
    const uint32_t sh = (bo & 3u) * 8u;
    const uint32_t a0 = w[0], a1 = w[1], a2 = w[2];
    return make_uint2(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh));
}

// BYTE_SRC: the samples are read where they lie, at any byte alignment (a.src_bytes / a.src_byte_off); staging starts at the
// 16-byte granule below the first sample and the lanes realign their quads with funnel shifts (two more instructions per quad)
template <bool BYTE_SRC>
__device__ __forceinline__ void svbzd_encode_body(const SvbEncodeArgs &a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    EncWarpSmem *smem = reinterpret_cast<EncWarpSmem *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int rot_src = (lane + 31) & 31;
    EncWarpSmem &ws = smem[threadIdx.x >> 5];
    const uint32_t bar0 = smem_u32(&ws.bar[0]);
    const uint32_t in0 = smem_u32(&ws.in[0][0]);
    if (lane == 0) {
        for (int s = 0; s < ENC_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t q = 0;  // chunks consumed by this warp so far (stage = q & 1, parity = (q >> 1) & 1)
    EncData d;
    d.buf = ws.dbuf[0];
    d.oth = ws.dbuf[S5B_ENC_NDBUF - 1];

    for (;;) {
        const uint64_t r = next_work(a.work_counter, lane);
        if (r >= a.n_reads) break;
        const uint32_t n = a.n_samples[r];
        const uint64_t ooff = a.svb_off[r];
        const uint64_t ocap = a.svb_off[r + 1] - ooff;
        const uint32_t nkeys = (n + 3) >> 2;
        int32_t st = S5B_OK;
        const uint8_t *src;        // 16-byte aligned start of the staged bytes
        uint32_t mis = 0;          // bytes from there to the first sample
        uint64_t slot_bytes16;     // bytes that may be bulk-copied from src: whole 16-byte granules inside the slab
        if (!BYTE_SRC) {
            const uint64_t soff = a.sig_off[r];
            const uint64_t scap = a.sig_off[r + 1] - soff;
            if ((soff & 7) || scap < n) st = S5B_ERR_ARG;
            src = reinterpret_cast<const uint8_t *>(a.sig + soff);
            // (the slot is a multiple of 8 samples except possibly the last one of the slab)
            slot_bytes16 = (scap * 2) & ~15ull;
        } else {
            const uint64_t boff = a.src_byte_off[r];
            if (boff + 2ull * n > a.src_capacity) st = S5B_ERR_ARG;
            const uint8_t *first = a.src_bytes + boff;
            mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u);
            src = first - mis;
            slot_bytes16 = st == S5B_OK ? (a.src_capacity - (boff - mis)) & ~15ull : 0;
        }
        if (st == S5B_OK && ocap < 4ull + nkeys + 3ull * n) st = S5B_ERR_NOSPACE;
        if (st != S5B_OK) {
            if (lane == 0) {
                a.status[r] = st;
                a.svb_len[r] = 0;
            }
            continue;
        }
        uint8_t *dst = a.svb + ooff;
        const uint64_t n_bytes = (uint64_t)mis + (uint64_t)n * 2;  // staged bytes that hold the read

        if (lane < 4) dst[lane] = (uint8_t)(n >> (8 * lane));  // u32 LE header, slow5_press.c:1074
        uint8_t *kdst = dst + 4;
        uint8_t *ddst = kdst + nkeys;
        d.head = d.pos = (uint32_t)(reinterpret_cast<uintptr_t>(ddst) & 15u);
        d.gbase = ddst - d.head;

        const uint32_t nchunks = (n + ENC_CH_SAMPLES - 1) / ENC_CH_SAMPLES;
        // Bulk-copyable bytes of the read.  Stage byte 0 of chunk k is src + k * ENC_CH_BYTES; a chunk whose samples start `mis`
        // bytes into the stage needs one more granule behind its ENC_CH_BYTES (the next chunk copies that granule again).
        uint64_t copyable = (n_bytes + 15) & ~15ull;
        if (copyable > slot_bytes16) copyable = slot_bytes16;
        // split into whole chunks and one last piece (32-bit from here on)
        const uint32_t copy_full = (uint32_t)(copyable / ENC_CH_BYTES);
        const uint32_t copy_last = (uint32_t)(copyable % ENC_CH_BYTES);
        auto issue = [&](uint32_t k, uint32_t qq) {
            // chunk k of this read -> stage qq & 1
            uint32_t bytes = k < copy_full ? (uint32_t)ENC_CH_BYTES : (k == copy_full ? copy_last : 0u);
            if (BYTE_SRC && mis && k < copy_full && (k + 1 < copy_full || copy_last)) bytes += 16;
            const uint32_t stage = qq & 1;
            if (lane == 0) {
                if (bytes) {
                    mbar_arrive_expect_tx(bar0 + 8 * stage, bytes);
                    bulk_g2s(in0 + stage * ENC_STAGE_BYTES, src + (size_t)k * ENC_CH_BYTES, bytes, bar0 + 8 * stage);
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
            if (bytes_cur < mis + chunk_samples * 2) {
                // ragged end of the slab: the bulk copy could not take the last (< 16 byte) piece
                if (!BYTE_SRC) {
                    const int16_t *g = reinterpret_cast<const int16_t *>(src) + (uint64_t)k * ENC_CH_SAMPLES;
                    int16_t *s = reinterpret_cast<int16_t *>(ws.in[stage]);
                    for (uint32_t i = bytes_cur / 2 + lane; i < chunk_samples; i += 32) s[i] = g[i];
                } else {
                    const uint8_t *g = src + (uint64_t)k * ENC_CH_BYTES;
                    for (uint32_t i = bytes_cur + lane; i < mis + chunk_samples * 2; i += 32) ws.in[stage][i] = g[i];
                }
                __syncwarp();
            }
            const uint8_t *in2 = ws.in[stage];
            const uint32_t full_iters = chunk_samples >> 8;
            for (uint32_t it = 0; it < full_iters; ++it)
                enc_iteration<false>(enc_load_quad<BYTE_SRC>(in2, mis, it * 64 + lane),
                                     enc_load_quad<BYTE_SRC>(in2, mis, it * 64 + 32 + lane), carry, lane, rot_src, 256, d,
                                     ws.kbuf + it * 64);
            const int tail = chunk_samples & 255;
            if (tail)
                enc_iteration<true>(enc_load_quad<BYTE_SRC>(in2, mis, full_iters * 64 + lane),
                                    enc_load_quad<BYTE_SRC>(in2, mis, full_iters * 64 + 32 + lane), carry, lane, rot_src,
                                    tail, d, ws.kbuf + full_iters * 64);
            __syncwarp();
            // ---- end of chunk: key bytes out, data segments out
            {
                const uint32_t nk = (chunk_samples + 3) >> 2;
                uint8_t *kg = kdst + (size_t)k * (ENC_CH_SAMPLES / 4);
                const uint32_t *ks = reinterpret_cast<const uint32_t *>(ws.kbuf);
                if (nk == ENC_CH_SAMPLES / 4 && (reinterpret_cast<uintptr_t>(kg) & 3u) == 0) {
                    uint32_t *kg32 = reinterpret_cast<uint32_t *>(kg);  // whole chunk, word aligned: two words per lane
                    kg32[lane] = ks[lane];
                    kg32[lane + 32] = ks[lane + 32];
                } else {
#pragma unroll 1
                    for (uint32_t i = lane; i < nk; i += 32) kg[i] = ws.kbuf[i];
                }
            }
            enc_drain(d, lane);
            ++q;
            bytes_cur = bytes_next;
            __syncwarp();  // stage, kbuf and the buffer fronts are free / visible before the next chunk touches them
        }
        // remaining (< 16) data bytes of the stream leave as byte stores
        if (lane >= (int)d.head && lane < (int)d.pos) d.gbase[lane] = d.buf[lane];
        const uint64_t data_bytes = (uint64_t)((d.gbase + d.pos) - ddst);
        if (lane == 0) {
            a.svb_len[r] = (uint32_t)(4 + nkeys + data_bytes);
            a.status[r] = S5B_OK;
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_read<0>();  // shared memory must outlive the last bulk copy's reads
}

__global__ void __launch_bounds__(ENC_WARPS * 32, S5B_ENC_MIN_CTAS) svbzd_encode_kernel(const SvbEncodeArgs a) {
    svbzd_encode_body<false>(a);
}
// the samples where they lie inside packed records (the transcoder's encode pass: no copy to an aligned slab first)
__global__ void __launch_bounds__(ENC_WARPS * 32, S5B_ENC_MIN_CTAS) svbzd_encode_bytes_kernel(const SvbEncodeArgs a) {
    svbzd_encode_body<true>(a);
}

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
constexpr int DEC_WARPS = 8;
constexpr int DEC_BLK = 1024;  // bytes per bulk copy
constexpr int DEC_NB = 4;      // ring blocks per warp
constexpr int DEC_RING = DEC_BLK * DEC_NB;
constexpr int DEC_MIRROR = 32;  // the ring's first bytes repeated behind its end: word loads never wrap

struct __align__(128) DecWarpSmem {
    uint8_t ring[DEC_RING + DEC_MIRROR];
    unsigned long long bar[DEC_NB];
};

// zigzag decode (streamvbyte_zigzag.c:23-25)
__device__ __forceinline__ uint32_t zz_dec(uint32_t v) { return (v >> 1) ^ (0u - (v & 1u)); }

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

// wait until ring blocks [waited, need) have landed; a block landing in slot 0 refreshes the mirror
__device__ __forceinline__ void dec_wait_blocks(DecState &s, const uint32_t need, uint8_t *ring, uint32_t &phase_bits,
                                                const uint32_t bar0, const int lane) {
    while (s.waited < need) {
        const uint32_t slot = s.waited % DEC_NB;
        mbar_wait(bar0 + 8 * slot, (phase_bits >> slot) & 1u);
        phase_bits ^= 1u << slot;
        ++s.waited;
        if (slot == 0) {
            uint32_t *r32 = reinterpret_cast<uint32_t *>(ring);
            if (lane < DEC_MIRROR / 4) r32[DEC_RING / 4 + lane] = r32[lane];
            __syncwarp();
        }
    }
}

// raw prmt: the __byte_perm intrinsic would spend an extra AND on masking every run-time selector
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// Widening table of the common path, one entry per control byte whose four values are 1 or 2 bytes long (flag of value j at
// bit 2j).  The four values sit in an 8-byte window (a, b):
//   x, y  PRMT selectors that pull values (0, 1) / (2, 3) out of the window as packed pairs, two bytes each (the byte after a
//         1-byte value comes along and is masked away below);
//   z, w  per pair, 0x7fff in every half whose value has 2 bytes and 0x007f otherwise: the AND mask of the zigzag decode's
//         ">> 1", which is where the stray byte disappears at no cost.
__device__ __forceinline__ uint4 dec_lut_entry(uint32_t k) {
    const uint32_t f0 = k & 1u, f1 = (k >> 2) & 1u, f2 = (k >> 4) & 1u, f3 = (k >> 6) & 1u;
    const uint32_t o1 = 1 + f0, o2 = o1 + 1 + f1, o3 = o2 + 1 + f2;
    uint4 e;
    e.x = 0u | (1u << 4) | (o1 << 8) | ((o1 + 1) << 12);
    e.y = o2 | ((o2 + 1) << 4) | (o3 << 8) | (((o3 + 1) & 7u) << 12);
    e.z = (f0 ? 0x7fffu : 0x007fu) | ((f1 ? 0x7fffu : 0x007fu) << 16);
    e.w = (f2 ? 0x7fffu : 0x007fu) | ((f3 ? 0x7fffu : 0x007fu) << 16);
    return e;
}
// packed zigzag decode of two 16-bit values (exact mod 2^16, which is all the truncating int16 store keeps); `keep` as above
__device__ __forceinline__ uint32_t zz_dec2(uint32_t p, uint32_t keep) {
    const uint32_t m = (p & 0x00010001u) * 0xFFFFu;  // 0xFFFF in every half whose value is odd
    return ((p >> 1) & keep) ^ m;
}

// inclusive warp scan of two independent 16-bit lanes per register (sums mod 2^16)
__device__ __forceinline__ uint32_t warp_incl_scan_16x2(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t;
        int p;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "shfl.sync.up.b32 %0|p, %2, %3, 0, 0xffffffff;\n\t"
            "selp.s32 %1, 1, 0, p;\n\t}"
            : "=r"(t), "=r"(p)
            : "r"(v), "r"(d));
        if (p) v = __vadd2(v, t);  // (folds into a predicated VIADD.16x2)
    }
    return v;
}

// Eight values of one lane, all 1 or 2 bytes long: the lane's 8..16 data bytes start at ring position ri; k0 / k1 are its two
// control bytes.  Returns the lane-local inclusive prefix sums of the zigzag-decoded values, two per register (mod 2^16).
__device__ __forceinline__ void dec_unpack8(const uint8_t *ring, const uint4 *lut, const uint32_t ri, const uint32_t k0,
                                            const uint32_t k1, const uint32_t sh4, uint32_t (&s)[4]) {
    // realigned into four registers (five aligned word loads + funnel shifts)
    const uint32_t *w = reinterpret_cast<const uint32_t *>(ring) + (ri >> 2);
    const uint32_t sh = (ri & 3u) * 8;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
    const uint4 e0 = lut[k0], e1 = lut[k1];
    const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh);
    const uint32_t a2 = __funnelshift_r(w2, w3, sh), a3 = __funnelshift_r(w3, w4, sh);
    // values 4..7 start at byte 4 + (2-byte values among 0..3); sh4 = 8 x that count
    const uint32_t c0 = __funnelshift_rc(a1, a2, sh4), c1 = __funnelshift_rc(a2, a3, sh4);
    // zigzag decode + running sum, two 16-bit lanes per register (streamvbyte_zigzag.c:23-25,34-40)
    // x * 0x10001 = (lo, lo + hi): the pair's own prefix; the carry-in from the left is added per half
    s[0] = zz_dec2(prmt(a0, a1, e0.x), e0.z) * 0x10001u;  // (d0, d0 + d1)
    s[1] = __vadd2(zz_dec2(prmt(a0, a1, e0.y), e0.w) * 0x10001u, __byte_perm(s[0], 0u, 0x3232));
    s[2] = __vadd2(zz_dec2(prmt(c0, c1, e1.x), e1.z) * 0x10001u, __byte_perm(s[1], 0u, 0x3232));
    s[3] = __vadd2(zz_dec2(prmt(c0, c1, e1.y), e1.w) * 0x10001u, __byte_perm(s[2], 0u, 0x3232));
}

// after an iteration: every lane is done with the ring bytes it consumed, their blocks can be refilled
__device__ __forceinline__ void dec_recycle(DecState &s, const uint32_t bar0, const uint32_t ring0, const int lane) {
    __syncwarp();
    const uint32_t done_blocks = (s.skew + s.pos) / DEC_BLK;
    while (s.issued < s.nblk && s.issued < done_blocks + DEC_NB) dec_issue_block(s, bar0, ring0, lane);
}

// 512 samples, the common path twice over ("A": values 8*lane.. of the first 256, "B": the same of the second 256): one scan
// over both halves' data lengths, one over both halves' sums.  ka / kb: the lane's control bytes (k0 | k1 << 8) of A / B, all
// 2-bit codes <= 1.  Returns false when the control bytes claim more data than the stream holds.
__device__ __forceinline__ bool dec_iteration512(DecState &s, uint8_t *ring, const uint4 *lut, const uint32_t ka,
                                                 const uint32_t kb, int16_t *o, const int lane, uint32_t &phase_bits,
                                                 const uint32_t bar0, const uint32_t ring0) {
    const uint32_t ka0 = ka & 0xffu, ka1 = ka >> 8, kb0 = kb & 0xffu, kb1 = kb >> 8;
    const uint32_t pa0 = __popc(ka0), pb0 = __popc(kb0);
    const uint32_t len_a = 8 + pa0 + __popc(ka1), len_b = 8 + pb0 + __popc(kb1);
    const uint32_t lens = len_a | (len_b << 16);                // a warp's half is <= 512 bytes
    const uint32_t incl = warp_incl_scan(lens);                  // prefix scan over the control-byte lengths -> data offsets
    const uint32_t tot = __shfl_sync(FULL, incl, 31);
    const uint32_t tot_a = tot & 0xffffu, total = tot_a + (tot >> 16);
    if (s.pos + total > s.D) return false;  // stream claims more data than it holds: never gather past it
    dec_wait_blocks(s, (s.skew + s.pos + total + DEC_BLK - 1) / DEC_BLK, ring, phase_bits, bar0, lane);
    const uint32_t excl = incl - lens;
    const uint32_t at = s.skew + s.pos;
    uint32_t sa[4], sb[4];
    dec_unpack8(ring, lut, (at + (excl & 0xffffu)) & (DEC_RING - 1), ka0, ka1, pa0 * 8, sa);
    dec_unpack8(ring, lut, (at + tot_a + (excl >> 16)) & (DEC_RING - 1), kb0, kb1, pb0 * 8, sb);
    // the lanes' totals (mod 2^16), scanned for both halves at once
    const uint32_t runs = __byte_perm(sa[3], sb[3], 0x7632);
    const uint32_t incl_sum = warp_incl_scan_16x2(runs);
    const uint32_t tots = __shfl_sync(FULL, incl_sum, 31);
    // carry-in: A starts from the running sum, B from the running sum plus all of A
    const uint32_t carry = (s.acc & 0xffffu) | ((s.acc + tots) << 16);
    const uint32_t base = __vadd2(__vsub2(incl_sum, runs), carry);
    s.acc += (tots & 0xffffu) + (tots >> 16);
    const uint32_t base_a = __byte_perm(base, 0u, 0x1010), base_b = __byte_perm(base, 0u, 0x3232);
    uint4 va, vb;
    va.x = __vadd2(sa[0], base_a);
    va.y = __vadd2(sa[1], base_a);
    va.z = __vadd2(sa[2], base_a);
    va.w = __vadd2(sa[3], base_a);
    vb.x = __vadd2(sb[0], base_b);
    vb.y = __vadd2(sb[1], base_b);
    vb.z = __vadd2(sb[2], base_b);
    vb.w = __vadd2(sb[3], base_b);
    *reinterpret_cast<uint4 *>(o) = va;
    *reinterpret_cast<uint4 *>(o + 256) = vb;
    s.pos += total;
    dec_recycle(s, bar0, ring0, lane);
    return true;
}

// One 256-sample iteration; lane owns values 8*lane .. 8*lane+7 with control bytes k_now (k0 | k1 << 8).  Returns false when
// the control bytes claim more data than the stream holds.
template <bool PARTIAL>
__device__ __forceinline__ bool dec_iteration(DecState &s, uint8_t *ring, const uint4 *lut, const uint32_t k_now, int16_t *o,
                                              const int nvalid, const int lane, uint32_t &phase_bits, const uint32_t bar0,
                                              const uint32_t ring0) {
    const bool wide = __any_sync(FULL, (k_now & 0xAAAAu) != 0);
    if (!PARTIAL && !wide) {
        // ---- common path: every value of the iteration is 1 or 2 bytes
        const uint32_t k0 = k_now & 0xffu, k1 = k_now >> 8;
        const uint32_t p0 = __popc(k0);
        const uint32_t lane_len = 8 + p0 + __popc(k1);
        const uint32_t incl = warp_incl_scan(lane_len);
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        if (s.pos + total > s.D) return false;
        dec_wait_blocks(s, (s.skew + s.pos + total + DEC_BLK - 1) / DEC_BLK, ring, phase_bits, bar0, lane);
        uint32_t sv[4];
        dec_unpack8(ring, lut, (s.skew + s.pos + incl - lane_len) & (DEC_RING - 1), k0, k1, p0 * 8, sv);
        const uint32_t run = sv[3] >> 16;  // the lane's total (mod 2^16)
        const uint32_t incl_sum = warp_incl_scan(run);
        const uint32_t base = (s.acc + incl_sum - run) & 0xFFFFu;
        s.acc += __shfl_sync(FULL, incl_sum, 31);
        const uint32_t base2 = base * 0x10001u;
        uint4 wv;
        wv.x = __vadd2(sv[0], base2);
        wv.y = __vadd2(sv[1], base2);
        wv.z = __vadd2(sv[2], base2);
        wv.w = __vadd2(sv[3], base2);
        *reinterpret_cast<uint4 *>(o) = wv;
        s.pos += total;
    } else {
        uint32_t c[8];
        uint32_t csum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = (k_now >> (2 * j)) & 3u;
            if (PARTIAL && lane * 8 + j >= nvalid) c[j] = 0;
            csum += c[j];
        }
        uint32_t lane_len = 8 + csum;
        if (PARTIAL) lane_len = min(max(nvalid - lane * 8, 0), 8) + csum;
        const uint32_t incl = warp_incl_scan(lane_len);
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        if (s.pos + total > s.D) return false;
        dec_wait_blocks(s, (s.skew + s.pos + total + DEC_BLK - 1) / DEC_BLK, ring, phase_bits, bar0, lane);
        uint32_t p = (s.skew + s.pos + incl - lane_len) & (DEC_RING - 1);
        uint32_t v[8];
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
            uint4 wv;
            wv.x = __byte_perm(sum[0] + base, sum[1] + base, 0x5410);
            wv.y = __byte_perm(sum[2] + base, sum[3] + base, 0x5410);
            wv.z = __byte_perm(sum[4] + base, sum[5] + base, 0x5410);
            wv.w = __byte_perm(sum[6] + base, sum[7] + base, 0x5410);
            *reinterpret_cast<uint4 *>(o) = wv;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (lane * 8 + j < nvalid) o[j] = (int16_t)(uint16_t)(sum[j] + base);
        }
        s.pos += total;
    }
    dec_recycle(s, bar0, ring0, lane);
    return true;
}

// the lane's two control bytes of a 256-sample group (k0 | k1 << 8): kp points at the first of them, ki is its index in the
// control bytes; positions past the last control byte read as 0
__device__ __forceinline__ uint32_t dec_load_keys(const uint8_t *kp, const uint32_t ki, const uint32_t nkeys, const bool guard) {
    uint32_t k0 = 0, k1 = 0;
    if (!guard) {
        k0 = __ldg(kp);
        k1 = __ldg(kp + 1);
    } else {
        if (ki < nkeys) k0 = __ldg(kp);
        if (ki + 1 < nkeys) k1 = __ldg(kp + 1);
    }
    return k0 | (k1 << 8);
}

__global__ void __launch_bounds__(DEC_WARPS * 32, S5B_DEC_MIN_CTAS) svbzd_decode_kernel(const SvbDecodeArgs a) {
    __shared__ DecWarpSmem smem[DEC_WARPS];
    __shared__ uint4 lut[256];
    const int lane = threadIdx.x & 31;
    DecWarpSmem &ws = smem[threadIdx.x >> 5];
    const uint32_t bar0 = smem_u32(&ws.bar[0]);
    const uint32_t ring0 = smem_u32(&ws.ring[0]);
    for (uint32_t k = threadIdx.x; k < 256; k += DEC_WARPS * 32) lut[k] = dec_lut_entry(k);
    if (lane == 0) {
        for (int s = 0; s < DEC_NB; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncthreads();
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
        // Control bytes are fetched one 512-sample step ahead (ka / kb: first / second 256 samples of the coming step), so
        // the loads have a whole step to land; positions past the last control byte read as 0.
        const uint32_t steps = n >> 9;
        const uint8_t *kp = keys + 2 * lane;  // the lane's control bytes of the step being fetched
        uint32_t ki = 2 * lane;
        uint32_t ka = dec_load_keys(kp, ki, nkeys, true), kb = dec_load_keys(kp + 64, ki + 64, nkeys, true);
        bool ok = true;
        for (uint32_t it = 0; it < steps; ++it) {
            const uint32_t ka_now = ka, kb_now = kb;
            const bool guard = it + 2 > steps;  // unguarded only when the next step is a whole one as well
            kp += 128;
            ki += 128;
            ka = dec_load_keys(kp, ki, nkeys, guard);
            kb = dec_load_keys(kp + 64, ki + 64, nkeys, guard);
            if (!__any_sync(FULL, ((ka_now | kb_now) & 0xAAAAu) != 0)) {
                ok = dec_iteration512(s, ws.ring, lut, ka_now, kb_now, out, lane, phase_bits, bar0, ring0);
            } else {
                ok = dec_iteration<false>(s, ws.ring, lut, ka_now, out, 256, lane, phase_bits, bar0, ring0) &&
                     dec_iteration<false>(s, ws.ring, lut, kb_now, out + 256, 256, lane, phase_bits, bar0, ring0);
            }
            if (!ok) break;
            out += 512;
        }
        if (ok) {
            uint32_t rest = n & 511u;
            if (rest >= 256) {
                ok = dec_iteration<false>(s, ws.ring, lut, ka, out, 256, lane, phase_bits, bar0, ring0);
                ka = kb;
                out += 256;
                rest -= 256;
            }
            if (ok && rest) ok = dec_iteration<true>(s, ws.ring, lut, ka, out, (int)rest, lane, phase_bits, bar0, ring0);
        }
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
// peek: u32 headers only
// ------------------------------------------------------------------------------------------------
__global__ void svbzd_peek_kernel(const uint8_t *svb, const uint64_t *svb_off, const uint32_t *svb_len,
                                  uint64_t n_reads, uint32_t *n_samples) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    uint32_t n = 0;
    if (svb_len[r] >= 4) {
        const uint8_t *p = svb + svb_off[r];
        n = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    n_samples[r] = n;
}

// ------------------------------------------------------------------------------------------------
// dense gather (slot layout -> packed slab): three-kernel exclusive scan of the rounded lengths,
// then one warp per stream copies it
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 1024;

__device__ __forceinline__ uint64_t round_up_u64(uint64_t v, uint32_t a) { return (v + a - 1) / a * a; }

__device__ uint64_t block_incl_scan(uint64_t v, uint64_t *warp_sums /*[32]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    if (lane == 31) warp_sums[wid] = v;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t t = __shfl_up_sync(FULL, w, d);
            if (lane >= d) w += t;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    if (wid > 0) v += warp_sums[wid - 1];
    return v;
}

__global__ void __launch_bounds__(SCAN_T) scan_block_sums_kernel(const uint32_t *len, uint64_t n, uint32_t align,
                                                                 uint64_t *block_sums) {
    __shared__ uint64_t ws[32];
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_T + threadIdx.x;
    uint64_t v = i < n ? round_up_u64(len[i], align) : 0;
    v = block_incl_scan(v, ws);
    if (threadIdx.x == SCAN_T - 1) block_sums[blockIdx.x] = v;
}
__global__ void __launch_bounds__(SCAN_T) scan_top_kernel(uint64_t *block_sums, uint64_t nblocks) {
    __shared__ uint64_t ws[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nblocks; base += SCAN_T) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < nblocks ? block_sums[i] : 0;
        const uint64_t incl = block_incl_scan(v, ws);
        const uint64_t carry = carry_s;
        if (i < nblocks) block_sums[i] = carry + incl - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == SCAN_T - 1) carry_s = carry + incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SCAN_T) scan_write_kernel(const uint32_t *len, uint64_t n, uint32_t align,
                                                            const uint64_t *block_sums, uint64_t *off) {
    __shared__ uint64_t ws[32];
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_T + threadIdx.x;
    const uint64_t v = i < n ? round_up_u64(len[i], align) : 0;
    const uint64_t incl = block_incl_scan(v, ws) + block_sums[blockIdx.x];
    if (i < n) {
        off[i] = incl - v;
        if (i == n - 1) off[n] = incl;
    }
}

constexpr int COPY_WARPS = 8;
__global__ void __launch_bounds__(COPY_WARPS * 32) gather_copy_kernel(const uint8_t *src, const uint64_t *src_off,
                                                                      const uint32_t *len, uint64_t n_reads,
                                                                      uint8_t *dst, const uint64_t *dst_off) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * COPY_WARPS + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * COPY_WARPS;
    for (uint64_t r = warp; r < n_reads; r += nwarps) {
        const uint8_t *s = src + src_off[r];
        uint8_t *d = dst + dst_off[r];
        const uint32_t n = len[r];
        if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15u) == 0) {
            const uint32_t n16 = n >> 4;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
            uint4 *d4 = reinterpret_cast<uint4 *>(d);
            for (uint32_t i = lane; i < n16; i += 32) d4[i] = s4[i];
            for (uint32_t i = (n16 << 4) + lane; i < n; i += 32) d[i] = s[i];
        } else {
            for (uint32_t i = lane; i < n; i += 32) d[i] = s[i];
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
    if (cudaFuncSetAttribute(svbzd_encode_bytes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(EncWarpSmem) * ENC_WARPS)) != cudaSuccess)
        return 0;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, svbzd_encode_kernel, ENC_WARPS * 32,
                                                      sizeof(EncWarpSmem) * ENC_WARPS) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, svbzd_encode_bytes_kernel, ENC_WARPS * 32,
                                                      sizeof(EncWarpSmem) * ENC_WARPS) != cudaSuccess)
        return 0;
    return n < nb ? n : nb;  // one grid size for both forms
}
int svbzd_decode_blocks_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, svbzd_decode_kernel, DEC_WARPS * 32, 0) != cudaSuccess)
        return 0;
    return n;
}

static unsigned persistent_grid(uint64_t n_reads, int warps, int num_sms, int blocks_per_sm) {
    uint64_t want = (n_reads + warps - 1) / warps;
    uint64_t cap = (uint64_t)num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
    uint64_t g = want < cap ? want : cap;
    return (unsigned)(g ? g : 1);
}

cudaError_t launch_svbzd_encode(const SvbEncodeArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const unsigned grid = persistent_grid(a.n_reads, ENC_WARPS, num_sms, blocks_per_sm);
    if (a.sig) svbzd_encode_kernel<<<grid, ENC_WARPS * 32, sizeof(EncWarpSmem) * ENC_WARPS, st>>>(a);
    else svbzd_encode_bytes_kernel<<<grid, ENC_WARPS * 32, sizeof(EncWarpSmem) * ENC_WARPS, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_svbzd_decode(const SvbDecodeArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    svbzd_decode_kernel<<<persistent_grid(a.n_reads, DEC_WARPS, num_sms, blocks_per_sm), DEC_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_svbzd_peek(const uint8_t *svb, const uint64_t *svb_off, const uint32_t *svb_len, uint64_t n_reads,
                              uint32_t *n_samples, cudaStream_t st) {
    if (n_reads == 0) return cudaSuccess;
    svbzd_peek_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(svb, svb_off, svb_len, n_reads, n_samples);
    return cudaGetLastError();
}

size_t compact_scratch_bytes(uint64_t n_reads) { return ((n_reads + SCAN_T - 1) / SCAN_T + 1) * sizeof(uint64_t); }

cudaError_t launch_scan(const uint32_t *len, uint64_t n, uint32_t align, uint64_t *off, void *scratch, cudaStream_t st) {
    if (n == 0) return cudaMemsetAsync(off, 0, sizeof(uint64_t), st);
    uint64_t *block_sums = static_cast<uint64_t *>(scratch);
    const uint64_t nblocks = (n + SCAN_T - 1) / SCAN_T;
    scan_block_sums_kernel<<<(unsigned)nblocks, SCAN_T, 0, st>>>(len, n, align, block_sums);
    scan_top_kernel<<<1, SCAN_T, 0, st>>>(block_sums, nblocks);
    scan_write_kernel<<<(unsigned)nblocks, SCAN_T, 0, st>>>(len, n, align, block_sums, off);
    return cudaGetLastError();
}

cudaError_t launch_compact(const uint8_t *src, const uint64_t *src_off, const uint32_t *len, uint64_t n_reads,
                           uint32_t align, uint8_t *dst, uint64_t *dst_off, void *scratch, cudaStream_t st,
                           int *n_launches) {
    *n_launches = 0;
    if (n_reads == 0) return cudaMemsetAsync(dst_off, 0, sizeof(uint64_t), st);
    uint64_t *block_sums = static_cast<uint64_t *>(scratch);
    const uint64_t nblocks = (n_reads + SCAN_T - 1) / SCAN_T;
    scan_block_sums_kernel<<<(unsigned)nblocks, SCAN_T, 0, st>>>(len, n_reads, align, block_sums);
    scan_top_kernel<<<1, SCAN_T, 0, st>>>(block_sums, nblocks);
    scan_write_kernel<<<(unsigned)nblocks, SCAN_T, 0, st>>>(len, n_reads, align, block_sums, dst_off);
    uint64_t g = (n_reads + COPY_WARPS - 1) / COPY_WARPS;
    if (g > 148ull * 8 * 4) g = 148ull * 8 * 4;
    gather_copy_kernel<<<(unsigned)g, COPY_WARPS * 32, 0, st>>>(src, src_off, len, n_reads, dst, dst_off);
    *n_launches = 4;
    return cudaGetLastError();
}

}  // namespace s5b
