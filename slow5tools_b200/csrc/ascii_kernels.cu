// ascii_kernels.cu -- the raw_signal column of SLOW5 text records on the GPU (second half).
//
// `slow5tools view` to SLOW5 prints every sample with sprintf("%d,") (slow5_rec_to_mem, slow5lib/src/slow5.c:3866-3878)
// and reads it back with strsep + slow5_ato_int16 (slow5_rec_parse, slow5.c:2754-2778; slow5_misc.c:122-139,303-319):
// the other per-sample loop of the record path next to the codecs.  Both directions here, one warp per read:
//   ascii_size_kernel / ascii_format_kernel   int16 samples -> "v0,v1,...,vN-1" (no trailing comma; '-' for negatives)
//   ascii_parse_kernel                        text -> int16 samples, with the reference's acceptance rules
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "../../include/slow5b200.h"

namespace s5b {

namespace asc {

constexpr int AW = 8;                 // warps per CTA
constexpr int STAGE = 2048 + 64;      // staged text bytes per warp (an iteration of 256 samples makes <= 1792)
constexpr int TILE = 1024;            // parse: text bytes per tile

__device__ __forceinline__ uint32_t ndigits(uint32_t a) {  // a < 32769
    return 1u + (a >= 10u) + (a >= 100u) + (a >= 1000u) + (a >= 10000u);
}

// characters of one sample, least significant byte first in the returned word: [-]digits followed by ','
// (the caller drops the comma of the last sample).  *k = number of characters including the comma.
__device__ __forceinline__ uint64_t sample_chars(int v, uint32_t *k) {
    const uint32_t neg = v < 0;
    uint32_t a = neg ? (uint32_t)(-v) : (uint32_t)v;
    const uint32_t nd = ndigits(a);
    uint64_t w = ',';
    for (uint32_t i = 0; i < nd; ++i) {
        w = (w << 8) | (uint64_t)('0' + a % 10u);
        a /= 10u;
    }
    if (neg) w = (w << 8) | (uint64_t)'-';
    *k = nd + neg + 1u;
    return w;
}

__global__ void __launch_bounds__(AW * 32) ascii_size_kernel(const int16_t *sig, const uint64_t *sig_off, const uint32_t *n_samples,
                                                             uint64_t n_reads, uint32_t *text_len) {
    const int lane = threadIdx.x & 31;
    const uint64_t r = (uint64_t)blockIdx.x * AW + (threadIdx.x >> 5);
    if (r >= n_reads) return;
    const int16_t *s = sig + sig_off[r];
    const uint32_t n = n_samples[r];
    uint32_t c = 0;
    for (uint32_t i = lane; i < n; i += 32) {
        const int v = s[i];
        c += ndigits((uint32_t)(v < 0 ? -v : v)) + (v < 0) + 1u;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(FULL, c, d);
    if (lane == 0) text_len[r] = n ? c - 1u : 0u;  // no comma after the last sample
}

__global__ void __launch_bounds__(AW * 32) ascii_format_kernel(const int16_t *sig, const uint64_t *sig_off, const uint32_t *n_samples,
                                                               uint64_t n_reads, uint8_t *text, const uint64_t *text_off) {
    __shared__ __align__(16) uint8_t stage_all[AW][STAGE];
    uint8_t *stage = stage_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint64_t r = (uint64_t)blockIdx.x * AW + (threadIdx.x >> 5);
    if (r >= n_reads) return;
    const int16_t *s = sig + sig_off[r];
    const uint32_t n = n_samples[r];
    uint8_t *dst = text + text_off[r];
    const uint64_t total = text_off[r + 1] - text_off[r];
    // stage[i] <-> gbase[i], gbase 16-byte aligned; [head, fill) are the valid staged bytes not yet stored
    uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
    uint8_t *gbase = dst - head;
    uint32_t fill = head;
    uint64_t written = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 256) {
        // lane owns samples i0 + 8 lane .. + 7 (the slab is padded to 8 samples per read: one 128-bit load)
        const uint32_t first = i0 + 8u * lane;
        int v[8];
        {
            const uint4 q = first < n ? *reinterpret_cast<const uint4 *>(s + first) : make_uint4(0, 0, 0, 0);
            v[0] = (int16_t)(q.x & 0xffffu), v[1] = (int16_t)(q.x >> 16), v[2] = (int16_t)(q.y & 0xffffu), v[3] = (int16_t)(q.y >> 16);
            v[4] = (int16_t)(q.z & 0xffffu), v[5] = (int16_t)(q.z >> 16), v[6] = (int16_t)(q.w & 0xffffu), v[7] = (int16_t)(q.w >> 16);
        }
        uint64_t w[8];
        uint32_t k[8], mine = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            w[j] = sample_chars(v[j], &k[j]);
            if (first + j >= n) k[j] = 0;
            else if (first + j == n - 1) --k[j];  // the last sample of the read has no comma
            mine += k[j];
        }
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        uint32_t at = fill + incl - mine;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint64_t x = w[j];
            for (uint32_t q = 0; q < k[j]; ++q) {
                stage[at++] = (uint8_t)x;
                x >>= 8;
            }
        }
        fill += __shfl_sync(FULL, incl, 31);
        __syncwarp();
        // store the complete 16-byte segments, carry the rest to the front
        const uint32_t nseg = fill >> 4;
        for (uint32_t seg = lane; seg < nseg; seg += 32) {
            const uint32_t lo = seg * 16;
            if (lo >= head) {
                *reinterpret_cast<uint4 *>(gbase + lo) = *reinterpret_cast<const uint4 *>(stage + lo);
            } else {
                for (uint32_t b = head; b < 16; ++b) gbase[b] = stage[b];  // ragged first segment of the read
            }
        }
        __syncwarp();
        if (nseg) {
            const uint32_t done = nseg * 16, keep = fill - done;
            uint8_t t = 0;
            if (lane < (int)keep) t = stage[done + lane];
            __syncwarp();
            if (lane < (int)keep) stage[lane] = t;
            written += done - head;
            gbase += done;
            head = 0;
            fill = keep;
            __syncwarp();
        }
    }
    for (uint32_t b = head + lane; b < fill; b += 32) gbase[b] = stage[b];
    (void)written;
    (void)total;
}

// ---- text -> samples ------------------------------------------------------------------------------------------------
// The reference accepts a token when it is non-empty, does not start with '0' unless it is "0" itself, and consists of
// digits and '-' only (slow5_int_check, slow5_misc.c:122-139); its value is strtol's: an optional leading '-', then
// digits up to the first other character ("-" alone is 0, "1-2" is 1); it must fit int16 (slow5_misc.c:303-319).
// status: 0, or S5B_ERR_ARG for a token the reference rejects / a count that differs from `expect`.
__global__ void __launch_bounds__(AW * 32) ascii_parse_kernel(const uint8_t *text, const uint64_t *text_off, const uint32_t *text_len,
                                                              uint64_t n_reads, int16_t *sig, const uint64_t *sig_off,
                                                              const uint32_t *expect, uint32_t *n_samples, int32_t *status) {
    __shared__ uint8_t tile_all[AW][TILE + 8];
    __shared__ uint16_t end_all[AW][TILE + 1];
    uint8_t *tile = tile_all[threadIdx.x >> 5];
    uint16_t *tok_end = end_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint64_t r = (uint64_t)blockIdx.x * AW + (threadIdx.x >> 5);
    if (r >= n_reads) return;
    const uint8_t *t = text + text_off[r];
    const uint32_t L = text_len[r];
    int16_t *out = sig + sig_off[r];
    const uint32_t cap = expect[r];
    uint32_t count = 0, w0 = 0;
    int bad = 0;
    // An empty column holds no samples.  Otherwise there are (commas + 1) tokens.
    while (w0 < L && !bad) {
        const uint32_t wn = min((uint32_t)TILE, L - w0);
        for (uint32_t i = lane; i < wn; i += 32) tile[i] = t[w0 + i];
        __syncwarp();
        const bool last_tile = w0 + wn == L;
        // token ends inside the tile: every comma, and the end of the text
        uint32_t ntok = 0;
        for (uint32_t i0 = 0; i0 < wn; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool is_end = i < wn && tile[i] == ',';
            const uint32_t m = __ballot_sync(FULL, is_end);
            if (is_end) tok_end[ntok + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
            ntok += __popc(m);
        }
        if (last_tile) {
            if (lane == 0) tok_end[ntok] = (uint16_t)wn;
            ++ntok;
        }
        __syncwarp();
        if (ntok == 0) {  // a token longer than a tile cannot be a sample
            bad = 1;
            break;
        }
        for (uint32_t j0 = 0; j0 < ntok; j0 += 32) {
            const uint32_t j = j0 + lane;
            int mybad = 0;
            if (j < ntok) {
                const uint32_t b = j ? tok_end[j - 1] + 1u : 0u, e = tok_end[j];
                const uint32_t len = e - b;
                if (len == 0 || len > 6 || (len > 1 && tile[b] == '0')) mybad = 1;
                int neg = 0, val = 0, stop = 0;
                for (uint32_t q = 0; q < len && q < 7; ++q) {
                    const uint32_t ch = tile[b + q];
                    if (ch == '-') {
                        if (q == 0) neg = 1;
                        else stop = 1;  // strtol stops at a '-' that is not the sign
                    } else if (ch >= '0' && ch <= '9') {
                        if (!stop) val = val * 10 + (int)(ch - '0');
                    } else {
                        mybad = 1;
                    }
                }
                if (neg) val = -val;
                if (val > 32767 || val < -32768) mybad = 1;
                const uint32_t idx = count + j;
                if (!mybad && idx < cap) out[idx] = (int16_t)val;
            }
            if (__any_sync(FULL, mybad)) bad = 1;
        }
        count += ntok;
        w0 += last_tile ? wn : tok_end[ntok - 1] + 1u;
        __syncwarp();
    }
    if (lane == 0) {
        n_samples[r] = count;
        status[r] = (bad || count != cap) ? S5B_ERR_ARG : S5B_OK;
    }
}

}  // namespace asc

cudaError_t launch_ascii_size(const int16_t *sig, const uint64_t *sig_off, const uint32_t *n_samples, uint64_t n_reads,
                              uint32_t *text_len, cudaStream_t st) {
    if (!n_reads) return cudaSuccess;
    asc::ascii_size_kernel<<<(unsigned)((n_reads + asc::AW - 1) / asc::AW), asc::AW * 32, 0, st>>>(sig, sig_off, n_samples, n_reads, text_len);
    return cudaGetLastError();
}
cudaError_t launch_ascii_format(const int16_t *sig, const uint64_t *sig_off, const uint32_t *n_samples, uint64_t n_reads,
                                uint8_t *text, const uint64_t *text_off, cudaStream_t st) {
    if (!n_reads) return cudaSuccess;
    asc::ascii_format_kernel<<<(unsigned)((n_reads + asc::AW - 1) / asc::AW), asc::AW * 32, 0, st>>>(sig, sig_off, n_samples, n_reads, text, text_off);
    return cudaGetLastError();
}
cudaError_t launch_ascii_parse(const uint8_t *text, const uint64_t *text_off, const uint32_t *text_len, uint64_t n_reads,
                               int16_t *sig, const uint64_t *sig_off, const uint32_t *expect, uint32_t *n_samples, int32_t *status,
                               cudaStream_t st) {
    if (!n_reads) return cudaSuccess;
    asc::ascii_parse_kernel<<<(unsigned)((n_reads + asc::AW - 1) / asc::AW), asc::AW * 32, 0, st>>>(text, text_off, text_len, n_reads, sig, sig_off,
                                                                                                 expect, n_samples, status);
    return cudaGetLastError();
}

}  // namespace s5b
