// record_kernels.cu -- the glue kernels that keep a whole BLOW5 batch on the device between the codec kernels:
// find the signal inside packed records (binary record layout, slow5.c:2811-2950 / :3928-4074), size the next
// stage's slots, gather packed records, and build the output file image ([u64 size][record]..., slow5.c:4055-4060)
// so the host only does one H2D, one D2H and one fwrite per batch.  Plain byte moving: one warp per record.
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "../../include/slow5b200.h"

namespace s5b {

namespace {

__device__ __forceinline__ uint64_t ld_u64_unaligned(const uint8_t *p) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// warp-cooperative copy of n bytes, any alignment on either side
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    const uintptr_t ds = reinterpret_cast<uintptr_t>(dst), ss = reinterpret_cast<uintptr_t>(src);
    if (((ds ^ ss) & 15u) == 0) {
        // same phase: byte head, 128-bit body, byte tail
        uint32_t head = (uint32_t)((16u - (ds & 15u)) & 15u);
        if (head > n) head = n;
        if (lane < (int)head) dst[lane] = src[lane];
        const uint32_t body = (n - head) >> 4;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
        // four loads in flight per lane before the first store (source and destination may alias as far as the compiler knows,
        // so it keeps the order written here)
        uint32_t i = lane;
        for (; i + 96 < body; i += 128) {
            const uint4 v0 = s4[i], v1 = s4[i + 32], v2 = s4[i + 64], v3 = s4[i + 96];
            d4[i] = v0;
            d4[i + 32] = v1;
            d4[i + 64] = v2;
            d4[i + 96] = v3;
        }
        for (; i < body; i += 32) d4[i] = s4[i];
        for (uint32_t t = head + (body << 4) + lane; t < n; t += 32) dst[t] = src[t];
    } else if (n >= 64) {
        // different 16-byte phase: destination-aligned 128-bit stores, every quad assembled from five aligned source
        // words with funnel shifts (the fifth word may lie up to 7 bytes past the last source byte: every slab this
        // file copies from is padded by >= 16 bytes)
        uint32_t head = (uint32_t)((16u - (ds & 15u)) & 15u);
        if (lane < (int)head) dst[lane] = src[lane];
        const uint32_t body = (n - head) >> 4;
        const uint8_t *s0 = src + head;
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s0) & 3u) * 8u;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(s0 - (sh >> 3));
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
        uint32_t i = lane;
        for (; i + 32 < body; i += 64) {  // two quads per lane in flight
            const uint32_t *q = w + 4 * i, *p = q + 128;
            const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
            const uint32_t b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3], b4 = p[4];
            uint4 v, u;
            v.x = __funnelshift_r(a0, a1, sh);
            v.y = __funnelshift_r(a1, a2, sh);
            v.z = __funnelshift_r(a2, a3, sh);
            v.w = __funnelshift_r(a3, a4, sh);
            u.x = __funnelshift_r(b0, b1, sh);
            u.y = __funnelshift_r(b1, b2, sh);
            u.z = __funnelshift_r(b2, b3, sh);
            u.w = __funnelshift_r(b3, b4, sh);
            d4[i] = v;
            d4[i + 32] = u;
        }
        for (; i < body; i += 32) {
            const uint32_t *q = w + 4 * i;
            const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
            uint4 v;
            v.x = __funnelshift_r(a0, a1, sh);
            v.y = __funnelshift_r(a1, a2, sh);
            v.z = __funnelshift_r(a2, a3, sh);
            v.w = __funnelshift_r(a3, a4, sh);
            d4[i] = v;
        }
        for (uint32_t t = head + (body << 4) + lane; t < n; t += 32) dst[t] = src[t];
    } else if (((ds ^ ss) & 3u) == 0) {
        uint32_t head = (uint32_t)((4u - (ds & 3u)) & 3u);
        if (head > n) head = n;
        if (lane < (int)head) dst[lane] = src[lane];
        const uint32_t body = (n - head) >> 2;
        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src + head);
        uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + head);
        for (uint32_t i = lane; i < body; i += 32) d4[i] = s4[i];
        for (uint32_t i = head + (body << 2) + lane; i < n; i += 32) dst[i] = src[i];
    } else {
        for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    }
}

constexpr int RK_WARPS = 8;

__global__ void rec_locate_kernel(const uint8_t *rec, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n,
                                  int sig_is_svb, RecArrays a, const int32_t *in_status, const AuxLayout lay) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t *p = rec + rec_off[r];
    const uint64_t len = rec_len[r];
    int32_t st = in_status ? in_status[r] : S5B_OK;  // a record whose decompression failed is not looked at
    uint32_t head = 0, ns = 0, sig_at = 0, sig_bytes = 0, aux = 0;
    if (st != S5B_OK) {
    } else if (len < 2) {
        st = S5B_ERR_PRESS;
    } else {
        const uint32_t rid = (uint32_t)p[0] | ((uint32_t)p[1] << 8);
        head = 2 + rid + 4 + 32;
        if ((uint64_t)head + 8 > len) {
            st = S5B_ERR_PRESS;
        } else {
            const uint64_t lrs = ld_u64_unaligned(p + head);
            sig_at = head + 8;
            if (lay.rg_n) {  // the record's read group must be one the header names (it is renumbered through a table later)
                const uint8_t *g = p + head - 36;
                const uint32_t rg = (uint32_t)g[0] | ((uint32_t)g[1] << 8) | ((uint32_t)g[2] << 16) | ((uint32_t)g[3] << 24);
                if (rg >= lay.rg_n) st = S5B_ERR_PRESS;
            }
            // the field counts samples for a raw signal and bytes for a compressed one (slow5.c:3983-3987)
            // sig_is_svb: 0 raw int16, 1 svb-zd stream (u32 sample count first), 2 ex-zd stream (u64 count at byte 1)
            const uint64_t sb = sig_is_svb ? lrs : lrs * 2;
            if (sb > len - sig_at || (sig_is_svb == 1 && sb < 4) || (sig_is_svb == 2 && sb < 16)) {
                st = S5B_ERR_PRESS;
            } else {
                sig_bytes = (uint32_t)sb;
                if (sig_is_svb == 1) {
                    const uint8_t *q = p + sig_at;
                    ns = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
                } else if (sig_is_svb == 2) {
                    const uint64_t nin = ld_u64_unaligned(p + sig_at + 1);  // slow5_press.c:1790-1793
                    if (nin > 0xFFFFFFFFull) st = S5B_ERR_PRESS;
                    ns = (uint32_t)nin;
                } else {
                    ns = (uint32_t)lrs;
                }
                // every sample of a compressed signal takes at least one stream byte (slow5_press.c:1062-1079, :1721-1776):
                // a count beyond that cannot decode, and the stages downstream size their slabs from it
                if (sig_is_svb && ns > sb) st = S5B_ERR_PRESS;
                aux = (uint32_t)(len - sig_at - sb);
                // the auxiliary section must be exactly the file's columns (slow5_rec_aux_parse, slow5.c:3088-3166): a primitive
                // takes its size, an array a u64 count and count elements
                if (st == S5B_OK && lay.n != AUX_LAYOUT_UNKNOWN) {
                    const uint8_t *q = p + sig_at + sb;
                    uint64_t at = 0;
                    bool ok = true;
                    for (uint32_t f = 0; f < lay.n && ok; ++f) {
                        uint64_t cnt = 1;
                        if ((lay.array_mask >> f) & 1ull) {
                            if (at + 8 > aux) {
                                ok = false;
                                break;
                            }
                            cnt = ld_u64_unaligned(q + at);
                            at += 8;
                        }
                        if (cnt > aux || at + cnt * lay.size[f] > aux) ok = false;
                        at += cnt * lay.size[f];
                    }
                    if (!ok || at != aux) st = S5B_ERR_PRESS;
                }
                // degrade, bit count chosen from the header: the record itself must belong to that dataset (slow5_reccmp,
                // src/degrade.c:195-211).  digitisation sits 32 bytes in front of the length field, sampling_rate 8.
                if (st == S5B_OK && lay.ds_check) {
                    const double dig = __longlong_as_double((long long)ld_u64_unaligned(p + head - 32));
                    const double rate = __longlong_as_double((long long)ld_u64_unaligned(p + head - 8));
                    if (dig != (double)lay.ds_digitisation || rate != (double)lay.ds_sampling_rate) st = S5B_ERR_DATASET;
                }
            }
        }
    }
    a.head_len[r] = head;
    a.n_samples[r] = st == S5B_OK ? ns : 0;
    a.sig_at[r] = sig_at;
    a.sig_bytes[r] = sig_bytes;
    a.aux_len[r] = aux;
    a.status[r] = st;
}

__global__ void rec_plan_kernel(int mode, uint64_t n, RecArrays a, const uint32_t *aux_in, uint32_t param, uint32_t *out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t v = 0;
    const uint32_t ns = a.n_samples[r];
    switch (mode) {
        case PLAN_SIG_SAMPLES: v = ns; break;                                       // scanned with align 8 (samples)
        case PLAN_SVB_BOUND: v = 4u + (ns + 3u) / 4u + 3u * ns; break;              // s5b_svbzd_bound
        case PLAN_EXZD_BOUND: v = 2u * ns + 1024u; break;                           // s5b_exzd_bound
        case PLAN_PACKED_LEN: v = a.head_len[r] + 8u + aux_in[r] + a.aux_len[r]; break;  // aux_in = signal bytes to store
        case PLAN_ZLIB_BOUND: v = aux_in[r] + 6u * (aux_in[r] / 6144u + 2u) + 8u; break; // == deflate_bound() (DEF_BLOCK 6144)
        case PLAN_IMAGE_LEN: v = aux_in[r] + 8u; break;
        case PLAN_INFLATE_GUESS: v = aux_in[r] * param + 1024u; break;
        case PLAN_SPLIT: v = a.head_len[r] + 8u + 4u + (ns + 3u) / 4u; break;       // start of the svb-zd data bytes
        case PLAN_SIG_BYTES_RAW: v = 2u * ns; break;
        case PLAN_PACKED_IMAGE_LEN: v = a.head_len[r] + 8u + aux_in[r] + a.aux_len[r] + 8u; break;  // packed record + size prefix
    }
    out[r] = v;
}

__global__ void __launch_bounds__(RK_WARPS * 32) sig_extract_kernel(const uint8_t *rec, const uint64_t *rec_off, RecArrays a,
                                                                    uint64_t n, int16_t *sig, const uint64_t *sig_off) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * RK_WARPS + (threadIdx.x >> 5), nw = (uint64_t)gridDim.x * RK_WARPS;
    for (uint64_t r = warp; r < n; r += nw) {
        if (a.status[r] != S5B_OK) continue;
        warp_copy(reinterpret_cast<uint8_t *>(sig + sig_off[r]), rec + rec_off[r] + a.sig_at[r], a.sig_bytes[r], lane);
    }
}

__global__ void __launch_bounds__(RK_WARPS * 32) rec_pack_kernel(const uint8_t *rec, const uint64_t *rec_off, RecArrays a,
                                                                 uint64_t n, const uint8_t *sig_src,
                                                                 const uint64_t *sig_src_off, const uint32_t *sig_src_len,
                                                                 int sig_src_is_samples, int sig_out_compressed, uint8_t *out,
                                                                 const uint64_t *out_off, int image, const uint64_t *base_ptr,
                                                                 const uint64_t *res, uint64_t *abs_off) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * RK_WARPS + (threadIdx.x >> 5), nw = (uint64_t)gridDim.x * RK_WARPS;
    // image mode: the packed records ARE the output ([u64 size][record] back to back at out + *base_ptr, slow5.c:4055-4060)
    if (image && res && (int32_t)res[1] == S5B_ERR_NOSPACE && res[2] == ~0ull) return;
    const uint64_t base = image && base_ptr ? *base_ptr : 0;
    for (uint64_t r = warp; r < n; r += nw) {
        if (image && abs_off && lane == 0) {
            abs_off[r] = base + out_off[r];
            if (r == n - 1) abs_off[n] = base + out_off[n];
        }
        if (a.status[r] != S5B_OK) continue;
        const uint8_t *in = rec + rec_off[r];
        uint8_t *o = out + base + out_off[r];
        if (image) {
            const uint64_t sz = out_off[r + 1] - out_off[r] - 8;
            if (lane < 8) o[lane] = (uint8_t)(sz >> (8 * lane));
            o += 8;
        }
        const uint32_t head = a.head_len[r];
        warp_copy(o, in, head, lane);
        // signal source: the input record itself (pass-through, sig_src == NULL), a byte slab (svb slots, offsets in
        // bytes) or the int16 signal slab (offsets in samples)
        const uint8_t *sbase = sig_src ? sig_src : in;
        const uint64_t soff = !sig_src ? a.sig_at[r] : sig_src_is_samples ? sig_src_off[r] * 2 : sig_src_off[r];
        const uint32_t sbytes = !sig_src ? a.sig_bytes[r] : sig_src_is_samples ? a.n_samples[r] * 2 : sig_src_len[r];
        const uint64_t lrs = sig_out_compressed ? (uint64_t)sbytes : (uint64_t)a.n_samples[r];  // slow5.c:3983-3987
        if (lane < 8) o[head + lane] = (uint8_t)(lrs >> (8 * lane));
        warp_copy(o + head + 8, sbase + soff, sbytes, lane);
        warp_copy(o + head + 8 + sbytes, in + a.sig_at[r] + a.sig_bytes[r], a.aux_len[r], lane);
    }
}

__global__ void __launch_bounds__(RK_WARPS * 32) image_gather_kernel(const uint8_t *src, const uint64_t *src_off,
                                                                     const uint32_t *len, uint64_t n, uint8_t *img,
                                                                     const uint64_t *img_off, const uint64_t *base_ptr,
                                                                     const uint64_t *res, uint64_t *abs_off) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * RK_WARPS + (threadIdx.x >> 5), nw = (uint64_t)gridDim.x * RK_WARPS;
    if (res && (int32_t)res[1] == S5B_ERR_NOSPACE && res[2] == ~0ull) return;  // the image does not fit: nothing is written
    const uint64_t base = base_ptr ? *base_ptr : 0;
    for (uint64_t r = warp; r < n; r += nw) {
        uint8_t *o = img + base + img_off[r];
        if (abs_off && lane == 0) {
            abs_off[r] = base + img_off[r];
            if (r == n - 1) abs_off[n] = base + img_off[n];
        }
        const uint64_t sz = len[r];
        if (lane < 8) o[lane] = (uint8_t)(sz >> (8 * lane));  // record size prefix, slow5.c:4055-4060
        warp_copy(o + 8, src + src_off[r], len[r], lane);
    }
}

// End of a transcoding pass over one chunk: the first failed record over up to four per-record status arrays (lowest
// record index wins, like a serial loop over the batch would report), the chunk's image size, the capacity check.
// res[0] = image bytes, res[1] = error code (int32), res[2] = its record index (~0 for a call-level error); res[3] is scratch.
// (two steps: every CTA folds its slice of the status arrays into res[3] with one atomicMin, then one thread writes the verdict;
// a single CTA walking 250 k records x 4 arrays took 0.28 ms per chunk)
__global__ void __launch_bounds__(256) recode_status_kernel(uint64_t n, const int32_t *s0, const int32_t *s1, const int32_t *s2,
                                                            const int32_t *s3, uint64_t *res) {
    unsigned long long mine = ~0ull;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        int32_t e = S5B_OK;
        if (s0 && s0[r] != S5B_OK) e = s0[r];
        else if (s1 && s1[r] != S5B_OK) e = s1[r];
        else if (s2 && s2[r] != S5B_OK) e = s2[r];
        else if (s3 && s3[r] != S5B_OK) e = s3[r];
        if (e != S5B_OK) {
            mine = (r << 8) | (uint32_t)(-e & 0xff);
            break;  // this thread's later records have higher indices
        }
    }
    if (mine != ~0ull) atomicMin(reinterpret_cast<unsigned long long *>(res + 3), mine);
}
__global__ void recode_finish_kernel(uint64_t n, const uint64_t *img_off, const uint64_t *base_ptr, uint64_t cap, uint64_t *res) {
    const unsigned long long best = res[3];
    const uint64_t total = img_off[n];
    const uint64_t base = base_ptr ? *base_ptr : 0;
    res[0] = total;
    if (best != ~0ull) {
        res[1] = (uint64_t)(int64_t)(-(int32_t)(best & 0xff));
        res[2] = best >> 8;
    } else if (cap && base + total > cap) {
        res[1] = (uint64_t)(int64_t)S5B_ERR_NOSPACE;
        res[2] = ~0ull;
    } else {
        res[1] = 0;
        res[2] = 0;
    }
}
// base += the chunk's image bytes, unless the chunk failed (a failed chunk is redone by the careful path or ends the call);
// acc (optional) accumulates: acc[0] = total image bytes so far, acc[1] = first error of the whole call
__global__ void recode_advance_kernel(uint64_t *base_ptr, const uint64_t *res, uint64_t *acc) {
    const int32_t e = (int32_t)res[1];
    if (e == S5B_OK) {
        if (base_ptr) *base_ptr += res[0];
        if (acc) acc[0] += res[0];
    } else if (acc && (int32_t)acc[1] == S5B_OK) {
        acc[1] = (uint64_t)(int64_t)e;
    }
}
__global__ void rebase_off_kernel(uint64_t *dst, const uint64_t *src, uint64_t n, uint64_t sub) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) dst[r] = src[r] - sub;
}

__global__ void rec_sig_abs_kernel(const uint64_t *rec_off, RecArrays a, uint64_t n, uint64_t *out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) out[r] = rec_off[r] + a.sig_at[r];
    if (r == n) out[n] = 0;
}

// read_id of every (decompressed prefix of a) record: len[r] = id bytes when the prefix holds the whole id, 0xFFFFFFFF when it
// does not (the caller decompresses that record in full), src[r] = where the id bytes start (slow5_idx.c:322-334)
__global__ void rec_ids_kernel(const uint64_t *rec_off, const uint32_t *have, const int32_t *status, uint64_t n, uint32_t *len,
                               uint64_t *src, const uint8_t *rec) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t l = 0xFFFFFFFFu;
    const uint64_t at = rec_off[r];
    if (status[r] == S5B_OK && have[r] >= 2) {
        const uint32_t rid = (uint32_t)rec[at] | ((uint32_t)rec[at + 1] << 8);
        if (have[r] >= 2 + rid) l = rid;
    }
    len[r] = l;
    src[r] = at + 2;
}

unsigned rk_grid(uint64_t n) {
    uint64_t g = (n + RK_WARPS - 1) / RK_WARPS;
    if (g > 148ull * 8) g = 148ull * 8;
    return (unsigned)(g ? g : 1);
}

}  // namespace

cudaError_t launch_rec_locate(const uint8_t *rec, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n,
                              int sig_is_svb, RecArrays a, cudaStream_t st, const int32_t *in_status, const AuxLayout *aux) {
    if (!n) return cudaSuccess;
    rec_locate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rec, rec_off, rec_len, n, sig_is_svb, a, in_status,
                                                                   aux ? *aux : AuxLayout());
    return cudaGetLastError();
}
cudaError_t launch_recode_finish(uint64_t n, const uint64_t *img_off, const uint64_t *base_ptr, uint64_t cap, const int32_t *s0,
                                 const int32_t *s1, const int32_t *s2, const int32_t *s3, uint64_t *res, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(res + 3, 0xff, sizeof(uint64_t), st);
    if (e != cudaSuccess) return e;
    if (n) {
        const uint64_t want = (n + 255) / 256;
        recode_status_kernel<<<(unsigned)(want < 1184 ? want : 1184), 256, 0, st>>>(n, s0, s1, s2, s3, res);
    }
    recode_finish_kernel<<<1, 1, 0, st>>>(n, img_off, base_ptr, cap, res);
    return cudaGetLastError();
}
__global__ void recode_publish_kernel(const uint64_t *res, volatile uint64_t *host) {
    host[0] = res[0];
    host[1] = res[1];
    host[2] = res[2];
    __threadfence_system();
}
cudaError_t launch_recode_publish(const uint64_t *res, uint64_t *host_mapped, cudaStream_t st) {
    recode_publish_kernel<<<1, 1, 0, st>>>(res, host_mapped);
    return cudaGetLastError();
}
cudaError_t launch_recode_advance(uint64_t *base_ptr, const uint64_t *res, uint64_t *acc, cudaStream_t st) {
    recode_advance_kernel<<<1, 1, 0, st>>>(base_ptr, res, acc);
    return cudaGetLastError();
}
cudaError_t launch_rebase_off(uint64_t *dst, const uint64_t *src, uint64_t n, uint64_t sub, cudaStream_t st) {
    if (!n) return cudaSuccess;
    rebase_off_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, src, n, sub);
    return cudaGetLastError();
}
cudaError_t launch_rec_sig_abs(const uint64_t *rec_off, RecArrays a, uint64_t n, uint64_t *out, cudaStream_t st) {
    rec_sig_abs_kernel<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(rec_off, a, n, out);
    return cudaGetLastError();
}
cudaError_t launch_rec_ids(const uint8_t *rec, const uint64_t *rec_off, const uint32_t *have, const int32_t *status, uint64_t n,
                           uint32_t *len, uint64_t *src, cudaStream_t st) {
    if (!n) return cudaSuccess;
    rec_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rec_off, have, status, n, len, src, rec);
    return cudaGetLastError();
}
cudaError_t launch_rec_plan(int mode, uint64_t n, RecArrays a, const uint32_t *aux_in, uint32_t param, uint32_t *out,
                            cudaStream_t st) {
    if (!n) return cudaSuccess;
    rec_plan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mode, n, a, aux_in, param, out);
    return cudaGetLastError();
}
cudaError_t launch_sig_extract(const uint8_t *rec, const uint64_t *rec_off, RecArrays a, uint64_t n, int16_t *sig,
                               const uint64_t *sig_off, cudaStream_t st) {
    if (!n) return cudaSuccess;
    sig_extract_kernel<<<rk_grid(n), RK_WARPS * 32, 0, st>>>(rec, rec_off, a, n, sig, sig_off);
    return cudaGetLastError();
}
__global__ void rec_rg_remap_kernel(uint8_t *rec, const uint64_t *rec_off, RecArrays a, uint64_t n, const uint32_t *rg_map) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || a.status[r] != S5B_OK) return;
    uint8_t *g = rec + rec_off[r] + a.head_len[r] - 36;  // the field in front of the four doubles
    const uint32_t rg = rg_map[(uint32_t)g[0] | ((uint32_t)g[1] << 8) | ((uint32_t)g[2] << 16) | ((uint32_t)g[3] << 24)];
    g[0] = (uint8_t)rg, g[1] = (uint8_t)(rg >> 8), g[2] = (uint8_t)(rg >> 16), g[3] = (uint8_t)(rg >> 24);
}
cudaError_t launch_rec_rg_remap(uint8_t *rec, const uint64_t *rec_off, RecArrays a, uint64_t n, const uint32_t *rg_map, cudaStream_t st) {
    if (!n) return cudaSuccess;
    rec_rg_remap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rec, rec_off, a, n, rg_map);
    return cudaGetLastError();
}
cudaError_t launch_rec_pack(const uint8_t *rec, const uint64_t *rec_off, RecArrays a, uint64_t n, const uint8_t *sig_src,
                            const uint64_t *sig_src_off, const uint32_t *sig_src_len, int sig_src_is_samples,
                            int sig_out_compressed, uint8_t *out, const uint64_t *out_off, cudaStream_t st, int image,
                            const uint64_t *base_ptr, const uint64_t *res, uint64_t *abs_off) {
    if (!n) return cudaSuccess;
    rec_pack_kernel<<<rk_grid(n), RK_WARPS * 32, 0, st>>>(rec, rec_off, a, n, sig_src, sig_src_off, sig_src_len,
                                                         sig_src_is_samples, sig_out_compressed, out, out_off, image, base_ptr,
                                                         res, abs_off);
    return cudaGetLastError();
}
cudaError_t launch_image_gather(const uint8_t *src, const uint64_t *src_off, const uint32_t *len, uint64_t n, uint8_t *img,
                                const uint64_t *img_off, cudaStream_t st, const uint64_t *base_ptr, const uint64_t *res,
                                uint64_t *abs_off) {
    if (!n) return cudaSuccess;
    image_gather_kernel<<<rk_grid(n), RK_WARPS * 32, 0, st>>>(src, src_off, len, n, img, img_off, base_ptr, res, abs_off);
    return cudaGetLastError();
}

}  // namespace s5b
