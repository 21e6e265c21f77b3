// qts_kernels.cu -- lossy signal degradation ("qts": quantise the signal), the per-sample step `slow5tools degrade` puts between
// record decode and re-encode (src/degrade.c:240-263 -> slow5_rec_qts_round -> slow5_arr_qts_round, slow5_press.c:1965-2019):
// the b least significant bits of every sample are zeroed by rounding to the nearest multiple of 2^b, halves up, and the int
// result is stored back into the int16 (so 32767 with b = 1 becomes -32768, as in the reference).
//
// Per sample that is (x + 2^(b-1)) & ~(2^b - 1) in 16-bit wrap-around arithmetic: adding half a step carries into the kept bits
// exactly when the dropped bits are >= half a step.  Two samples are done per 32-bit register (a 16-bit-lane add and one AND),
// eight per 128-bit load / store.  Pure streaming work: 2 bytes in, 2 bytes out per sample, bound by HBM.
//
// The transcoder calls it on a lane's whole sample slab, whose length is known only on the device (the scan of the per-record
// sample counts): the kernel reads the count from there and the grid strides over it.
#include "s5b_kernels.h"

namespace s5b {

namespace {

constexpr int QTS_THREADS = 256;
constexpr int QTS_CTAS_PER_SM = 8;

__device__ __forceinline__ uint32_t qts_pair(uint32_t x, uint32_t half2, uint32_t keep2) { return __vadd2(x, half2) & keep2; }

__global__ void __launch_bounds__(QTS_THREADS) qts_round_kernel(int16_t *sig, uint64_t n_imm, const uint64_t *n_dev, int bits) {
    const uint64_t n = n_dev ? *n_dev : n_imm;
    const uint32_t half = 1u << (bits - 1);
    const uint32_t keep = ~((1u << bits) - 1u) & 0xffffu;
    const uint32_t half2 = half | (half << 16), keep2 = keep | (keep << 16);
    // samples in front of the first 16-byte boundary, whole 128-bit words, and the rest
    const uintptr_t addr = reinterpret_cast<uintptr_t>(sig);
    uint64_t head = ((16u - (addr & 15u)) & 15u) >> 1;
    if (head > n) head = n;
    const uint64_t body = (n - head) >> 3;
    const uint64_t tail_at = head + (body << 3);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthr = (uint64_t)gridDim.x * blockDim.x;
    uint4 *v = reinterpret_cast<uint4 *>(sig + head);
    // two words in flight per thread
    uint64_t i = tid;
    for (; i + nthr < body; i += 2 * nthr) {
        uint4 a = v[i], b = v[i + nthr];
        a.x = qts_pair(a.x, half2, keep2);
        a.y = qts_pair(a.y, half2, keep2);
        a.z = qts_pair(a.z, half2, keep2);
        a.w = qts_pair(a.w, half2, keep2);
        b.x = qts_pair(b.x, half2, keep2);
        b.y = qts_pair(b.y, half2, keep2);
        b.z = qts_pair(b.z, half2, keep2);
        b.w = qts_pair(b.w, half2, keep2);
        v[i] = a;
        v[i + nthr] = b;
    }
    if (i < body) {
        uint4 a = v[i];
        a.x = qts_pair(a.x, half2, keep2);
        a.y = qts_pair(a.y, half2, keep2);
        a.z = qts_pair(a.z, half2, keep2);
        a.w = qts_pair(a.w, half2, keep2);
        v[i] = a;
    }
    if (blockIdx.x == 0) {
        uint16_t *s = reinterpret_cast<uint16_t *>(sig);
        if (threadIdx.x < head) s[threadIdx.x] = (uint16_t)((s[threadIdx.x] + half) & keep);
        const uint64_t t = tail_at + threadIdx.x;
        if (threadIdx.x < 8 && t < n) s[t] = (uint16_t)((s[t] + half) & keep);
    }
}

}  // namespace

cudaError_t launch_qts_round(int16_t *sig, uint64_t n_samples, const uint64_t *d_n_samples, int bits, int num_sms, cudaStream_t st) {
    if (bits < 1 || bits > 16) return cudaErrorInvalidValue;
    uint64_t ctas = (uint64_t)num_sms * QTS_CTAS_PER_SM;
    if (!d_n_samples) {  // a known length: no more CTAs than there is work
        const uint64_t want = (n_samples / 8 + QTS_THREADS - 1) / QTS_THREADS;
        if (want < ctas) ctas = want ? want : 1;
    }
    qts_round_kernel<<<(unsigned)ctas, QTS_THREADS, 0, st>>>(sig, n_samples, d_n_samples, bits);
    return cudaGetLastError();
}

}  // namespace s5b
