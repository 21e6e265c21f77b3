// ascii_api.cu -- host-batch entry points for the raw_signal column of SLOW5 text records (include/slow5b200.h):
// stored signals (raw / svb-zd / ex-zd) -> comma-joined decimal text, and text -> int16 samples, a batch per call.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/slow5b200.h"
#include "s5b_ctx.h"

using namespace s5b;

extern "C" {

int s5b_signal_to_ascii_batch_host(s5b_ctx_t *ctx, int sig_method, const void *const *ptrs, const size_t *counts, size_t n,
                                   char **out_ptrs, size_t *out_n) {
    if (!ctx || !ptrs || !counts || !out_ptrs || !out_n) return S5B_ERR_ARG;
    if (sig_method != S5B_COMPRESS_NONE && sig_method != S5B_COMPRESS_SVB_ZD && sig_method != S5B_COMPRESS_EX_ZD) return S5B_ERR_ARG;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = nullptr;
        out_n[i] = 0;
    }
    if (n == 0) return S5B_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->slot[0].stream;
    // ---- layout: stored bytes at 16-byte aligned offsets, samples at 8-sample aligned offsets
    std::vector<uint64_t> in_off(n + 1), sig_off(n + 1);
    std::vector<uint32_t> in_len(n), ns(n);
    uint64_t tot = 0, sig_tot = 0;
    for (size_t i = 0; i < n; ++i) {
        if ((!ptrs[i] && counts[i]) || counts[i] > 0xfffffff0ull) return S5B_ERR_ARG;
        in_len[i] = (uint32_t)counts[i];
        in_off[i] = tot;
        tot += round_up(counts[i], 16);
        uint64_t nn = 0;
        if (sig_method == S5B_COMPRESS_NONE) {
            nn = counts[i] / 2;
        } else if (sig_method == S5B_COMPRESS_SVB_ZD) {
            uint32_t v = 0;
            if (counts[i] >= 4) memcpy(&v, ptrs[i], 4);
            nn = v;
        } else if (counts[i] >= 9) {
            memcpy(&nn, static_cast<const uint8_t *>(ptrs[i]) + 1, 8);  // slow5_press.c:1790-1793
        }
        // a sample takes at least one stored byte: a larger count cannot decode (the decoders say so), do not size for it
        if (sig_method != S5B_COMPRESS_NONE && nn > counts[i]) nn = 0;
        ns[i] = (uint32_t)nn;
        sig_off[i] = sig_tot;
        sig_tot += round_up(nn, 8);
    }
    in_off[n] = tot;
    sig_off[n] = sig_tot;
    CU(ctx->h_stage_in.reserve(tot + 16));
    uint8_t *hin = static_cast<uint8_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i)
        if (in_len[i]) memcpy(hin + in_off[i], ptrs[i], in_len[i]);
    const size_t n1 = n + 1;
    CU(ctx->r_meta.reserve(4 * n1 * 8 + 6 * n * 4 + 256));
    CU(ctx->r_scratch.reserve(compact_scratch_bytes(n)));
    uint64_t *u64p = static_cast<uint64_t *>(ctx->r_meta.p);
    uint64_t *d_in_off = u64p, *d_sig_off = u64p + n1, *d_text_off = u64p + 2 * n1;
    uint32_t *u32p = reinterpret_cast<uint32_t *>(u64p + 4 * n1);
    uint32_t *d_in_len = u32p, *d_ns = u32p + n, *d_ns2 = u32p + 2 * n, *d_text_len = u32p + 3 * n;
    int32_t *d_status = reinterpret_cast<int32_t *>(u32p + 4 * n);
    CU(ctx->r_in.reserve(tot + 32));
    CU(cudaMemcpyAsync(ctx->r_in.p, hin, tot, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_off, in_off.data(), n1 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_sig_off, sig_off.data(), n1 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_len, in_len.data(), n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ns, ns.data(), n * 4, cudaMemcpyHostToDevice, st));
    const int16_t *d_sig;
    std::vector<int32_t> status(n, S5B_OK);
    if (sig_method == S5B_COMPRESS_NONE) {
        // the staged bytes are the sample slab (16-byte offsets = 8-sample offsets)
        for (size_t i = 0; i <= n; ++i) sig_off[i] = in_off[i] / 2;
        CU(cudaMemcpyAsync(d_sig_off, sig_off.data(), n1 * 8, cudaMemcpyHostToDevice, st));
        d_sig = static_cast<const int16_t *>(ctx->r_in.p);
    } else {
        CU(ctx->r_sig.reserve(sig_tot * 2 + 32));
        SvbDecodeArgs da{static_cast<const uint8_t *>(ctx->r_in.p), d_in_off, d_in_len, round_up(tot, 16), n,
                         static_cast<int16_t *>(ctx->r_sig.p), d_sig_off, d_ns2, d_status, ctx->slot[0].d_counter};
        if (sig_method == S5B_COMPRESS_SVB_ZD) CU(launch_svbzd_decode(da, ctx->num_sms, ctx->dec_bps, st));
        else CU(launch_exzd_decode(da, ctx->num_sms, ctx->xd_bps, st));
        ctx->launches += 1;
        CU(cudaMemcpyAsync(status.data(), d_status, n * 4, cudaMemcpyDeviceToHost, st));
        d_sig = static_cast<const int16_t *>(ctx->r_sig.p);
    }
    // ---- text sizes, offsets, text
    CU(launch_ascii_size(d_sig, d_sig_off, d_ns, n, d_text_len, st));
    CU(launch_scan(d_text_len, n, 1, d_text_off, ctx->r_scratch.p, st));
    std::vector<uint64_t> text_off(n1);
    CU(cudaMemcpyAsync(text_off.data(), d_text_off, n1 * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int first = S5B_OK;
    for (size_t i = 0; i < n; ++i)
        if (status[i] != S5B_OK && first == S5B_OK) first = status[i];
    if (first != S5B_OK) return first;
    const uint64_t text_tot = text_off[n];
    CU(ctx->r_img.reserve(text_tot + 64));
    CU(ctx->h_stage_out.reserve(text_tot + 64));
    CU(launch_ascii_format(d_sig, d_sig_off, d_ns, n, static_cast<uint8_t *>(ctx->r_img.p), d_text_off, st));
    ctx->launches += 5;
    CU(cudaMemcpyAsync(ctx->h_stage_out.p, ctx->r_img.p, text_tot, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const char *hout = static_cast<const char *>(ctx->h_stage_out.p);
    for (size_t i = 0; i < n; ++i) {
        const size_t len = (size_t)(text_off[i + 1] - text_off[i]);
        char *m = static_cast<char *>(malloc(len + 1));
        if (!m) return S5B_ERR_MEM;
        memcpy(m, hout + text_off[i], len);
        m[len] = '\0';
        out_ptrs[i] = m;
        out_n[i] = len;
    }
    return S5B_OK;
}

int s5b_ascii_to_signal_batch_host(s5b_ctx_t *ctx, const char *const *ptrs, const size_t *counts, const uint64_t *expect, size_t n,
                                   int16_t **out_ptrs, size_t *out_n) {
    if (!ctx || !ptrs || !counts || !expect || !out_ptrs || !out_n) return S5B_ERR_ARG;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = nullptr;
        out_n[i] = 0;
    }
    if (n == 0) return S5B_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->slot[0].stream;
    std::vector<uint64_t> in_off(n + 1), sig_off(n + 1);
    std::vector<uint32_t> in_len(n), ex(n);
    std::vector<char> pre_bad(n, 0);
    uint64_t tot = 0, sig_tot = 0;
    for (size_t i = 0; i < n; ++i) {
        if ((!ptrs[i] && counts[i]) || counts[i] > 0xfffffff0ull || expect[i] > 0x7ffffff0ull) return S5B_ERR_ARG;
        in_len[i] = (uint32_t)counts[i];
        ex[i] = (uint32_t)expect[i];
        // every sample but the last takes at least two characters ("d,"): a column shorter than that cannot hold them --
        // that record fails, and nothing is sized from its count
        if (expect[i] && counts[i] + 1 < 2 * expect[i]) {
            pre_bad[i] = 1;
            ex[i] = 0;
        }
        in_off[i] = tot;
        tot += round_up(counts[i], 16);
        sig_off[i] = sig_tot;
        sig_tot += round_up(ex[i], 8);
    }
    in_off[n] = tot;
    sig_off[n] = sig_tot;
    CU(ctx->h_stage_in.reserve(tot + 16));
    uint8_t *hin = static_cast<uint8_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i)
        if (in_len[i]) memcpy(hin + in_off[i], ptrs[i], in_len[i]);
    const size_t n1 = n + 1;
    CU(ctx->r_meta.reserve(2 * n1 * 8 + 4 * n * 4 + 256));
    uint64_t *u64p = static_cast<uint64_t *>(ctx->r_meta.p);
    uint64_t *d_in_off = u64p, *d_sig_off = u64p + n1;
    uint32_t *u32p = reinterpret_cast<uint32_t *>(u64p + 2 * n1);
    uint32_t *d_in_len = u32p, *d_ex = u32p + n, *d_ns = u32p + 2 * n;
    int32_t *d_status = reinterpret_cast<int32_t *>(u32p + 3 * n);
    CU(ctx->r_in.reserve(tot + 32));
    CU(ctx->r_sig.reserve(sig_tot * 2 + 32));
    CU(ctx->h_stage_out.reserve(sig_tot * 2 + 32));
    CU(cudaMemcpyAsync(ctx->r_in.p, hin, tot, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_off, in_off.data(), n1 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_sig_off, sig_off.data(), n1 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_len, in_len.data(), n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ex, ex.data(), n * 4, cudaMemcpyHostToDevice, st));
    CU(launch_ascii_parse(static_cast<const uint8_t *>(ctx->r_in.p), d_in_off, d_in_len, n, static_cast<int16_t *>(ctx->r_sig.p),
                          d_sig_off, d_ex, d_ns, d_status, st));
    ctx->launches += 1;
    std::vector<int32_t> status(n);
    CU(cudaMemcpyAsync(status.data(), d_status, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_stage_out.p, ctx->r_sig.p, sig_tot * 2, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int first = S5B_OK;
    const int16_t *hout = static_cast<const int16_t *>(ctx->h_stage_out.p);
    for (size_t i = 0; i < n; ++i) {
        if (pre_bad[i]) status[i] = S5B_ERR_ARG;
        if (status[i] != S5B_OK) {
            if (first == S5B_OK) first = status[i];
            continue;
        }
        const size_t bytes = (size_t)ex[i] * 2;
        int16_t *m = static_cast<int16_t *>(malloc(bytes ? bytes : 1));
        if (!m) return S5B_ERR_MEM;
        memcpy(m, hout + sig_off[i], bytes);
        out_ptrs[i] = m;
        out_n[i] = ex[i];
    }
    return first;
}

}  // extern "C"
