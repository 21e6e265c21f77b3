// s5b_capi.cu -- C-ABI (include/slow5b200.h) over the codec kernels, plus the CUDA-stream batch
// scheduler that replaces the reference's fork-join pthread pool for this path
// (src/thread.c:19-114 work_db/pthread_db/pthread_single; slow5lib/src/slow5_mt.c:202-316).
//
// Scheduler shape: a host batch is cut into sub-batches of ~S5B_CHUNK_MB of payload; sub-batch i runs
// on pipeline slot i % 2, each slot owning a stream, device slabs and pinned metadata, so the H2D copy
// of sub-batch i+1, the kernels of sub-batch i and the D2H copy of sub-batch i-1 overlap.  There is no
// CPU codec in this file: if CUDA is unusable every entry point reports S5B_ERR_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/slow5b200.h"
#include "s5b_kernels.h"
#include "s5b_ctx.h"
#include "zstd_core.h"

using namespace s5b;

namespace {
thread_local int tl_last_error = 0;
}  // namespace

extern "C" {

const char *s5b_version(void) { return "slow5b200 0.1.0 (sm_100a)"; }

const char *s5b_strerror(int err) {
    switch (err) {
        case S5B_OK: return "ok";
        case S5B_ERR_ARG: return "bad argument";
        case S5B_ERR_MEM: return "out of memory";
        case S5B_ERR_PRESS: return "malformed compressed stream";
        case S5B_ERR_NOSPACE: return "output slot too small";
        case S5B_ERR_DEVICE: return "CUDA device unavailable or CUDA error";
        case S5B_ERR_DATASET: return "a record's digitisation / sampling rate does not match the dataset";
        default: return "unknown error";
    }
}

int s5b_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

uint64_t s5b_svbzd_bound(uint32_t n) { return 4ull + ((uint64_t)n + 3) / 4 + 3ull * n; }
uint64_t s5b_svbzd_slot(uint32_t n) { return round_up(s5b_svbzd_bound(n), 16); }

int s5b_ctx_create(int device, s5b_ctx_t **out) {
    if (!out) return S5B_ERR_ARG;
    *out = nullptr;
    const bool timing = getenv("S5B_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    auto lap = [&](const char *what) {
        if (timing) {
            const double t1 = now();
            fprintf(stderr, "[timing]   ctx: %-28s %.3f s\n", what, t1 - t0);
            t0 = t1;
        }
    };
    int ndev = s5b_device_count();
    if (ndev <= 0) return S5B_ERR_DEVICE;
    lap("driver init / device count");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return S5B_ERR_DEVICE;
    }
    if (device >= ndev) return S5B_ERR_ARG;
    s5b_ctx *ctx = new (std::nothrow) s5b_ctx();
    if (!ctx) return S5B_ERR_MEM;
    ctx->device = device;
    DeviceGuard g(device);
    // (one attribute, not cudaGetDeviceProperties: that call fills in dozens of fields and takes tens of milliseconds)
    int n_sm = 0;
    if (!g.ok || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n_sm <= 0) {
        delete ctx;
        return S5B_ERR_DEVICE;
    }
    ctx->num_sms = n_sm;
    lap("set device / properties");
    ctx->enc_bps = svbzd_encode_blocks_per_sm();
    ctx->dec_bps = svbzd_decode_blocks_per_sm();
    ctx->inf_bps = inflate_blocks_per_sm();
    ctx->def_bps = deflate_blocks_per_sm();
    ctx->zd_bps = zstd_decode_blocks_per_sm();
    ctx->ze_bps = zstd_encode_blocks_per_sm();
    ctx->xe_bps = exzd_encode_blocks_per_sm();
    ctx->xd_bps = exzd_decode_blocks_per_sm();
    if (ctx->xe_bps <= 0 || ctx->xd_bps <= 0 || ctx->enc_bps <= 0 || ctx->dec_bps <= 0 || ctx->inf_bps <= 0 || ctx->def_bps <= 0 || ctx->zd_bps <= 0 ||
        ctx->ze_bps <= 0) {  // no sm_100a image for this device
        (void)cudaGetLastError();
        delete ctx;
        return S5B_ERR_DEVICE;
    }
    lap("kernel occupancy queries");
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_counter, 256) == cudaSuccess;
    for (int i = 0; ok && i < NSLOT; ++i) {
        ok = ok && cudaStreamCreateWithFlags(&ctx->slot[i].stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->slot[i].done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(&ctx->slot[i].d_counter, 256) == cudaSuccess;
    }
    lap("streams, events, counters");
    if (const char *e = getenv("S5B_CHUNK_MB")) {
        long v = atol(e);
        if (v > 0) ctx->chunk_bytes = (size_t)v << 20;
    }
    if (const char *e = getenv("S5B_RECODE_CHUNK")) {
        long v = atol(e);
        if (v > 0) ctx->recode_chunk_records = (size_t)v;
    }
    if (const char *e = getenv("S5B_RECODE_LANES")) {
        long v = atol(e);
        if (v >= 2 && v <= NLANE) ctx->n_lanes = (int)v;
    }
    if (const char *e = getenv("S5B_RECODE_DEV_CHUNK")) {
        long v = atol(e);
        if (v > 0) ctx->recode_dev_chunk_records = (size_t)v;
    }
    if (const char *e = getenv("S5B_RECODE_DEV_CHUNK_MB")) {
        long v = atol(e);
        if (v > 0) ctx->recode_dev_chunk_bytes = (size_t)v << 20;
    }
    if (const char *e = getenv("S5B_RECODE_CHUNK_MB")) {
        long v = atol(e);
        if (v > 0) ctx->recode_chunk_bytes = (size_t)v << 20;
    }
    if (!ok) {
        s5b_ctx_destroy(ctx);
        return S5B_ERR_DEVICE;
    }
    *out = ctx;
    return S5B_OK;
}

void s5b_ctx_destroy(s5b_ctx_t *ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < NSLOT; ++i) {
        PipeSlot &s = ctx->slot[i];
        s.d_a.release();
        s.d_b.release();
        s.d_c.release();
        s.d_meta.release();
        s.d_scratch.release();
        s.d_work.release();
        s.h_meta.release();
        if (s.d_counter) cudaFree(s.d_counter);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    ctx->d_scratch.release();
    ctx->zd_scratch.release();
    for (DevBuf *b : {&ctx->r_in, &ctx->r_infl, &ctx->r_sig, &ctx->r_svb, &ctx->r_packed, &ctx->r_z, &ctx->r_img, &ctx->r_meta,
                      &ctx->r_scratch, &ctx->r_work, &ctx->def_work, &ctx->inf_work})
        b->release();
    recode_lanes_release(ctx);
    ctx->h_stage_in.release();
    ctx->h_stage_out.release();
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    (void)cudaGetLastError();
    delete ctx;
}

const char *s5b_ctx_last_cuda_error(const s5b_ctx_t *ctx) { return ctx ? ctx->last_cuda_error.c_str() : ""; }
uint64_t s5b_ctx_launch_count(const s5b_ctx_t *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------
// device-resident batches
// ---------------------------------------------------------------------------------------------
int s5b_svbzd_encode_dev(s5b_ctx_t *ctx, const int16_t *d_sig, const uint64_t *d_sig_off, const uint32_t *d_n_samples,
                         uint64_t n_reads, uint8_t *d_svb, const uint64_t *d_svb_off, uint32_t *d_svb_len,
                         int32_t *d_status, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_sig || !d_sig_off || !d_n_samples || !d_svb || !d_svb_off || !d_svb_len || !d_status) return S5B_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_sig) & 15u) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    SvbEncodeArgs a{d_sig, d_sig_off, d_n_samples, n_reads, d_svb, d_svb_off, d_svb_len, d_status, ctx->d_counter};
    CU(launch_svbzd_encode(a, ctx->num_sms, ctx->enc_bps, st));
    ctx->launches += 1;
    return S5B_OK;
}

int s5b_svbzd_decode_dev(s5b_ctx_t *ctx, const uint8_t *d_svb, const uint64_t *d_svb_off, const uint32_t *d_svb_len,
                         uint64_t svb_capacity, uint64_t n_reads, int16_t *d_sig, const uint64_t *d_sig_off,
                         uint32_t *d_n_samples, int32_t *d_status, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_svb || !d_svb_off || !d_svb_len || !d_sig || !d_sig_off || !d_n_samples || !d_status) return S5B_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_svb) & 15u) || (reinterpret_cast<uintptr_t>(d_sig) & 15u) || (svb_capacity & 15u))
        return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    // the decode kernel's work counter must not alias an encode still in flight on another stream
    SvbDecodeArgs a{d_svb, d_svb_off, d_svb_len, svb_capacity, n_reads, d_sig, d_sig_off, d_n_samples, d_status,
                    ctx->d_counter + 8};
    CU(launch_svbzd_decode(a, ctx->num_sms, ctx->dec_bps, st));
    ctx->launches += 1;
    return S5B_OK;
}

int s5b_svbzd_peek_dev(s5b_ctx_t *ctx, const uint8_t *d_svb, const uint64_t *d_svb_off, const uint32_t *d_svb_len,
                       uint64_t n_reads, uint32_t *d_n_samples, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_svb || !d_svb_off || !d_svb_len || !d_n_samples) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    CU(launch_svbzd_peek(d_svb, d_svb_off, d_svb_len, n_reads, d_n_samples, st));
    ctx->launches += 1;
    return S5B_OK;
}

uint64_t s5b_exzd_bound(uint32_t n) { return exzd_bound(n); }
uint64_t s5b_exzd_slot(uint32_t n) { return round_up(exzd_bound(n), 16); }

int s5b_exzd_encode_dev(s5b_ctx_t *ctx, const int16_t *d_sig, const uint64_t *d_sig_off, const uint32_t *d_n_samples,
                        uint64_t n_reads, uint8_t *d_out, const uint64_t *d_out_off, uint32_t *d_out_len,
                        int32_t *d_status, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_sig || !d_sig_off || !d_n_samples || !d_out || !d_out_off || !d_out_len || !d_status) return S5B_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_sig) & 15u) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    SvbEncodeArgs a{d_sig, d_sig_off, d_n_samples, n_reads, d_out, d_out_off, d_out_len, d_status, ctx->d_counter + 24};
    CU(launch_exzd_encode(a, ctx->num_sms, ctx->xe_bps, st));
    ctx->launches += 1;
    return S5B_OK;
}

int s5b_exzd_decode_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                        uint64_t in_capacity, uint64_t n_reads, int16_t *d_sig, const uint64_t *d_sig_off,
                        uint32_t *d_n_samples, int32_t *d_status, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_in || !d_in_off || !d_in_len || !d_sig || !d_sig_off || !d_n_samples || !d_status) return S5B_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_sig) & 15u) || (reinterpret_cast<uintptr_t>(d_in) & 15u) || (in_capacity & 15u))
        return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    SvbDecodeArgs a{d_in, d_in_off, d_in_len, in_capacity, n_reads, d_sig, d_sig_off, d_n_samples, d_status,
                    ctx->d_counter + 28};
    CU(launch_exzd_decode(a, ctx->num_sms, ctx->xd_bps, st));
    ctx->launches += 1;
    return S5B_OK;
}

int s5b_zlib_inflate_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                         uint64_t in_capacity, uint64_t n_reads, uint8_t *d_out, const uint64_t *d_out_off,
                         uint32_t *d_out_len, int32_t *d_status, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_in || !d_in_off || !d_in_len || !d_out || !d_out_off || !d_out_len || !d_status) return S5B_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (in_capacity & 15u)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    InflateArgs a{d_in, d_in_off, d_in_len, in_capacity, n_reads, d_out, d_out_off, d_out_len, d_status,
                  ctx->d_counter + 16};
    CU(launch_inflate_ws(ctx->inf_work, a, ctx->num_sms, ctx->inf_bps, st));
    ctx->launches += 1;
    return S5B_OK;
}

int s5b_zstd_content_size(const void *frame, size_t len, uint64_t *size) {
    if (!frame || !size) return S5B_ERR_ARG;
    s5bz::FrameInfo fi;
    if (s5bz::parse_frame_header(static_cast<const uint8_t *>(frame), len, fi) != s5bz::Z_OK || !fi.has_content_size)
        return S5B_ERR_PRESS;
    *size = fi.content_size;
    return S5B_OK;
}

}  // extern "C"
int s5b::zstd_launch(s5b_ctx *ctx, const InflateArgs &a, cudaStream_t st) {
    CU(ctx->zd_scratch.reserve(zstd_decode_scratch_bytes(ctx->num_sms, ctx->zd_bps)));
    CU(launch_zstd_decode(a, ctx->num_sms, ctx->zd_bps, ctx->zd_scratch.p, st));
    ctx->launches += 1;
    return S5B_OK;
}
extern "C" {

int s5b_zstd_decode_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                        uint64_t in_capacity, uint64_t n_reads, uint8_t *d_out, const uint64_t *d_out_off,
                        uint32_t *d_out_len, int32_t *d_status, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_in || !d_in_off || !d_in_len || !d_out || !d_out_off || !d_out_len || !d_status) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    InflateArgs a{d_in, d_in_off, d_in_len, in_capacity, n_reads, d_out, d_out_off, d_out_len, d_status,
                  ctx->d_counter + 28};
    return zstd_launch(ctx, a, st);
}

uint64_t s5b_zlib_bound(uint64_t len) { return round_up(deflate_bound(len), 16); }

int s5b_zlib_deflate_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                         uint64_t in_capacity, const uint32_t *d_split, uint64_t n_reads, uint8_t *d_out,
                         const uint64_t *d_out_off, uint32_t *d_out_len, int32_t *d_status, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_in || !d_in_off || !d_in_len || !d_out || !d_out_off || !d_out_len || !d_status) return S5B_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (in_capacity & 15u)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    DeflateArgs a{d_in, d_in_off, d_in_len, in_capacity, d_split, n_reads, d_out, d_out_off, d_out_len, d_status,
                  ctx->d_counter + 24};
    CU(launch_deflate_ws(ctx->def_work, a, ctx->num_sms, ctx->def_bps, st));
    ctx->launches += 7;
    return S5B_OK;
}

uint64_t s5b_zstd_bound(uint64_t len) { return round_up(zstd_encode_bound(len), 16); }

int s5b_zstd_encode_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                        uint64_t in_capacity, const uint32_t *d_split, uint64_t n_reads, uint8_t *d_out,
                        const uint64_t *d_out_off, uint32_t *d_out_len, int32_t *d_status, void *stream) {
    if (!ctx) return S5B_ERR_ARG;
    if (n_reads == 0) return S5B_OK;
    if (!d_in || !d_in_off || !d_in_len || !d_out || !d_out_off || !d_out_len || !d_status) return S5B_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (in_capacity & 15u)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    DeflateArgs a{d_in, d_in_off, d_in_len, in_capacity, d_split, n_reads, d_out, d_out_off, d_out_len, d_status,
                  ctx->d_counter + 20};
    CU(launch_zstd_encode(a, ctx->num_sms, ctx->ze_bps, st));
    ctx->launches += 1;
    return S5B_OK;
}

int s5b_compact_dev(s5b_ctx_t *ctx, const uint8_t *d_src, const uint64_t *d_src_off, const uint32_t *d_len,
                    uint64_t n_reads, uint32_t align, uint8_t *d_dst, uint64_t *d_dst_off, void *stream) {
    if (!ctx || !d_dst_off) return S5B_ERR_ARG;
    if (align != 1 && align != 16) return S5B_ERR_ARG;
    if (n_reads && (!d_src || !d_src_off || !d_len || !d_dst)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    CU(ctx->d_scratch.reserve(compact_scratch_bytes(n_reads)));
    int nl = 0;
    CU(launch_compact(d_src, d_src_off, d_len, n_reads, align, d_dst, d_dst_off, ctx->d_scratch.p, st, &nl));
    ctx->launches += nl;
    return S5B_OK;
}

// ---------------------------------------------------------------------------------------------
// host batches: slab form
// ---------------------------------------------------------------------------------------------
// Cuts [0, n) into runs whose payload stays under `budget` bytes (at least one read per run).
static void cut_runs(const uint64_t *off, uint64_t n, uint64_t unit, uint64_t budget,
                     std::vector<uint64_t> &cuts) {
    cuts.clear();
    cuts.push_back(0);
    uint64_t start = 0;
    for (uint64_t r = 1; r <= n; ++r) {
        if (r == n) break;
        if ((off[r + 1] - off[start]) * unit > budget && r > start) {
            cuts.push_back(r);
            start = r;
        }
    }
    cuts.push_back(n);
}

int s5b_svbzd_encode_host(s5b_ctx_t *ctx, const int16_t *h_sig, const uint64_t *h_sig_off, const uint32_t *h_n_samples,
                          uint64_t n_reads, uint8_t *h_svb, uint64_t h_svb_capacity, uint64_t *h_svb_off,
                          uint32_t *h_svb_len, int32_t *h_status) {
    if (!ctx || !h_svb_off) return S5B_ERR_ARG;
    h_svb_off[0] = 0;
    if (n_reads == 0) return S5B_OK;
    if (!h_sig || !h_sig_off || !h_n_samples || !h_svb || !h_svb_len || !h_status) return S5B_ERR_ARG;
    for (uint64_t r = 0; r < n_reads; ++r) {
        if ((h_sig_off[r] & 7) || h_sig_off[r + 1] < h_sig_off[r] || h_sig_off[r + 1] - h_sig_off[r] < h_n_samples[r])
            return S5B_ERR_ARG;
    }
    DeviceGuard g(ctx->device);
    std::vector<uint64_t> cuts;
    cut_runs(h_sig_off, n_reads, 2, ctx->chunk_bytes, cuts);
    const size_t nrun = cuts.size() - 1;

    struct Pending {
        bool active = false;
        uint64_t r0 = 0, r1 = 0;
        uint64_t host_base = 0;
    } pend[NSLOT];
    uint64_t host_total = 0;
    int first_err = S5B_OK;

    // finish(): wait for the slot's kernels, learn the dense size, queue the payload D2H, fix up offsets
    auto finish = [&](int si) -> int {
        PipeSlot &s = ctx->slot[si];
        Pending &p = pend[si];
        if (!p.active) return S5B_OK;
        p.active = false;
        const uint64_t m = p.r1 - p.r0;
        CU(cudaEventSynchronize(s.done));
        // pinned meta layout (written by the D2H below): [dense_off (m+1) u64][len m u32][status m i32]
        const uint64_t *dense_off = static_cast<const uint64_t *>(s.h_meta.p);
        const uint32_t *len = reinterpret_cast<const uint32_t *>(dense_off + (m + 1));
        const int32_t *status = reinterpret_cast<const int32_t *>(len + m);
        const uint64_t dense_total = dense_off[m];
        if (host_total + dense_total > h_svb_capacity) return S5B_ERR_NOSPACE;
        CU(cudaMemcpyAsync(h_svb + host_total, s.d_c.p, dense_total, cudaMemcpyDeviceToHost, s.stream));
        for (uint64_t i = 0; i < m; ++i) {
            h_svb_off[p.r0 + i] = host_total + dense_off[i];
            h_svb_len[p.r0 + i] = len[i];
            h_status[p.r0 + i] = status[i];
            if (status[i] != S5B_OK && first_err == S5B_OK) first_err = status[i];
        }
        host_total += dense_total;
        h_svb_off[p.r1] = host_total;
        CU(cudaEventRecord(s.done, s.stream));
        return S5B_OK;
    };

    for (size_t run = 0; run < nrun; ++run) {
        const int si = (int)(run % NSLOT);
        PipeSlot &s = ctx->slot[si];
        int rc = finish(si);
        if (rc != S5B_OK) return rc;
        CU(cudaEventSynchronize(s.done));  // previous payload D2H of this slot has landed; buffers are free
        const uint64_t r0 = cuts[run], r1 = cuts[run + 1], m = r1 - r0;
        const uint64_t s0 = h_sig_off[r0], s1 = h_sig_off[r1];
        const uint64_t sig_bytes = round_up((s1 - s0) * 2, 16);
        // host-side metadata for this run, rebased to the run's slabs
        // layout in h_meta (up): [sig_off (m+1) u64][slot_off (m+1) u64][n m u32]
        const size_t up_bytes = (2 * (m + 1)) * 8 + m * 4;
        const size_t down_bytes = (m + 1) * 8 + m * 8;
        CU(s.h_meta.reserve(up_bytes > down_bytes ? up_bytes : down_bytes));
        uint64_t *u_sig_off = static_cast<uint64_t *>(s.h_meta.p);
        uint64_t *u_slot_off = u_sig_off + (m + 1);
        uint32_t *u_n = reinterpret_cast<uint32_t *>(u_slot_off + (m + 1));
        uint64_t slot_total = 0;
        for (uint64_t i = 0; i < m; ++i) {
            u_sig_off[i] = h_sig_off[r0 + i] - s0;
            u_slot_off[i] = slot_total;
            u_n[i] = h_n_samples[r0 + i];
            slot_total += s5b_svbzd_slot(u_n[i]);
        }
        u_sig_off[m] = sig_bytes / 2;  // padded: the last read may be bulk-copied in whole granules
        u_slot_off[m] = slot_total;
        // device meta layout: [sig_off][slot_off][n] | [dense_off (m+1) u64][len m u32][status m i32]
        const size_t d_down_at = round_up(up_bytes, 16);
        CU(s.d_meta.reserve(d_down_at + down_bytes));
        CU(s.d_a.reserve(sig_bytes + 16));
        CU(s.d_b.reserve(slot_total + 16));
        CU(s.d_c.reserve(slot_total + 16));
        CU(s.d_scratch.reserve(compact_scratch_bytes(m)));
        uint8_t *dm = static_cast<uint8_t *>(s.d_meta.p);
        uint64_t *d_sig_off = reinterpret_cast<uint64_t *>(dm);
        uint64_t *d_slot_off = d_sig_off + (m + 1);
        uint32_t *d_n = reinterpret_cast<uint32_t *>(d_slot_off + (m + 1));
        uint64_t *d_dense_off = reinterpret_cast<uint64_t *>(dm + d_down_at);
        uint32_t *d_len = reinterpret_cast<uint32_t *>(d_dense_off + (m + 1));
        int32_t *d_status = reinterpret_cast<int32_t *>(d_len + m);

        CU(cudaMemcpyAsync(dm, s.h_meta.p, up_bytes, cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.d_a.p, h_sig + s0, (s1 - s0) * 2, cudaMemcpyHostToDevice, s.stream));
        SvbEncodeArgs a{static_cast<const int16_t *>(s.d_a.p), d_sig_off, d_n, m, static_cast<uint8_t *>(s.d_b.p),
                        d_slot_off, d_len, d_status, s.d_counter};
        CU(launch_svbzd_encode(a, ctx->num_sms, ctx->enc_bps, s.stream));
        int nl = 0;
        CU(launch_compact(static_cast<const uint8_t *>(s.d_b.p), d_slot_off, d_len, m, 16,
                          static_cast<uint8_t *>(s.d_c.p), d_dense_off, s.d_scratch.p, s.stream, &nl));
        ctx->launches += 1 + nl;
        // the up-metadata in h_meta has been consumed by the H2D above once the stream reaches here;
        // the same pinned block receives the down-metadata
        CU(cudaMemcpyAsync(s.h_meta.p, d_dense_off, down_bytes, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaEventRecord(s.done, s.stream));
        pend[si].active = true;
        pend[si].r0 = r0;
        pend[si].r1 = r1;
    }
    // drain in submission order so host offsets stay monotone
    for (size_t k = 0; k < NSLOT; ++k) {
        const int si = (int)((nrun + k) % NSLOT);
        int rc = finish(si);
        if (rc != S5B_OK) return rc;
    }
    for (int si = 0; si < NSLOT; ++si) CU(cudaEventSynchronize(ctx->slot[si].done));
    return first_err;
}

int s5b_svbzd_decode_host(s5b_ctx_t *ctx, const uint8_t *h_svb, const uint64_t *h_svb_off, const uint32_t *h_svb_len,
                          uint64_t n_reads, int16_t *h_sig, uint64_t h_sig_capacity, uint64_t *h_sig_off,
                          uint32_t *h_n_samples, int32_t *h_status) {
    if (!ctx || !h_sig_off) return S5B_ERR_ARG;
    h_sig_off[0] = 0;
    if (n_reads == 0) return S5B_OK;
    if (!h_svb || !h_svb_off || !h_svb_len || !h_sig || !h_n_samples || !h_status) return S5B_ERR_ARG;
    // sample counts come from the stream headers (slow5_press.c:1120); the signal layout follows from them
    uint64_t sig_total = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        if (h_svb_off[r + 1] < h_svb_off[r] || h_svb_off[r + 1] - h_svb_off[r] < h_svb_len[r]) return S5B_ERR_ARG;
        uint32_t n = 0;
        if (h_svb_len[r] >= 4) memcpy(&n, h_svb + h_svb_off[r], 4);
        h_n_samples[r] = n;
        h_sig_off[r] = sig_total;
        sig_total += round_up(n, 8);
    }
    h_sig_off[n_reads] = sig_total;
    if (sig_total > h_sig_capacity) return S5B_ERR_NOSPACE;
    DeviceGuard g(ctx->device);
    // cut by output payload (the larger side)
    std::vector<uint64_t> cuts;
    cut_runs(h_sig_off, n_reads, 2, ctx->chunk_bytes, cuts);
    const size_t nrun = cuts.size() - 1;
    struct Pending {
        bool active = false;
        uint64_t r0 = 0, r1 = 0;
    } pend[NSLOT];
    int first_err = S5B_OK;
    auto finish = [&](int si) -> int {
        PipeSlot &s = ctx->slot[si];
        Pending &p = pend[si];
        if (!p.active) return S5B_OK;
        p.active = false;
        CU(cudaEventSynchronize(s.done));
        const uint64_t m = p.r1 - p.r0;
        // pinned down-meta: [n m u32][status m i32]
        const uint32_t *n = static_cast<const uint32_t *>(s.h_meta.p);
        const int32_t *status = reinterpret_cast<const int32_t *>(n + m);
        for (uint64_t i = 0; i < m; ++i) {
            h_status[p.r0 + i] = status[i];
            if (status[i] != S5B_OK && first_err == S5B_OK) first_err = status[i];
        }
        return S5B_OK;
    };
    for (size_t run = 0; run < nrun; ++run) {
        const int si = (int)(run % NSLOT);
        PipeSlot &s = ctx->slot[si];
        int rc = finish(si);
        if (rc != S5B_OK) return rc;
        const uint64_t r0 = cuts[run], r1 = cuts[run + 1], m = r1 - r0;
        // input byte range, widened to 16-byte granules of the HOST slab offsets so relative alignment
        // of every stream is preserved on the device
        const uint64_t b0 = h_svb_off[r0] & ~15ull;
        const uint64_t b1 = h_svb_off[r1 - 1] + h_svb_len[r1 - 1];
        const uint64_t in_bytes = round_up(b1 - b0, 16);
        const uint64_t s0 = h_sig_off[r0], s1 = h_sig_off[r1];
        const size_t up_bytes = 2 * (m + 1) * 8 + m * 4;
        const size_t down_bytes = m * 8;
        CU(s.h_meta.reserve(up_bytes > down_bytes ? up_bytes : down_bytes));
        uint64_t *u_svb_off = static_cast<uint64_t *>(s.h_meta.p);
        uint64_t *u_sig_off = u_svb_off + (m + 1);
        uint32_t *u_len = reinterpret_cast<uint32_t *>(u_sig_off + (m + 1));
        for (uint64_t i = 0; i < m; ++i) {
            u_svb_off[i] = h_svb_off[r0 + i] - b0;
            u_sig_off[i] = h_sig_off[r0 + i] - s0;
            u_len[i] = h_svb_len[r0 + i];
        }
        u_svb_off[m] = in_bytes;
        u_sig_off[m] = s1 - s0;
        const size_t d_down_at = round_up(up_bytes, 16);
        CU(s.d_meta.reserve(d_down_at + down_bytes));
        CU(s.d_a.reserve(in_bytes + 16));
        CU(s.d_b.reserve((s1 - s0) * 2 + 16));
        uint8_t *dm = static_cast<uint8_t *>(s.d_meta.p);
        uint64_t *d_svb_off = reinterpret_cast<uint64_t *>(dm);
        uint64_t *d_sig_off = d_svb_off + (m + 1);
        uint32_t *d_len = reinterpret_cast<uint32_t *>(d_sig_off + (m + 1));
        uint32_t *d_n = reinterpret_cast<uint32_t *>(dm + d_down_at);
        int32_t *d_status = reinterpret_cast<int32_t *>(d_n + m);
        CU(cudaMemcpyAsync(dm, s.h_meta.p, up_bytes, cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.d_a.p, h_svb + b0, b1 - b0, cudaMemcpyHostToDevice, s.stream));
        SvbDecodeArgs a{static_cast<const uint8_t *>(s.d_a.p), d_svb_off, d_len, in_bytes, m,
                        static_cast<int16_t *>(s.d_b.p), d_sig_off, d_n, d_status, s.d_counter};
        CU(launch_svbzd_decode(a, ctx->num_sms, ctx->dec_bps, s.stream));
        ctx->launches += 1;
        CU(cudaMemcpyAsync(h_sig + s0, s.d_b.p, (s1 - s0) * 2, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(s.h_meta.p, d_n, down_bytes, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaEventRecord(s.done, s.stream));
        pend[si].active = true;
        pend[si].r0 = r0;
        pend[si].r1 = r1;
    }
    for (size_t k = 0; k < NSLOT; ++k) {
        int rc = finish((int)((nrun + k) % NSLOT));
        if (rc != S5B_OK) return rc;
    }
    return first_err;
}

// ---------------------------------------------------------------------------------------------
// host batches: pointer-array form (db_t / slow5_batch_t shape)
// ---------------------------------------------------------------------------------------------
static int svbzd_compress_ptrs(s5b_ctx_t *ctx, const void *const *ptrs, const size_t *counts, size_t n,
                               void **out_ptrs, size_t *out_n) {
    std::vector<uint64_t> sig_off(n + 1), svb_off(n + 1);
    std::vector<uint32_t> ns(n), lens(n);
    std::vector<int32_t> status(n);
    uint64_t tot = 0, bound = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!ptrs[i] && counts[i]) return S5B_ERR_ARG;
        if (counts[i] / 2 > 0xffffffffull) return S5B_ERR_ARG;
        ns[i] = (uint32_t)(counts[i] / 2);  // count is BYTES, slow5_press.c:1088
        sig_off[i] = tot;
        tot += round_up(ns[i], 8);
        bound += s5b_svbzd_slot(ns[i]);
    }
    sig_off[n] = tot;
    CU(ctx->h_stage_in.reserve(tot * 2 + 16));
    CU(ctx->h_stage_out.reserve(bound + 16));
    int16_t *hin = static_cast<int16_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i) memcpy(hin + sig_off[i], ptrs[i], (size_t)ns[i] * 2);
    int rc = s5b_svbzd_encode_host(ctx, hin, sig_off.data(), ns.data(), n, static_cast<uint8_t *>(ctx->h_stage_out.p),
                                   bound + 16, svb_off.data(), lens.data(), status.data());
    if (rc != S5B_OK && rc != S5B_ERR_PRESS && rc != S5B_ERR_NOSPACE && rc != S5B_ERR_ARG) return rc;
    const uint8_t *hout = static_cast<const uint8_t *>(ctx->h_stage_out.p);
    int first = S5B_OK;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = nullptr;
        out_n[i] = 0;
        if (status[i] != S5B_OK) {
            if (first == S5B_OK) first = status[i];
            continue;
        }
        void *m = malloc(lens[i] ? lens[i] : 1);
        if (!m) {
            if (first == S5B_OK) first = S5B_ERR_MEM;
            continue;
        }
        memcpy(m, hout + svb_off[i], lens[i]);
        out_ptrs[i] = m;
        out_n[i] = lens[i];
    }
    return first;
}

static int svbzd_depress_ptrs(s5b_ctx_t *ctx, const void *const *ptrs, const size_t *counts, size_t n,
                              void **out_ptrs, size_t *out_n) {
    std::vector<uint64_t> svb_off(n + 1), sig_off(n + 1);
    std::vector<uint32_t> lens(n), ns(n);
    std::vector<int32_t> status(n);
    uint64_t tot = 0, sig_tot = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!ptrs[i] && counts[i]) return S5B_ERR_ARG;
        if (counts[i] > 0xffffffffull) return S5B_ERR_ARG;
        lens[i] = (uint32_t)counts[i];
        svb_off[i] = tot;
        tot += round_up(lens[i], 16);
        uint32_t nn = 0;
        if (counts[i] >= 4) memcpy(&nn, ptrs[i], 4);
        sig_tot += round_up(nn, 8);
    }
    svb_off[n] = tot;
    CU(ctx->h_stage_in.reserve(tot + 16));
    CU(ctx->h_stage_out.reserve(sig_tot * 2 + 16));
    uint8_t *hin = static_cast<uint8_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i) memcpy(hin + svb_off[i], ptrs[i], lens[i]);
    int rc = s5b_svbzd_decode_host(ctx, hin, svb_off.data(), lens.data(), n, static_cast<int16_t *>(ctx->h_stage_out.p),
                                   sig_tot, sig_off.data(), ns.data(), status.data());
    if (rc != S5B_OK && rc != S5B_ERR_PRESS && rc != S5B_ERR_NOSPACE && rc != S5B_ERR_ARG) return rc;
    const int16_t *hout = static_cast<const int16_t *>(ctx->h_stage_out.p);
    int first = S5B_OK;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = nullptr;
        out_n[i] = 0;
        if (status[i] != S5B_OK) {
            if (first == S5B_OK) first = status[i];
            continue;
        }
        const size_t bytes = (size_t)ns[i] * 2;
        void *m = malloc(bytes ? bytes : 1);
        if (!m) {
            if (first == S5B_OK) first = S5B_ERR_MEM;
            continue;
        }
        memcpy(m, hout + sig_off[i], bytes);
        out_ptrs[i] = m;
        out_n[i] = bytes;
    }
    return first;
}

// ex-zd: one shot (H2D, kernel, D2H) on slot 0's stream; the stream header carries the sample count (:1790-1793)
static int exzd_ptrs(s5b_ctx_t *ctx, bool compress, const void *const *ptrs, const size_t *counts, size_t n,
                     void **out_ptrs, size_t *out_n) {
    std::vector<uint64_t> in_off(n + 1), out_off(n + 1);
    std::vector<uint32_t> in_len(n);
    uint64_t in_tot = 0, out_tot = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!ptrs[i] && counts[i]) return S5B_ERR_ARG;
        in_off[i] = in_tot;
        out_off[i] = out_tot;
        if (compress) {
            if (counts[i] / 2 > 0xffffffffull) return S5B_ERR_ARG;
            in_len[i] = (uint32_t)(counts[i] / 2);  // samples; count is BYTES (:1723)
            in_tot += round_up(in_len[i], 8) * 2;
            out_tot += s5b_exzd_slot(in_len[i]);
        } else {
            if (counts[i] > 0xffffffffull) return S5B_ERR_ARG;
            in_len[i] = (uint32_t)counts[i];
            in_tot += round_up(in_len[i], 16);
            uint64_t nin = 0;
            if (counts[i] >= 9) memcpy(&nin, static_cast<const uint8_t *>(ptrs[i]) + 1, 8);
            if (nin > 0xffffffffull) nin = 0;  // the kernel reports the bad header
            out_tot += round_up(nin, 8) * 2;
        }
    }
    in_off[n] = in_tot;
    out_off[n] = out_tot;
    DeviceGuard g(ctx->device);
    PipeSlot &s = ctx->slot[0];
    CU(ctx->h_stage_in.reserve(in_tot + 16));
    CU(ctx->h_stage_out.reserve(out_tot + 16));
    uint8_t *hin = static_cast<uint8_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i) memcpy(hin + in_off[i], ptrs[i], counts[i]);
    // device meta: [in_off (n+1) u64][out_off (n+1) u64][in_len n u32][out_len n u32][status n i32]
    const size_t meta_bytes = 2 * (n + 1) * 8 + 3 * n * 4;
    CU(s.h_meta.reserve(meta_bytes));
    CU(s.d_meta.reserve(meta_bytes));
    CU(s.d_a.reserve(in_tot + 16));
    CU(s.d_b.reserve(out_tot + 16));
    uint64_t *m_in_off = static_cast<uint64_t *>(s.h_meta.p), *m_out_off = m_in_off + (n + 1);
    uint32_t *m_in_len = reinterpret_cast<uint32_t *>(m_out_off + (n + 1)), *m_out_len = m_in_len + n;
    int32_t *m_status = reinterpret_cast<int32_t *>(m_out_len + n);
    for (size_t i = 0; i <= n; ++i) {
        m_in_off[i] = compress ? in_off[i] / 2 : in_off[i];     // samples on the signal side
        m_out_off[i] = compress ? out_off[i] : out_off[i] / 2;
    }
    for (size_t i = 0; i < n; ++i) m_in_len[i] = in_len[i];
    uint8_t *dm = static_cast<uint8_t *>(s.d_meta.p);
    uint64_t *d_in_off = reinterpret_cast<uint64_t *>(dm), *d_out_off = d_in_off + (n + 1);
    uint32_t *d_in_len = reinterpret_cast<uint32_t *>(d_out_off + (n + 1)), *d_out_len = d_in_len + n;
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    CU(cudaMemcpyAsync(dm, s.h_meta.p, 2 * (n + 1) * 8 + n * 4, cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpyAsync(s.d_a.p, hin, in_tot, cudaMemcpyHostToDevice, s.stream));
    if (compress) {
        SvbEncodeArgs a{static_cast<const int16_t *>(s.d_a.p), d_in_off, d_in_len, n, static_cast<uint8_t *>(s.d_b.p),
                        d_out_off, d_out_len, d_status, s.d_counter};
        CU(launch_exzd_encode(a, ctx->num_sms, ctx->xe_bps, s.stream));
    } else {
        SvbDecodeArgs a{static_cast<const uint8_t *>(s.d_a.p), d_in_off, d_in_len, round_up(in_tot, 16), n,
                        static_cast<int16_t *>(s.d_b.p), d_out_off, d_out_len, d_status, s.d_counter};
        CU(launch_exzd_decode(a, ctx->num_sms, ctx->xd_bps, s.stream));
    }
    ctx->launches += 1;
    CU(cudaMemcpyAsync(ctx->h_stage_out.p, s.d_b.p, out_tot, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaMemcpyAsync(m_out_len, d_out_len, 2 * n * 4, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    const uint8_t *hout = static_cast<const uint8_t *>(ctx->h_stage_out.p);
    int first = S5B_OK;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = nullptr;
        out_n[i] = 0;
        if (m_status[i] != S5B_OK) {
            if (first == S5B_OK) first = m_status[i];
            continue;
        }
        const size_t bytes = compress ? (size_t)m_out_len[i] : (size_t)m_out_len[i] * 2;  // decode reports samples
        void *m = malloc(bytes ? bytes : 1);
        if (!m) {
            if (first == S5B_OK) first = S5B_ERR_MEM;
            continue;
        }
        memcpy(m, hout + out_off[i], bytes);
        out_ptrs[i] = m;
        out_n[i] = bytes;
    }
    return first;
}

// slow5_arr_qts_round for a batch of host arrays: packed back to back (even offsets), one H2D, one launch over the whole slab,
// one D2H
static int qts_ptrs(s5b_ctx_t *ctx, int bits, const void *const *ptrs, const size_t *counts, size_t n, void **out_ptrs,
                    size_t *out_n) {
    std::vector<uint64_t> off(n + 1);
    uint64_t tot = 0;
    for (size_t i = 0; i < n; ++i) {
        if ((!ptrs[i] && counts[i]) || (counts[i] & 1)) return S5B_ERR_ARG;
        off[i] = tot;
        tot += counts[i];
    }
    off[n] = tot;
    DeviceGuard g(ctx->device);
    PipeSlot &s = ctx->slot[0];
    CU(ctx->h_stage_in.reserve(tot + 16));
    uint8_t *h = static_cast<uint8_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i) memcpy(h + off[i], ptrs[i], counts[i]);
    CU(s.d_a.reserve(tot + 16));
    CU(cudaMemcpyAsync(s.d_a.p, h, tot, cudaMemcpyHostToDevice, s.stream));
    if (tot) {
        CU(launch_qts_round(static_cast<int16_t *>(s.d_a.p), tot / 2, nullptr, bits, ctx->num_sms, s.stream));
        ctx->launches += 1;
    }
    CU(cudaMemcpyAsync(h, s.d_a.p, tot, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    int first = S5B_OK;
    for (size_t i = 0; i < n; ++i) {
        out_n[i] = 0;
        out_ptrs[i] = malloc(counts[i] ? counts[i] : 1);
        if (!out_ptrs[i]) {
            if (first == S5B_OK) first = S5B_ERR_MEM;
            continue;
        }
        memcpy(out_ptrs[i], h + off[i], counts[i]);
        out_n[i] = counts[i];
    }
    return first;
}

// zlib streams: the inflated size is not stored anywhere (slow5_press.c:985-1003 grows its buffer in 256 KiB
// steps), so slots are sized from a guess and the (rare) streams that overflow are run again with the exact
// size the first pass reported.
static int zlib_depress_ptrs(s5b_ctx_t *ctx, const void *const *ptrs, const size_t *counts, size_t n,
                             void **out_ptrs, size_t *out_n) {
    std::vector<uint64_t> in_off(n + 1), out_off(n + 1);
    std::vector<uint32_t> in_len(n), out_len(n);
    std::vector<int32_t> status(n);
    uint64_t tot = 0, otot = 0;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = nullptr;
        out_n[i] = 0;
        if (!ptrs[i] && counts[i]) return S5B_ERR_ARG;
        if (counts[i] > 0xffffffffull) return S5B_ERR_ARG;
        in_len[i] = (uint32_t)counts[i];
        in_off[i] = tot;
        tot += round_up(in_len[i], 16);
        out_off[i] = otot;
        otot += round_up(4ull * in_len[i] + 1024, 16);
    }
    in_off[n] = tot;
    out_off[n] = otot;
    PipeSlot &s = ctx->slot[0];
    cudaStream_t st = s.stream;
    CU(ctx->h_stage_in.reserve(tot + 16));
    uint8_t *hin = static_cast<uint8_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i) memcpy(hin + in_off[i], ptrs[i], in_len[i]);
    const size_t meta_bytes = 2 * (n + 1) * 8 + 2 * n * 4;
    CU(s.d_meta.reserve(meta_bytes + 64));
    CU(s.d_a.reserve(tot + 16));
    CU(s.d_b.reserve(otot + 16));
    uint64_t *d_in_off = static_cast<uint64_t *>(s.d_meta.p);
    uint64_t *d_out_off = d_in_off + (n + 1);
    uint32_t *d_in_len = reinterpret_cast<uint32_t *>(d_out_off + (n + 1));
    uint32_t *d_out_len = d_in_len + n;
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    CU(s.d_meta.reserve(meta_bytes + n * 4 + 64));
    d_in_off = static_cast<uint64_t *>(s.d_meta.p);
    d_out_off = d_in_off + (n + 1);
    d_in_len = reinterpret_cast<uint32_t *>(d_out_off + (n + 1));
    d_out_len = d_in_len + n;
    d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    CU(cudaMemcpyAsync(s.d_a.p, hin, tot, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_off, in_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_out_off, out_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_len, in_len.data(), n * 4, cudaMemcpyHostToDevice, st));
    InflateArgs a{static_cast<const uint8_t *>(s.d_a.p), d_in_off, d_in_len, round_up(tot, 16), n,
                  static_cast<uint8_t *>(s.d_b.p), d_out_off, d_out_len, d_status, s.d_counter};
    CU(launch_inflate_ws(ctx->inf_work, a, ctx->num_sms, ctx->inf_bps, st));
    ctx->launches += 1;
    CU(cudaMemcpyAsync(out_len.data(), d_out_len, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(status.data(), d_status, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // second pass for the streams whose slot was too small
    std::vector<size_t> redo;
    for (size_t i = 0; i < n; ++i)
        if (status[i] == S5B_ERR_NOSPACE) redo.push_back(i);
    std::vector<uint64_t> r_out_off;
    if (!redo.empty()) {
        const size_t m = redo.size();
        std::vector<uint64_t> r_in_off(m + 1);
        std::vector<uint32_t> r_in_len(m), r_out_len(m);
        std::vector<int32_t> r_status(m);
        r_out_off.resize(m + 1);
        uint64_t ro = 0;
        for (size_t k = 0; k < m; ++k) {
            // the input ranges are not contiguous any more: give every redo stream its own [off, off+len) and let
            // the entry after it bound nothing (capacity is checked against the whole slab)
            r_in_off[k] = in_off[redo[k]];
            r_in_len[k] = in_len[redo[k]];
            r_out_off[k] = ro;
            ro += round_up((uint64_t)out_len[redo[k]] + 16, 16);
        }
        r_in_off[m] = tot;
        r_out_off[m] = ro;
        CU(s.d_c.reserve(ro + 16));
        CU(cudaMemcpyAsync(d_in_off, r_in_off.data(), (m + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_out_off, r_out_off.data(), (m + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_in_len, r_in_len.data(), m * 4, cudaMemcpyHostToDevice, st));
        InflateArgs b{static_cast<const uint8_t *>(s.d_a.p), d_in_off, d_in_len, round_up(tot, 16), m,
                      static_cast<uint8_t *>(s.d_c.p), d_out_off, d_out_len, d_status, s.d_counter};
        CU(launch_inflate_ws(ctx->inf_work, b, ctx->num_sms, ctx->inf_bps, st));
        ctx->launches += 1;
        CU(cudaMemcpyAsync(r_out_len.data(), d_out_len, m * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(r_status.data(), d_status, m * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (size_t k = 0; k < m; ++k) {
            status[redo[k]] = r_status[k] == S5B_OK ? 1 /* marker: result lives in d_c */ : r_status[k];
            out_len[redo[k]] = r_out_len[k];
        }
    }
    // results back: pass-1 slots are sparse, copy each stream on its own range of one bulk D2H
    CU(ctx->h_stage_out.reserve(otot + 16));
    uint8_t *hout = static_cast<uint8_t *>(ctx->h_stage_out.p);
    CU(cudaMemcpyAsync(hout, s.d_b.p, otot, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int first = S5B_OK;
    std::vector<uint8_t> tmp;
    for (size_t i = 0; i < n; ++i) {
        if (status[i] != S5B_OK && status[i] != 1) {
            if (first == S5B_OK) first = status[i];
            continue;
        }
        void *mem = malloc(out_len[i] ? out_len[i] : 1);
        if (!mem) {
            if (first == S5B_OK) first = S5B_ERR_MEM;
            continue;
        }
        if (status[i] == S5B_OK) memcpy(mem, hout + out_off[i], out_len[i]);
        out_ptrs[i] = mem;
        out_n[i] = out_len[i];
    }
    for (size_t k = 0; k < redo.size(); ++k) {
        const size_t i = redo[k];
        if (status[i] == 1 && out_ptrs[i])
            CU(cudaMemcpy(out_ptrs[i], static_cast<uint8_t *>(s.d_c.p) + r_out_off[k], out_len[i], cudaMemcpyDeviceToHost));
    }
    return first;
}

// record / buffer entropy coding for the pointer-array forms: method is S5B_COMPRESS_ZLIB or S5B_COMPRESS_ZSTD
static int entropy_compress_ptrs(s5b_ctx_t *ctx, int method, const void *const *ptrs, const size_t *counts,
                                 const uint32_t *splits, size_t n, void **out_ptrs, size_t *out_n) {
    std::vector<uint64_t> in_off(n + 1), out_off(n + 1);
    std::vector<uint32_t> in_len(n), out_len(n);
    std::vector<int32_t> status(n);
    uint64_t tot = 0, otot = 0;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = nullptr;
        out_n[i] = 0;
        if (!ptrs[i] && counts[i]) return S5B_ERR_ARG;
        if (counts[i] > 0xfffffff0ull) return S5B_ERR_ARG;
        in_len[i] = (uint32_t)counts[i];
        in_off[i] = tot;
        tot += round_up(in_len[i], 16);
        out_off[i] = otot;
        otot += method == S5B_COMPRESS_ZSTD ? s5b_zstd_bound(in_len[i]) : s5b_zlib_bound(in_len[i]);
    }
    in_off[n] = tot;
    out_off[n] = otot;
    PipeSlot &s = ctx->slot[0];
    cudaStream_t st = s.stream;
    CU(ctx->h_stage_in.reserve(tot + 16));
    CU(ctx->h_stage_out.reserve(otot + 16));
    uint8_t *hin = static_cast<uint8_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i) memcpy(hin + in_off[i], ptrs[i], in_len[i]);
    CU(s.d_meta.reserve(2 * (n + 1) * 8 + 4 * n * 4 + 64));
    CU(s.d_a.reserve(tot + 16));
    CU(s.d_b.reserve(otot + 16));
    uint64_t *d_in_off = static_cast<uint64_t *>(s.d_meta.p);
    uint64_t *d_out_off = d_in_off + (n + 1);
    uint32_t *d_in_len = reinterpret_cast<uint32_t *>(d_out_off + (n + 1));
    uint32_t *d_out_len = d_in_len + n;
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    uint32_t *d_split = reinterpret_cast<uint32_t *>(d_status + n);
    if (splits) CU(cudaMemcpyAsync(d_split, splits, n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.d_a.p, hin, tot, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_off, in_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_out_off, out_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_len, in_len.data(), n * 4, cudaMemcpyHostToDevice, st));
    DeflateArgs a{static_cast<const uint8_t *>(s.d_a.p), d_in_off, d_in_len, round_up(tot, 16), splits ? d_split : nullptr, n,
                  static_cast<uint8_t *>(s.d_b.p), d_out_off, d_out_len, d_status, s.d_counter};
    if (method == S5B_COMPRESS_ZSTD) CU(launch_zstd_encode(a, ctx->num_sms, ctx->ze_bps, st));
    else CU(launch_deflate_ws(s.d_work, a, ctx->num_sms, ctx->def_bps, st));
    ctx->launches += 1;
    CU(cudaMemcpyAsync(out_len.data(), d_out_len, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(status.data(), d_status, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_stage_out.p, s.d_b.p, otot, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint8_t *hout = static_cast<const uint8_t *>(ctx->h_stage_out.p);
    int first = S5B_OK;
    for (size_t i = 0; i < n; ++i) {
        if (status[i] != S5B_OK) {
            if (first == S5B_OK) first = status[i];
            continue;
        }
        void *mem = malloc(out_len[i] ? out_len[i] : 1);
        if (!mem) {
            if (first == S5B_OK) first = S5B_ERR_MEM;
            continue;
        }
        memcpy(mem, hout + out_off[i], out_len[i]);
        out_ptrs[i] = mem;
        out_n[i] = out_len[i];
    }
    return first;
}

static int zstd_depress_ptrs(s5b_ctx_t *ctx, const void *const *ptrs, const size_t *counts, size_t n, void **out_ptrs,
                             size_t *out_n) {
    std::vector<uint64_t> in_off(n + 1), out_off(n + 1);
    std::vector<uint32_t> in_len(n), out_len(n);
    std::vector<int32_t> status(n);
    uint64_t tot = 0, otot = 0;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = nullptr;
        out_n[i] = 0;
        if (!ptrs[i] && counts[i]) return S5B_ERR_ARG;
        if (counts[i] > 0xffffffffull) return S5B_ERR_ARG;
        in_len[i] = (uint32_t)counts[i];
        in_off[i] = tot;
        tot += round_up(in_len[i], 16);
        uint64_t sz = 0;
        // frames without a content size are rejected by the kernel with the reference's verdict; give them no room
        if (s5b_zstd_content_size(ptrs[i], counts[i], &sz) != S5B_OK || sz > 0xfffffff0ull) sz = 0;
        out_off[i] = otot;
        otot += round_up(sz, 16);
    }
    in_off[n] = tot;
    out_off[n] = otot;
    PipeSlot &s = ctx->slot[0];
    cudaStream_t st = s.stream;
    CU(ctx->h_stage_in.reserve(tot + 16));
    CU(ctx->h_stage_out.reserve(otot + 16));
    uint8_t *hin = static_cast<uint8_t *>(ctx->h_stage_in.p);
    for (size_t i = 0; i < n; ++i) memcpy(hin + in_off[i], ptrs[i], in_len[i]);
    CU(s.d_meta.reserve(2 * (n + 1) * 8 + 3 * n * 4 + 64));
    CU(s.d_a.reserve(tot + 16));
    CU(s.d_b.reserve(otot + 16));
    uint64_t *d_in_off = static_cast<uint64_t *>(s.d_meta.p);
    uint64_t *d_out_off = d_in_off + (n + 1);
    uint32_t *d_in_len = reinterpret_cast<uint32_t *>(d_out_off + (n + 1));
    uint32_t *d_out_len = d_in_len + n;
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    CU(cudaMemcpyAsync(s.d_a.p, hin, tot, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_off, in_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_out_off, out_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_in_len, in_len.data(), n * 4, cudaMemcpyHostToDevice, st));
    InflateArgs a{static_cast<const uint8_t *>(s.d_a.p), d_in_off, d_in_len, round_up(tot, 16), n,
                  static_cast<uint8_t *>(s.d_b.p), d_out_off, d_out_len, d_status, s.d_counter};
    int rc = zstd_launch(ctx, a, st);
    if (rc != S5B_OK) return rc;
    CU(cudaMemcpyAsync(out_len.data(), d_out_len, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(status.data(), d_status, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_stage_out.p, s.d_b.p, otot, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint8_t *hout = static_cast<const uint8_t *>(ctx->h_stage_out.p);
    int first = S5B_OK;
    for (size_t i = 0; i < n; ++i) {
        if (status[i] != S5B_OK) {
            if (first == S5B_OK) first = status[i] == S5B_ERR_NOSPACE ? S5B_ERR_PRESS : status[i];
            continue;
        }
        void *mem = malloc(out_len[i] ? out_len[i] : 1);
        if (!mem) {
            if (first == S5B_OK) first = S5B_ERR_MEM;
            continue;
        }
        memcpy(mem, hout + out_off[i], out_len[i]);
        out_ptrs[i] = mem;
        out_n[i] = out_len[i];
    }
    return first;
}

static int copy_ptrs(const void *const *ptrs, const size_t *counts, size_t n, void **out_ptrs, size_t *out_n) {
    // SLOW5_COMPRESS_NONE: malloc + memcpy (slow5_press.c:340-350, :449-459)
    int first = S5B_OK;
    for (size_t i = 0; i < n; ++i) {
        out_ptrs[i] = malloc(counts[i] ? counts[i] : 1);
        if (!out_ptrs[i]) {
            out_n[i] = 0;
            if (first == S5B_OK) first = S5B_ERR_MEM;
            continue;
        }
        memcpy(out_ptrs[i], ptrs[i], counts[i]);
        out_n[i] = counts[i];
    }
    return first;
}

int s5b_compress_batch_host(s5b_ctx_t *ctx, int method, const void *const *ptrs, const size_t *counts, size_t n,
                            void **out_ptrs, size_t *out_n) {
    if (!ctx || (n && (!ptrs || !counts || !out_ptrs || !out_n))) return S5B_ERR_ARG;
    if (n == 0) return S5B_OK;
    switch (method) {
        case S5B_COMPRESS_NONE: return copy_ptrs(ptrs, counts, n, out_ptrs, out_n);
        case S5B_COMPRESS_SVB_ZD: return svbzd_compress_ptrs(ctx, ptrs, counts, n, out_ptrs, out_n);
        case S5B_COMPRESS_EX_ZD: return exzd_ptrs(ctx, true, ptrs, counts, n, out_ptrs, out_n);
        case S5B_COMPRESS_ZLIB:
        case S5B_COMPRESS_ZSTD: {
            DeviceGuard g(ctx->device);
            return entropy_compress_ptrs(ctx, method, ptrs, counts, nullptr, n, out_ptrs, out_n);
        }
        default: return S5B_ERR_ARG;
    }
}

int s5b_qts_round_batch_host(s5b_ctx_t *ctx, int bits, const void *const *ptrs, const size_t *counts, size_t n, void **out_ptrs,
                             size_t *out_n) {
    if (!ctx || bits < 0 || bits > 16 || (n && (!ptrs || !counts || !out_ptrs || !out_n))) return S5B_ERR_ARG;
    if (n == 0) return S5B_OK;
    if (bits == 0) return copy_ptrs(ptrs, counts, n, out_ptrs, out_n);  // slow5_press.c:1995-1996
    return qts_ptrs(ctx, bits, ptrs, counts, n, out_ptrs, out_n);
}

int s5b_qts_round_dev(s5b_ctx_t *ctx, int16_t *d_sig, uint64_t n_samples, int bits, void *stream) {
    if (!ctx || bits < 0 || bits > 16) return S5B_ERR_ARG;
    if (bits == 0 || n_samples == 0) return S5B_OK;
    if (!d_sig || (reinterpret_cast<uintptr_t>(d_sig) & 1u)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    CU(launch_qts_round(d_sig, n_samples, nullptr, bits, ctx->num_sms, st));
    ctx->launches += 1;
    return S5B_OK;
}

int s5b_compress_records_host(s5b_ctx_t *ctx, int method, const void *const *ptrs, const size_t *counts,
                              const uint32_t *splits, size_t n, void **out_ptrs, size_t *out_n) {
    if (!ctx || (n && (!ptrs || !counts || !out_ptrs || !out_n))) return S5B_ERR_ARG;
    if (method != S5B_COMPRESS_ZLIB && method != S5B_COMPRESS_ZSTD) return S5B_ERR_ARG;
    if (n == 0) return S5B_OK;
    DeviceGuard g(ctx->device);
    return entropy_compress_ptrs(ctx, method, ptrs, counts, splits, n, out_ptrs, out_n);
}

int s5b_depress_batch_host(s5b_ctx_t *ctx, int method, const void *const *ptrs, const size_t *counts, size_t n,
                           void **out_ptrs, size_t *out_n) {
    if (!ctx || (n && (!ptrs || !counts || !out_ptrs || !out_n))) return S5B_ERR_ARG;
    if (n == 0) return S5B_OK;
    switch (method) {
        case S5B_COMPRESS_NONE: return copy_ptrs(ptrs, counts, n, out_ptrs, out_n);
        case S5B_COMPRESS_SVB_ZD: return svbzd_depress_ptrs(ctx, ptrs, counts, n, out_ptrs, out_n);
        case S5B_COMPRESS_EX_ZD: return exzd_ptrs(ctx, false, ptrs, counts, n, out_ptrs, out_n);
        case S5B_COMPRESS_ZLIB: {
            DeviceGuard g(ctx->device);
            return zlib_depress_ptrs(ctx, ptrs, counts, n, out_ptrs, out_n);
        }
        case S5B_COMPRESS_ZSTD: {
            DeviceGuard g(ctx->device);
            return zstd_depress_ptrs(ctx, ptrs, counts, n, out_ptrs, out_n);
        }
        default: return S5B_ERR_ARG;
    }
}

// ---------------------------------------------------------------------------------------------
// whole-batch record transcoding, device resident between one H2D and one D2H
// ---------------------------------------------------------------------------------------------
void *s5b_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}
void s5b_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
int s5b::recode_chunk_sync(s5b_ctx *ctx, int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *h_in,
                           uint64_t in_bytes, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n,
                           uint8_t *h_out, uint64_t out_cap, uint64_t *out_bytes) {
    if (!ctx || !out_bytes) return S5B_ERR_ARG;
    *out_bytes = 0;
    if (n == 0) return S5B_OK;
    if (!h_in || !rec_off || !rec_len || !h_out) return S5B_ERR_ARG;
    auto rec_ok = [](int m) { return m == S5B_COMPRESS_NONE || m == S5B_COMPRESS_ZLIB || m == S5B_COMPRESS_ZSTD; };
    auto sig_ok = [](int m) { return m == S5B_COMPRESS_NONE || m == S5B_COMPRESS_SVB_ZD || m == S5B_COMPRESS_EX_ZD; };
    if (!rec_ok(in_rec) || !rec_ok(out_rec) || !sig_ok(in_sig) || !sig_ok(out_sig))
        return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    const int qts = ctx->qts_bits;                 // degrade: the samples change, so the signal is stored anew in any case
    const bool resig = in_sig != out_sig || qts > 0;
    cudaStream_t st = ctx->slot[0].stream;
    unsigned long long *counter = ctx->slot[0].d_counter;
    // ---- per-record arrays: 16 x u32[n] and 8 x u64[n+1]
    const size_t n1 = n + 1;
    CU(ctx->r_meta.reserve(16 * n * 4 + 8 * n1 * 8 + 256));
    CU(ctx->r_scratch.reserve(compact_scratch_bytes(n)));
    uint64_t *u64p = static_cast<uint64_t *>(ctx->r_meta.p);
    uint64_t *d_rec_off = u64p, *d_infl_off = u64p + n1, *d_sig_off = u64p + 2 * n1, *d_svb_off = u64p + 3 * n1,
             *d_packed_off = u64p + 4 * n1, *d_z_off = u64p + 5 * n1, *d_img_off = u64p + 6 * n1, *d_sigabs = u64p + 7 * n1;
    uint32_t *u32p = reinterpret_cast<uint32_t *>(u64p + 8 * n1);
    uint32_t *d_rec_len = u32p, *d_tmp = u32p + n, *d_infl_len = u32p + 2 * n, *d_svb_len = u32p + 3 * n,
             *d_packed_len = u32p + 4 * n, *d_z_len = u32p + 5 * n, *d_split = u32p + 6 * n, *d_ns2 = u32p + 7 * n;
    RecArrays ra{u32p + 8 * n, u32p + 9 * n, u32p + 10 * n, u32p + 11 * n, u32p + 12 * n,
                 reinterpret_cast<int32_t *>(u32p + 13 * n)};
    int32_t *d_st2 = reinterpret_cast<int32_t *>(u32p + 14 * n);
    int32_t *d_st3 = reinterpret_cast<int32_t *>(u32p + 15 * n);
    std::vector<int32_t> h_st(n);
    int first_err = S5B_OK;
    auto check_status = [&](const int32_t *d) -> int {
        if (cudaMemcpyAsync(h_st.data(), d, n * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return S5B_ERR_DEVICE;
        for (uint64_t i = 0; i < n; ++i)
            if (h_st[i] != S5B_OK && first_err == S5B_OK) first_err = h_st[i];
        return S5B_OK;
    };
    // exclusive scan of len[] (rounded to align) into off[0..n]; returns the total through *total
    auto scan_total = [&](const uint32_t *len, uint32_t align, uint64_t *off, uint64_t *total) -> cudaError_t {
        cudaError_t e = launch_scan(len, n, align, off, ctx->r_scratch.p, st);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(total, off + n, 8, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        ctx->launches += 3;
        return cudaStreamSynchronize(st);
    };

    // ---- input up
    const uint64_t in_cap = round_up(in_bytes, 16);
    CU(ctx->r_in.reserve(in_cap + 16));
    CU(cudaMemcpyAsync(ctx->r_in.p, h_in, in_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_rec_off, rec_off, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_rec_len, rec_len, n * 4, cudaMemcpyHostToDevice, st));
    const uint8_t *cur = static_cast<const uint8_t *>(ctx->r_in.p);
    const uint64_t *cur_off = d_rec_off;
    const uint32_t *cur_len = d_rec_len;
    uint64_t cur_cap = in_cap;

    // ---- record decompression (slow5.c:2586)
    if (in_rec == S5B_COMPRESS_ZLIB) {
        CU(launch_rec_plan(PLAN_INFLATE_GUESS, n, ra, d_rec_len, 4, d_tmp, st));
        for (int attempt = 0; attempt < 2; ++attempt) {
            uint64_t total = 0;
            CU(scan_total(d_tmp, 16, d_infl_off, &total));
            CU(ctx->r_infl.reserve(total + 16));
            InflateArgs ia{cur, cur_off, cur_len, cur_cap, n, static_cast<uint8_t *>(ctx->r_infl.p), d_infl_off, d_infl_len,
                           d_st2, counter};
            CU(launch_inflate_ws(ctx->inf_work, ia, ctx->num_sms, ctx->inf_bps, st));
            ctx->launches += 2;
            CU(cudaMemcpyAsync(h_st.data(), d_st2, n * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            bool overflow = false;
            for (uint64_t i = 0; i < n; ++i) overflow |= h_st[i] == S5B_ERR_NOSPACE;
            if (!overflow || attempt == 1) {
                for (uint64_t i = 0; i < n; ++i)
                    if (h_st[i] != S5B_OK && first_err == S5B_OK) first_err = h_st[i];
                cur_cap = round_up(total, 16);
                break;
            }
            // some slots were too small: the first pass reported the sizes needed -> exact slots, run again
            CU(cudaMemcpyAsync(d_tmp, d_infl_len, n * 4, cudaMemcpyDeviceToDevice, st));
        }
        if (first_err != S5B_OK) return first_err;
        cur = static_cast<const uint8_t *>(ctx->r_infl.p);
        cur_off = d_infl_off;
        cur_len = d_infl_len;
    }
    if (in_rec == S5B_COMPRESS_ZSTD) {
        // content sizes come from the frame headers (slow5_press.c:1206): read on the host, exact slots
        std::vector<uint32_t> sizes(n);
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t sz = 0;
            if (s5b_zstd_content_size(h_in + rec_off[i], rec_len[i], &sz) != S5B_OK || sz > 0xfffffff0ull) return S5B_ERR_PRESS;
            sizes[i] = (uint32_t)sz;
        }
        uint64_t total = 0;
        CU(cudaMemcpyAsync(d_tmp, sizes.data(), n * 4, cudaMemcpyHostToDevice, st));
        CU(scan_total(d_tmp, 16, d_infl_off, &total));
        CU(ctx->r_infl.reserve(total + 16));
        InflateArgs ia{cur, cur_off, cur_len, cur_cap, n, static_cast<uint8_t *>(ctx->r_infl.p), d_infl_off, d_infl_len, d_st2,
                       counter};
        {
            const int rc = zstd_launch(ctx, ia, st);
            if (rc != S5B_OK) return rc;
        }
        if (check_status(d_st2) != S5B_OK) return S5B_ERR_DEVICE;
        if (first_err != S5B_OK) return first_err == S5B_ERR_NOSPACE ? S5B_ERR_PRESS : first_err;
        cur = static_cast<const uint8_t *>(ctx->r_infl.p);
        cur_off = d_infl_off;
        cur_len = d_infl_len;
        cur_cap = round_up(total, 16);
    }
    // ---- where is the signal (slow5.c:2811-2927)
    {
        s5b::AuxLayout lay = ctx->aux_layout;
        lay.rg_n = ctx->rg_map_n;
        if (!qts) lay.ds_check = 0;
        CU(launch_rec_locate(cur, cur_off, cur_len, n, in_sig == S5B_COMPRESS_SVB_ZD ? 1 : (in_sig == S5B_COMPRESS_EX_ZD ? 2 : 0), ra,
                             st, nullptr, &lay));
        ctx->launches += 1;
        if (ctx->rg_map_n) {  // merge: read groups renumbered in the context's own copy of the records
            CU(launch_rec_rg_remap(const_cast<uint8_t *>(cur), cur_off, ra, n, ctx->d_rg_map, st));
            ctx->launches += 1;
        }
    }
    {
        int rc = check_status(ra.status);
        if (rc != S5B_OK) return cuda_fail(ctx, cudaGetLastError());
        if (first_err != S5B_OK) return first_err;
    }
    // ---- signal stage
    const uint8_t *sig_src = nullptr;  // nullptr = pass the stored bytes through
    const uint64_t *sig_src_off = nullptr;
    const uint32_t *sig_src_len = ra.sig_bytes;
    int sig_src_is_samples = 0;
    if (resig) {
        uint64_t total = 0;
        CU(launch_rec_plan(PLAN_SIG_SAMPLES, n, ra, nullptr, 0, d_tmp, st));
        CU(scan_total(d_tmp, 8, d_sig_off, &total));
        CU(ctx->r_sig.reserve(total * 2 + 32));
        ctx->launches += 1;
        // the raw samples of every read, in an aligned slab: decoded from the stored stream (slow5.c:2915) or copied out
        if (in_sig != S5B_COMPRESS_NONE) {
            CU(launch_rec_sig_abs(cur_off, ra, n, d_sigabs, st));
            SvbDecodeArgs da{cur, d_sigabs, ra.sig_bytes, cur_cap, n, static_cast<int16_t *>(ctx->r_sig.p), d_sig_off, d_ns2,
                             d_st2, counter};
            if (in_sig == S5B_COMPRESS_SVB_ZD) CU(launch_svbzd_decode(da, ctx->num_sms, ctx->dec_bps, st));
            else CU(launch_exzd_decode(da, ctx->num_sms, ctx->xd_bps, st));
            ctx->launches += 2;
            if (check_status(d_st2) != S5B_OK) return S5B_ERR_DEVICE;
            if (first_err != S5B_OK) return first_err;
        } else {
            CU(launch_sig_extract(cur, cur_off, ra, n, static_cast<int16_t *>(ctx->r_sig.p), d_sig_off, st));
            ctx->launches += 1;
        }
        if (qts) {  // src/degrade.c:255
            CU(launch_qts_round(static_cast<int16_t *>(ctx->r_sig.p), total, nullptr, qts, ctx->num_sms, st));
            ctx->launches += 1;
        }
        if (out_sig == S5B_COMPRESS_NONE) {
            sig_src = static_cast<const uint8_t *>(ctx->r_sig.p);
            sig_src_off = d_sig_off;
            sig_src_is_samples = 1;
        } else {  // encode (slow5.c:3973) into worst-case slots
            const bool svb = out_sig == S5B_COMPRESS_SVB_ZD;
            CU(launch_rec_plan(svb ? PLAN_SVB_BOUND : PLAN_EXZD_BOUND, n, ra, nullptr, 0, d_tmp, st));
            CU(scan_total(d_tmp, 16, d_svb_off, &total));
            CU(ctx->r_svb.reserve(total + 32));
            SvbEncodeArgs ea{static_cast<const int16_t *>(ctx->r_sig.p), d_sig_off, ra.n_samples, n,
                             static_cast<uint8_t *>(ctx->r_svb.p), d_svb_off, d_svb_len, d_st2, counter};
            if (svb) CU(launch_svbzd_encode(ea, ctx->num_sms, ctx->enc_bps, st));
            else CU(launch_exzd_encode(ea, ctx->num_sms, ctx->xe_bps, st));
            ctx->launches += 2;
            if (check_status(d_st2) != S5B_OK) return S5B_ERR_DEVICE;
            if (first_err != S5B_OK) return first_err;
            sig_src = static_cast<const uint8_t *>(ctx->r_svb.p);
            sig_src_off = d_svb_off;
            sig_src_len = d_svb_len;
        }
    }
    // ---- pack (slow5.c:3928-4044)
    const uint8_t *fin = cur;
    const uint64_t *fin_off = cur_off;
    const uint32_t *fin_len = cur_len;
    uint64_t fin_cap = cur_cap;
    if (resig) {
        uint64_t total = 0;
        if (sig_src_is_samples) {  // raw signal goes into the record: 2 * n_samples bytes
            CU(launch_rec_plan(PLAN_SIG_BYTES_RAW, n, ra, nullptr, 0, d_svb_len, st));
            sig_src_len = d_svb_len;
        }
        CU(launch_rec_plan(PLAN_PACKED_LEN, n, ra, sig_src_len, 0, d_packed_len, st));
        CU(scan_total(d_packed_len, 16, d_packed_off, &total));
        CU(ctx->r_packed.reserve(total + 32));
        CU(launch_rec_pack(cur, cur_off, ra, n, sig_src, sig_src_off, sig_src_len, sig_src_is_samples,
                           out_sig != S5B_COMPRESS_NONE, static_cast<uint8_t *>(ctx->r_packed.p), d_packed_off, st));
        ctx->launches += 3;
        fin = static_cast<const uint8_t *>(ctx->r_packed.p);
        fin_off = d_packed_off;
        fin_len = d_packed_len;
        fin_cap = round_up(total, 16);
    }
    // ---- record compression (slow5.c:4050)
    if (out_rec == S5B_COMPRESS_ZLIB || out_rec == S5B_COMPRESS_ZSTD) {
        if (in_rec == out_rec && !resig && !ctx->rg_map_n) {
            // nothing changed inside the records: the stored compressed records are the answer
            fin = static_cast<const uint8_t *>(ctx->r_in.p);
            fin_off = d_rec_off;
            fin_len = d_rec_len;
        } else {
            uint64_t total = 0;
            CU(launch_rec_plan(PLAN_ZLIB_BOUND, n, ra, fin_len, 0, d_tmp, st));
            CU(scan_total(d_tmp, 16, d_z_off, &total));
            CU(ctx->r_z.reserve(total + 32));
            const uint32_t *split = nullptr;
            if (out_sig == S5B_COMPRESS_SVB_ZD) {
                CU(launch_rec_plan(PLAN_SPLIT, n, ra, nullptr, 0, d_split, st));
                split = d_split;
            }
            DeflateArgs za{fin, fin_off, fin_len, fin_cap, split, n, static_cast<uint8_t *>(ctx->r_z.p), d_z_off, d_z_len, d_st3,
                           counter};
            // PLAN_ZLIB_BOUND slots also cover zstd_encode_bound() (3 bytes of header per block instead of 6)
            if (out_rec == S5B_COMPRESS_ZSTD) CU(launch_zstd_encode(za, ctx->num_sms, ctx->ze_bps, st));
            else CU(launch_deflate_ws(ctx->r_work, za, ctx->num_sms, ctx->def_bps, st));
            ctx->launches += 3;
            if (check_status(d_st3) != S5B_OK) return S5B_ERR_DEVICE;
            if (first_err != S5B_OK) return first_err;
            fin = static_cast<const uint8_t *>(ctx->r_z.p);
            fin_off = d_z_off;
            fin_len = d_z_len;
        }
    }
    // ---- file image: [u64 size][record] ... (slow5.c:4055-4060), one D2H
    {
        uint64_t total = 0;
        CU(launch_rec_plan(PLAN_IMAGE_LEN, n, ra, fin_len, 0, d_tmp, st));
        CU(scan_total(d_tmp, 1, d_img_off, &total));
        *out_bytes = total;
        if (total > out_cap) return S5B_ERR_NOSPACE;
        CU(ctx->r_img.reserve(total + 32));
        CU(launch_image_gather(fin, fin_off, fin_len, n, static_cast<uint8_t *>(ctx->r_img.p), d_img_off, st));
        ctx->launches += 2;
        CU(cudaMemcpyAsync(h_out, ctx->r_img.p, total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return first_err;
}

extern "C" {
// ---------------------------------------------------------------------------------------------
// read ids of a batch of stored records: the per-record work of slow5_idx_build (slow5_idx.c:283-334)
// ---------------------------------------------------------------------------------------------
int s5b_blow5_read_ids_host(s5b_ctx_t *ctx, int in_rec, const uint8_t *h_in, uint64_t in_bytes, const uint64_t *rec_off,
                            const uint32_t *rec_len, uint64_t n, uint8_t *h_ids, uint64_t ids_cap, uint64_t *id_off) {
    if (!id_off) return S5B_ERR_ARG;
    id_off[0] = 0;
    if (n == 0) return S5B_OK;
    if (!h_in || !rec_off || !rec_len || !h_ids) return S5B_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i)
        if (rec_off[i] + rec_len[i] > in_bytes) return S5B_ERR_ARG;
    if (in_rec == S5B_COMPRESS_NONE) {  // the id is right there
        uint64_t tot = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (rec_len[i] < 2) return S5B_ERR_PRESS;
            uint16_t rid;
            memcpy(&rid, h_in + rec_off[i], 2);
            if (2u + rid > rec_len[i]) return S5B_ERR_PRESS;
            if (tot + rid > ids_cap) return S5B_ERR_NOSPACE;
            memcpy(h_ids + tot, h_in + rec_off[i] + 2, rid);
            tot += rid;
            id_off[i + 1] = tot;
        }
        return S5B_OK;
    }
    if (!ctx || (in_rec != S5B_COMPRESS_ZLIB && in_rec != S5B_COMPRESS_ZSTD)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->slot[0].stream;
    unsigned long long *counter = ctx->slot[0].d_counter;
    // zlib: like the reference (slow5_idx.c:290-310) only the first 256 bytes of every record are decompressed -- a cut
    // stream is not an error for inflate, it yields the bytes decoded so far; zstd frames are decoded in full (:289)
    const bool zl = in_rec == S5B_COMPRESS_ZLIB;
    const uint32_t PFX = 256, SLOT = 4096;
    std::vector<uint32_t> idlen(n);
    uint64_t tot = 0;
    std::vector<uint64_t> redo;  // records whose prefix did not reach the end of the id
    const uint64_t SUB = 200000;
    for (uint64_t b0 = 0; b0 < n; b0 += SUB) {
        const uint64_t m = (n - b0 < SUB) ? n - b0 : SUB;
        // ---- pinned input slab + offsets
        std::vector<uint64_t> ioff(m + 1), ooff(m + 1);
        std::vector<uint32_t> ilen(m);
        uint64_t itot = 0, otot = 0;
        for (uint64_t i = 0; i < m; ++i) {
            const uint32_t l = zl ? (rec_len[b0 + i] < PFX ? rec_len[b0 + i] : PFX) : rec_len[b0 + i];
            ioff[i] = itot;
            ilen[i] = l;
            itot += round_up(l, 16);
            ooff[i] = otot;
            if (zl) {
                otot += SLOT;
            } else {
                uint64_t sz = 0;
                if (s5b_zstd_content_size(h_in + rec_off[b0 + i], rec_len[b0 + i], &sz) != S5B_OK || sz > 0xfffffff0ull)
                    return S5B_ERR_PRESS;
                otot += round_up(sz, 16);
            }
        }
        ioff[m] = itot;
        ooff[m] = otot;
        CU(ctx->h_stage_in.reserve(itot + 16));
        uint8_t *hin = static_cast<uint8_t *>(ctx->h_stage_in.p);
        for (uint64_t i = 0; i < m; ++i) memcpy(hin + ioff[i], h_in + rec_off[b0 + i], ilen[i]);
        // ---- device buffers: [ioff][ooff][src] u64 (m+1 each) | [ilen][olen][idlen] u32 | [status] i32
        const size_t meta = 4 * (m + 1) * 8 + 4 * m * 4;
        CU(ctx->r_meta.reserve(meta + 64));
        CU(ctx->r_in.reserve(itot + 16));
        CU(ctx->r_infl.reserve(otot + 16));
        CU(ctx->r_scratch.reserve(compact_scratch_bytes(m)));
        uint64_t *d_ioff = static_cast<uint64_t *>(ctx->r_meta.p), *d_ooff = d_ioff + (m + 1), *d_src = d_ooff + (m + 1),
                 *d_dense = d_src + (m + 1);
        uint32_t *d_ilen = reinterpret_cast<uint32_t *>(d_dense + (m + 1)), *d_olen = d_ilen + m, *d_idlen = d_olen + m;
        int32_t *d_st = reinterpret_cast<int32_t *>(d_idlen + m);
        CU(cudaMemcpyAsync(d_ioff, ioff.data(), (m + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_ooff, ooff.data(), (m + 1) * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_ilen, ilen.data(), m * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->r_in.p, hin, itot, cudaMemcpyHostToDevice, st));
        InflateArgs ia{static_cast<const uint8_t *>(ctx->r_in.p), d_ioff, d_ilen, round_up(itot, 16), m,
                       static_cast<uint8_t *>(ctx->r_infl.p), d_ooff, d_olen, d_st, counter};
        if (zl) {
            CU(launch_inflate_ws(ctx->inf_work, ia, ctx->num_sms, ctx->inf_bps, st));
            ctx->launches += 1;
        } else {
            const int rc = zstd_launch(ctx, ia, st);
            if (rc != S5B_OK) return rc;
        }
        CU(launch_rec_ids(static_cast<const uint8_t *>(ctx->r_infl.p), d_ooff, d_olen, d_st, m, d_idlen, d_src, st));
        CU(cudaMemcpyAsync(idlen.data() + b0, d_idlen, m * 4, cudaMemcpyDeviceToHost, st));
        std::vector<int32_t> h_st(m);
        CU(cudaMemcpyAsync(h_st.data(), d_st, m * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        // records to redo get length 0 in the gather; data errors are errors (a zstd frame that fails, a zlib prefix that is
        // not a zlib stream); NOSPACE on a prefix just means the slot was too small to hold what 256 bytes inflate to
        std::vector<uint32_t> glen(m);
        for (uint64_t i = 0; i < m; ++i) {
            if (h_st[i] == S5B_ERR_PRESS) return S5B_ERR_PRESS;
            if (idlen[b0 + i] == 0xFFFFFFFFu) {
                if (!zl) return S5B_ERR_PRESS;
                redo.push_back(b0 + i);
                glen[i] = 0;
            } else {
                glen[i] = idlen[b0 + i];
            }
        }
        CU(cudaMemcpyAsync(d_idlen, glen.data(), m * 4, cudaMemcpyHostToDevice, st));
        int nl = 0;
        // dense gather of the id bytes, then one small D2H
        uint64_t sub_tot = 0;
        for (uint64_t i = 0; i < m; ++i) sub_tot += glen[i];
        CU(ctx->r_packed.reserve(sub_tot + 64));
        CU(launch_compact(static_cast<const uint8_t *>(ctx->r_infl.p), d_src, d_idlen, m, 1,
                          static_cast<uint8_t *>(ctx->r_packed.p), d_dense, ctx->r_scratch.p, st, &nl));
        ctx->launches += 1 + nl;
        if (tot + sub_tot > ids_cap) return S5B_ERR_NOSPACE;
        CU(cudaMemcpyAsync(h_ids + tot, ctx->r_packed.p, sub_tot, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (uint64_t i = 0; i < m; ++i) {
            tot += glen[i];
            id_off[b0 + i + 1] = tot;
        }
    }
    if (!redo.empty()) {
        // ids longer than what the prefix gave: decompress those records in full (slow5_idx.c:312-320), then rebuild the slab
        std::vector<const void *> ptrs(redo.size());
        std::vector<size_t> counts(redo.size()), out_n(redo.size());
        std::vector<void *> outs(redo.size(), nullptr);
        for (size_t k = 0; k < redo.size(); ++k) {
            ptrs[k] = h_in + rec_off[redo[k]];
            counts[k] = rec_len[redo[k]];
        }
        const int rc = s5b_depress_batch_host(ctx, S5B_COMPRESS_ZLIB, ptrs.data(), counts.data(), redo.size(), outs.data(),
                                              out_n.data());
        std::vector<std::string> ids(redo.size());
        int bad = rc;
        for (size_t k = 0; k < redo.size(); ++k) {
            if (outs[k] && out_n[k] >= 2) {
                uint16_t rid;
                memcpy(&rid, outs[k], 2);
                if (2u + rid <= out_n[k]) ids[k].assign(static_cast<const char *>(outs[k]) + 2, rid);
                else bad = S5B_ERR_PRESS;
            } else if (bad == S5B_OK) {
                bad = S5B_ERR_PRESS;
            }
            free(outs[k]);
        }
        if (bad != S5B_OK) return bad;
        // splice: walk backwards so the dense slab can be expanded in place
        uint64_t extra = 0;
        for (const auto &sid : ids) extra += sid.size();
        if (tot + extra > ids_cap) return S5B_ERR_NOSPACE;
        std::vector<uint64_t> new_off(n + 1);
        uint64_t acc = 0;
        size_t k = 0;
        for (uint64_t i = 0; i < n; ++i) {
            new_off[i] = acc;
            if (k < redo.size() && redo[k] == i) acc += ids[k++].size();
            else acc += id_off[i + 1] - id_off[i];
        }
        new_off[n] = acc;
        k = redo.size();
        for (uint64_t i = n; i-- > 0;) {
            if (k > 0 && redo[k - 1] == i) {
                --k;
                memcpy(h_ids + new_off[i], ids[k].data(), ids[k].size());
            } else {
                memmove(h_ids + new_off[i], h_ids + id_off[i], id_off[i + 1] - id_off[i]);
            }
        }
        for (uint64_t i = 0; i <= n; ++i) id_off[i] = new_off[i];
    }
    return S5B_OK;
}

// ---------------------------------------------------------------------------------------------
// single buffers
// ---------------------------------------------------------------------------------------------
namespace {
struct TlCtx {
    s5b_ctx_t *ctx = nullptr;
    ~TlCtx() {
        if (ctx) s5b_ctx_destroy(ctx);
    }
};
s5b_ctx_t *thread_ctx(int *err) {
    static thread_local TlCtx t;
    if (!t.ctx) {
        int rc = s5b_ctx_create(-1, &t.ctx);
        if (rc != S5B_OK) {
            *err = rc;
            return nullptr;
        }
    }
    return t.ctx;
}
}  // namespace

void *s5b_ptr_compress_solo(int method, const void *ptr, size_t count, size_t *n) {
    void *out = nullptr;
    size_t out_n = 0;
    int rc = S5B_OK;
    if (!ptr || !n) {
        rc = S5B_ERR_ARG;  // slow5_press.c:334-338
    } else if (method == S5B_COMPRESS_NONE) {  // a copy (slow5_press.c:340-349): no device involved
        out = malloc(count ? count : 1);
        if (out) {
            memcpy(out, ptr, count);
            out_n = count;
        } else {
            rc = S5B_ERR_MEM;
        }
    } else if (s5b_ctx_t *ctx = thread_ctx(&rc)) {
        const void *ptrs[1] = {ptr};
        size_t counts[1] = {count};
        rc = s5b_compress_batch_host(ctx, method, ptrs, counts, 1, &out, &out_n);
    }
    tl_last_error = rc;
    if (rc != S5B_OK) {
        free(out);
        out = nullptr;
        out_n = 0;
    }
    if (n) *n = out_n;
    return out;
}

void *s5b_ptr_depress_solo(int method, const void *ptr, size_t count, size_t *n) {
    void *out = nullptr;
    size_t out_n = 0;
    int rc = S5B_OK;
    if (!ptr || !n) {
        rc = S5B_ERR_ARG;  // slow5_press.c:443-447
    } else if (method == S5B_COMPRESS_NONE) {  // a copy (slow5_press.c:449-458): no device involved
        out = malloc(count ? count : 1);
        if (out) {
            memcpy(out, ptr, count);
            out_n = count;
        } else {
            rc = S5B_ERR_MEM;
        }
    } else if (s5b_ctx_t *ctx = thread_ctx(&rc)) {
        const void *ptrs[1] = {ptr};
        size_t counts[1] = {count};
        rc = s5b_depress_batch_host(ctx, method, ptrs, counts, 1, &out, &out_n);
    }
    tl_last_error = rc;
    if (rc != S5B_OK) {
        free(out);
        out = nullptr;
        out_n = 0;
    }
    if (n) *n = out_n;
    return out;
}

int s5b_last_error(void) { return tl_last_error; }

}  // extern "C"
