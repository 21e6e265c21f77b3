// s5b_ctx.h -- internal: the codec context shared by the C-ABI translation units (s5b_capi.cu, recode_engine.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/slow5b200.h"
#include "s5b_kernels.h"

namespace s5b {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t slack = bytes / 4;
        if (slack > (64u << 20)) slack = 64u << 20;
        size_t want = bytes + slack + 256;
        want = (want + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr int NSLOT = 2;

struct PipeSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf d_a, d_b, d_c;       // payload in / slotted out / dense out
    DevBuf d_meta;              // offsets, lengths, statuses
    DevBuf d_scratch;           // scan scratch
    DevBuf d_work;              // deflate workspace
    PinBuf h_meta;              // pinned mirror of d_meta (both directions)
    unsigned long long *d_counter = nullptr;
};

// ---- the record transcoder's pipeline lanes (recode_engine.cu) --------------------------------
constexpr int NLANE = 6;   // lanes that exist; s5b_ctx::n_lanes of them are used (3 unless S5B_RECODE_LANES says otherwise)
// stages of a transcoding pass, for the optional per-stage CUDA-event timing (s5b_ctx_stage_timing)
enum Stage { ST_H2D = 0, ST_REC_DEPRESS, ST_GLUE, ST_SIG_DEPRESS, ST_SIG_PRESS, ST_PACK, ST_REC_PRESS, ST_IMAGE, ST_D2H,
             ST_SIG_EXTRACT, ST_SIG_DEGRADE, ST_COUNT };

struct RecodeLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t front = nullptr;     // everything up to the (size, status) read-back of the chunk in flight
    DevBuf in, infl, sig, svb, packed, z, img, meta, scratch, zd_scratch, tab, work, iwork;
    PinBuf h_tab;                    // pinned staging of the chunk's record table (up) and image offsets (down)
    unsigned long long *d_counter = nullptr;
    uint64_t *d_res = nullptr;       // [0] image bytes of the chunk, [1] first error (int32 in the low half), [2] its record
    uint64_t *h_res = nullptr;       // pinned mirror (mapped: the device writes it through h_res_dev)
    uint64_t *h_res_dev = nullptr;
    void release();
};

struct StageTimer {
    bool enabled = false;
    bool isolate = false;            // S5B_STAGE_ISOLATE: synchronise the stream around every stage (development aid)
    struct Mark { cudaEvent_t a, b; int stage; };
    std::vector<Mark> marks;         // recorded, not yet read
    std::vector<cudaEvent_t> pool;   // free events
    double ms[ST_COUNT] = {0};
    uint64_t count[ST_COUNT] = {0};
};

}  // namespace s5b

struct s5b_ctx {
    int device = 0;
    int num_sms = 0;
    int enc_bps = 0, dec_bps = 0, inf_bps = 0, def_bps = 0, zd_bps = 0, ze_bps = 0, xe_bps = 0, xd_bps = 0;
    s5b::DevBuf zd_scratch;
    cudaStream_t stream = nullptr;  // default stream for *_dev calls
    unsigned long long *d_counter = nullptr;
    s5b::DevBuf d_scratch;
    s5b::PipeSlot slot[s5b::NSLOT];
    s5b::PinBuf h_stage_in, h_stage_out;  // pointer-array forms
    s5b::DevBuf r_in, r_infl, r_sig, r_svb, r_packed, r_z, r_img, r_meta, r_scratch, r_work;  // the careful (synchronous) transcoder
    s5b::DevBuf def_work;                 // deflate workspace of the *_dev entry points
    s5b::DevBuf inf_work;                 // inflate scratch rows of everything that is not a pipeline lane
    s5b::RecodeLane lane[s5b::NLANE];     // the pipelined transcoder (lazily created)
    bool lanes_ready = false;
    int n_lanes = 3;
    uint64_t *d_img_base = nullptr;       // running output offset of a device-resident transcoding pass
    // the device-resident form uploads the caller's record table through pinned staging (a copy from pageable memory waits for
    // the stream to drain and leaves the GPU idle meanwhile); two buffers take turns, each guarded by the event of its last upload
    s5b::PinBuf dev_tab[2];
    cudaEvent_t dev_tab_done[2] = {nullptr, nullptr};
    unsigned dev_tab_next = 0;
    s5b::StageTimer timer;
    uint32_t *d_rg_map = nullptr;     // read_group renumbering of the file being transcoded (s5b_ctx_set_rg_map), rg_map_n entries
    uint32_t rg_map_n = 0;
    int qts_bits = 0;                 // > 0: the transcoder degrades every signal by this many bits (s5b_ctx_set_degrade)
    s5b::AuxLayout aux_layout;        // auxiliary columns of the file being transcoded (s5b_ctx_set_aux_layout), unknown by default
    uint64_t launches = 0;
    size_t chunk_bytes = 32u << 20;  // e2e is flat between 16 and 128 MiB (PCIe bound), 32 MiB marginally best
    // records / stored bytes per pipeline chunk of the host-form transcoder (S5B_RECODE_CHUNK, S5B_RECODE_CHUNK_MB).  Measured
    // on the north-star step (tools/gpu_e2e_sweep2.sh, 1 M records end to end): 8 Ki 2.67e6, 16 Ki 3.28e6, 24 Ki 3.58e6,
    // 32 Ki 3.51e6, 64 Ki 3.38e6, 128 Ki 3.06e6 reads/s -- small chunks shorten the pipeline's fill and drain, too small ones
    // starve the kernels (a chunk must still fill the GPU with 32-record rounds).
    size_t recode_chunk_records = 24576;
    size_t recode_chunk_bytes = 256u << 20;
    // the device-resident form has no copies to overlap: large chunks amortise launches and kernel tails.  0 = as large as the
    // free device memory allows (recode_engine.cu, s5b_blow5_recode_dev)
    size_t recode_dev_chunk_records = 0;
    size_t recode_dev_chunk_bytes = 0;
    uint64_t recode_dev_workspace = 0;    // s5b_ctx_set_recode_workspace: the budget in bytes, 0 = from the free memory
    std::string last_cuda_error;
};

namespace s5b {

inline int cuda_fail(s5b_ctx *c, cudaError_t e) {
    if (c) c->last_cuda_error = cudaGetErrorString(e);
    (void)cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? S5B_ERR_MEM : S5B_ERR_DEVICE;
}
#define CU(call)                                  \
    do {                                          \
        cudaError_t e__ = (call);                 \
        if (e__ != cudaSuccess) return s5b::cuda_fail(ctx, e__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// inflate's thread-per-stream kernel keeps its per-lane symbol lists in a scratch buffer: reserve it in `buf` and launch
inline cudaError_t launch_inflate_ws(DevBuf &buf, InflateArgs a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    cudaError_t e = buf.reserve(inflate_work_bytes(num_sms));
    if (e != cudaSuccess) return e;
    a.work = buf.p;
    a.work_bytes = buf.cap;
    return launch_inflate(a, num_sms, blocks_per_sm, st);
}

// deflate needs a workspace (sorted symbol lists, code lengths per block): reserve it in `buf` and launch
inline cudaError_t launch_deflate_ws(DevBuf &buf, DeflateArgs a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    const size_t need = deflate_work_bytes(a.in_capacity, a.n_reads);
    cudaError_t e = buf.reserve(need);
    if (e != cudaSuccess) return e;
    a.work = buf.p;
    a.work_bytes = buf.cap;
    return launch_deflate(a, num_sms, blocks_per_sm, st);
}

// the careful transcoder of one chunk (host syncs between stages, inflate-slot retry, exact error reporting); lives in
// s5b_capi.cu and is the fallback of the pipelined engine
int recode_chunk_sync(s5b_ctx *ctx, int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *h_in, uint64_t in_bytes,
                      const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n, uint8_t *h_out, uint64_t out_cap,
                      uint64_t *out_bytes);
int zstd_launch(s5b_ctx *ctx, const InflateArgs &a, cudaStream_t st);
void recode_lanes_release(s5b_ctx *ctx);

}  // namespace s5b
