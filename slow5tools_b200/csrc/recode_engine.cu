// recode_engine.cu -- the pipelined whole-batch BLOW5 record transcoder: the CUDA-stream batch scheduler that takes the
// place of the reference's fork-join pool around its per-record worker (src/view.c:35-57 depress_parse_rec_to_mem run by
// work_db, src/thread.c:114; driver loop slow5_convert_parallel src/view.c:254-301).
//
// A batch is cut into chunks (<= 24 Ki records / 256 MiB of stored bytes).  A chunk goes through
//     H2D -> record decompression (inflate | zstd) -> locate -> signal decode -> signal encode -> pack ->
//     record compression (deflate | zstd) -> file image -> D2H
// entirely on one stream with NO host synchronisation in between: every slab a stage needs is reserved up front from
// bounds that follow from the chunk's stored bytes alone (a compressed record is given 4x + 1 KiB to inflate into, a
// compressed signal holds at most one sample per byte, svb-zd needs <= 3.25 B/sample, ...), the per-record sizes a stage
// produces are scanned on the device and consumed there by the next stage, and the per-record statuses of all stages are
// folded into one (code, record) pair next to the chunk's image size.  Only that 24-byte verdict travels back before
// the payload D2H is queued.  Three lanes (stream + workspace each) keep the H2D of chunk i+1, the kernels of chunk i
// and the D2H of chunk i-1 in flight together; the host thread blocks once per chunk, on a chunk that finished its
// kernels two submissions ago.  A chunk the fast path cannot settle (a record that inflates to more than its slot, any
// malformed record) is redone by the careful transcoder (recode_chunk_sync, s5b_capi.cu), which sizes slots exactly
// and reports errors the way a serial loop over the records would.
//
// The device-resident form (s5b_blow5_recode_dev) is the same chunk pass with the payload already in HBM and the image
// written straight to its final place: no copies, no host synchronisation at all.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/slow5b200.h"
#include "s5b_ctx.h"

using namespace s5b;

namespace s5b {

void RecodeLane::release() {
    for (DevBuf *b : {&in, &infl, &sig, &svb, &packed, &z, &img, &meta, &scratch, &zd_scratch, &tab, &work, &iwork}) b->release();
    h_tab.release();
    if (d_counter) cudaFree(d_counter);
    if (d_res) cudaFree(d_res);
    if (h_res) cudaFreeHost(h_res);
    if (front) cudaEventDestroy(front);
    if (stream) cudaStreamDestroy(stream);
    d_counter = nullptr;
    d_res = nullptr;
    h_res = nullptr;
    front = nullptr;
    stream = nullptr;
}

void recode_lanes_release(s5b_ctx *ctx) {
    for (int i = 0; i < NLANE; ++i) ctx->lane[i].release();
    if (ctx->d_img_base) cudaFree(ctx->d_img_base);
    ctx->d_img_base = nullptr;
    for (int i = 0; i < 2; ++i) {
        ctx->dev_tab[i].release();
        if (ctx->dev_tab_done[i]) cudaEventDestroy(ctx->dev_tab_done[i]);
        ctx->dev_tab_done[i] = nullptr;
    }
    if (ctx->d_rg_map) cudaFree(ctx->d_rg_map);
    ctx->d_rg_map = nullptr;
    ctx->rg_map_n = 0;
    for (auto &m : ctx->timer.marks) {
        cudaEventDestroy(m.a);
        cudaEventDestroy(m.b);
    }
    for (cudaEvent_t e : ctx->timer.pool) cudaEventDestroy(e);
    ctx->timer.marks.clear();
    ctx->timer.pool.clear();
    ctx->lanes_ready = false;
}

}  // namespace s5b

namespace {

int lanes_init(s5b_ctx *ctx) {
    if (ctx->lanes_ready) return S5B_OK;
    for (int i = 0; i < NLANE; ++i) {
        RecodeLane &l = ctx->lane[i];
        CU(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&l.front, cudaEventDisableTiming));
        CU(cudaMalloc(&l.d_counter, 256));
        CU(cudaMalloc(&l.d_res, 64));
        CU(cudaHostAlloc(&l.h_res, 64, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer(reinterpret_cast<void **>(&l.h_res_dev), l.h_res, 0));
    }
    CU(cudaMalloc(&ctx->d_img_base, 64));
    for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&ctx->dev_tab_done[i], cudaEventDisableTiming));
    ctx->lanes_ready = true;
    return S5B_OK;
}

// ---- optional per-stage timing: a pair of events around every stage, read back by s5b_ctx_stage_report ----
struct StageScope {
    s5b_ctx *ctx;
    cudaStream_t st;
    int stage;
    cudaEvent_t a = nullptr, b = nullptr;
    StageScope(s5b_ctx *c, cudaStream_t s, int stage_) : ctx(c), st(s), stage(stage_) {
        if (!ctx->timer.enabled) return;
        a = take();
        b = take();
        if (ctx->timer.isolate) cudaStreamSynchronize(st);  // dev: stage boundaries with an idle GPU on both sides
        if (a && b) cudaEventRecord(a, st);
    }
    cudaEvent_t take() {
        StageTimer &t = ctx->timer;
        if (!t.pool.empty()) {
            cudaEvent_t e = t.pool.back();
            t.pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        return e;
    }
    ~StageScope() {
        if (!a || !b) return;
        cudaEventRecord(b, st);
        if (ctx->timer.isolate) cudaStreamSynchronize(st);
        ctx->timer.marks.push_back({a, b, stage});
    }
};

struct Job {
    int in_rec, in_sig, out_rec, out_sig;
    const uint8_t *src;   // host or device payload
    bool src_dev;
    uint64_t src_bytes;
    const uint64_t *rec_off;  // host tables
    const uint32_t *rec_len;
    uint64_t n;
    uint8_t *dst;         // host or device image
    bool dst_dev;
    uint64_t dst_cap;
    uint64_t *abs_off;    // optional, n+1 entries, same side as dst
    uint64_t *d_acc;      // device form: [0] total image bytes, [1] first error
    const uint64_t *d_tab_off;  // device form: the whole record table, uploaded once per call
    const uint32_t *d_tab_len;
    bool edits = false;   // the records themselves change (read groups renumbered, samples degraded): stored records cannot be passed through
    int qts = 0;          // > 0: slow5_rec_qts_round with this many bits between signal decode and encode (s5b_ctx_set_degrade)
    bool resig() const { return in_sig != out_sig || qts > 0; }  // the signal is decoded and stored anew
};

constexpr uint64_t DEV_CHUNK_RECORDS_MAX = 4u << 20;  // records per chunk of the device-resident form, at most

struct Chunk {
    uint64_t first = 0, count = 0;
    uint64_t span0 = 0, span1 = 0;  // stored bytes [span0, span1) of the source hold the chunk's records
};

// workspace bounds of one chunk from its stored size alone (see the header comment)
struct Bounds {
    uint64_t infl, sig_samples, svb, packed, z, img;
};
Bounds chunk_bounds(const Job &j, uint64_t span, uint64_t n) {
    Bounds b{};
    const uint64_t D = j.in_rec == S5B_COMPRESS_NONE ? span : 4 * span + 1040 * n;  // decompressed record bytes (slots)
    b.infl = j.in_rec == S5B_COMPRESS_NONE ? 0 : D + 64;
    const uint64_t samples = j.in_sig == S5B_COMPRESS_NONE ? D / 2 : D;
    uint64_t sig_out = D;  // signal bytes inside the output records
    if (j.resig()) {
        b.sig_samples = samples + 8 * n + 64;
        if (j.out_sig == S5B_COMPRESS_SVB_ZD) b.svb = 3 * samples + samples / 4 + 40 * n + 64;
        else if (j.out_sig == S5B_COMPRESS_EX_ZD) b.svb = 2 * samples + 1040 * n + 64;
        sig_out = j.out_sig == S5B_COMPRESS_NONE ? 2 * samples : b.svb;
        b.packed = D + sig_out + 32 * n + 64;
    }
    const uint64_t P = j.resig() ? b.packed : D;
    const bool untouched = j.in_rec == j.out_rec && !j.resig() && !j.edits;  // stored records are the answer
    if ((j.out_rec == S5B_COMPRESS_ZLIB || j.out_rec == S5B_COMPRESS_ZSTD) && !untouched)
        b.z = P + 6 * (P / 6144 + 2 * n) + 48 * n + 64;
    const uint64_t F = b.z ? b.z : (untouched ? span : P);
    b.img = F + 8 * n + 64;
    if (j.resig() && j.out_rec == S5B_COMPRESS_NONE) b.img = b.packed + 8 * n + 64;
    return b;
}

// Enqueues the whole pass over one chunk on the lane's stream.  Host form: the image lands in lane.img and the verdict
// in lane.h_res (read after lane.front).  Device form: the image lands in j.dst at *ctx->d_img_base, which is advanced.
int enqueue_chunk(s5b_ctx *ctx, RecodeLane &L, const Job &j, const Chunk &c) {
    cudaStream_t st = L.stream;
    const uint64_t n = c.count, n1 = n + 1;
    const uint64_t span = c.span1 - c.span0;
    const Bounds B = chunk_bounds(j, span, n);
    CU(L.meta.reserve(18 * n * 4 + 8 * n1 * 8 + 256));
    CU(L.scratch.reserve(compact_scratch_bytes(n)));
    uint64_t *u64p = static_cast<uint64_t *>(L.meta.p);
    uint64_t *d_rec_off = u64p, *d_infl_off = u64p + n1, *d_sig_off = u64p + 2 * n1, *d_svb_off = u64p + 3 * n1,
             *d_packed_off = u64p + 4 * n1, *d_z_off = u64p + 5 * n1, *d_img_off = u64p + 6 * n1, *d_sigabs = u64p + 7 * n1;
    uint32_t *u32p = reinterpret_cast<uint32_t *>(u64p + 8 * n1);
    const uint32_t *d_rec_len = u32p;
    uint32_t *d_tmp = u32p + n, *d_infl_len = u32p + 2 * n, *d_svb_len = u32p + 3 * n,
             *d_packed_len = u32p + 4 * n, *d_z_len = u32p + 5 * n, *d_split = u32p + 6 * n, *d_ns2 = u32p + 7 * n;
    RecArrays ra{u32p + 8 * n, u32p + 9 * n, u32p + 10 * n, u32p + 11 * n, u32p + 12 * n,
                 reinterpret_cast<int32_t *>(u32p + 13 * n)};
    // one status array per stage; the ones a pass does not run stay out of the verdict (null below)
    int32_t *d_st_dep = reinterpret_cast<int32_t *>(u32p + 14 * n), *d_st_sdec = reinterpret_cast<int32_t *>(u32p + 15 * n),
            *d_st_senc = reinterpret_cast<int32_t *>(u32p + 16 * n), *d_st_z = reinterpret_cast<int32_t *>(u32p + 17 * n);
    const int32_t *st_dep = nullptr, *st_sdec = nullptr, *st_senc = nullptr, *st_z = nullptr;
    auto scan = [&](const uint32_t *len, uint32_t align, uint64_t *off) -> cudaError_t {
        ctx->launches += 3;
        return launch_scan(len, n, align, off, L.scratch.p, st);
    };

    // ---- input: the chunk's stored bytes and its slice of the record table
    const uint8_t *cur;
    uint64_t cur_cap;
    {
        StageScope ts(ctx, st, ST_H2D);
        // the slab keeps the source's 16-byte phase so aligned records stay aligned
        const uint64_t skew = c.span0 & 15u;
        if (j.src_dev) {
            cur = j.src + (c.span0 - skew);
            cur_cap = round_up(skew + span, 16);
        } else {
            CU(L.in.reserve(round_up(skew + span, 16) + 32));
            CU(cudaMemcpyAsync(static_cast<uint8_t *>(L.in.p) + skew, j.src + c.span0, span, cudaMemcpyHostToDevice, st));
            cur = static_cast<const uint8_t *>(L.in.p);
            cur_cap = round_up(skew + span, 16);
        }
        if (j.d_tab_off) {
            d_rec_len = j.d_tab_len + c.first;
            CU(launch_rebase_off(d_rec_off, j.d_tab_off + c.first, n, c.span0 - skew, st));
        } else {
            // through pinned staging: a copy from pageable memory would stall the host until the stream drains
            CU(L.h_tab.reserve(n * 20));
            uint64_t *t_off = static_cast<uint64_t *>(L.h_tab.p);
            uint32_t *t_len = reinterpret_cast<uint32_t *>(t_off + n);
            memcpy(t_off, j.rec_off + c.first, n * 8);
            memcpy(t_len, j.rec_len + c.first, n * 4);
            CU(cudaMemcpyAsync(d_rec_off, t_off, n * 8, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(u32p, t_len, n * 4, cudaMemcpyHostToDevice, st));
            CU(launch_rebase_off(d_rec_off, d_rec_off, n, c.span0 - skew, st));
        }
        ctx->launches += 1;
    }
    const uint64_t *cur_off = d_rec_off;
    const uint32_t *cur_len = d_rec_len;

    // ---- record decompression (slow5.c:2586)
    if (j.in_rec == S5B_COMPRESS_ZLIB || j.in_rec == S5B_COMPRESS_ZSTD) {
        StageScope ts(ctx, st, ST_REC_DEPRESS);
        CU(L.infl.reserve(B.infl));
        if (j.in_rec == S5B_COMPRESS_ZLIB) {
            CU(launch_rec_plan(PLAN_INFLATE_GUESS, n, ra, d_rec_len, 4, d_tmp, st));
        } else {
            CU(launch_zstd_sizes(cur, cur_off, cur_len, n, 4, 1024, d_tmp, d_st_dep, st));
        }
        CU(scan(d_tmp, 16, d_infl_off));
        InflateArgs ia{cur, cur_off, cur_len, cur_cap, n, static_cast<uint8_t *>(L.infl.p), d_infl_off, d_infl_len, d_st_dep,
                       L.d_counter};
        if (j.in_rec == S5B_COMPRESS_ZLIB) {
            CU(launch_inflate_ws(L.iwork, ia, ctx->num_sms, ctx->inf_bps, st));
        } else {
            CU(L.zd_scratch.reserve(zstd_decode_scratch_bytes(ctx->num_sms, ctx->zd_bps)));
            CU(launch_zstd_decode(ia, ctx->num_sms, ctx->zd_bps, L.zd_scratch.p, st));
        }
        ctx->launches += 2;
        st_dep = d_st_dep;
        cur = static_cast<const uint8_t *>(L.infl.p);
        cur_off = d_infl_off;
        cur_len = d_infl_len;
        cur_cap = round_up(B.infl - 64, 16);
    }
    // ---- where is the signal (slow5.c:2811-2927)
    {
        StageScope ts(ctx, st, ST_GLUE);
        s5b::AuxLayout lay = ctx->aux_layout;
        lay.rg_n = ctx->rg_map_n;
        if (!j.qts) lay.ds_check = 0;
        CU(launch_rec_locate(cur, cur_off, cur_len, n,
                             j.in_sig == S5B_COMPRESS_SVB_ZD ? 1 : (j.in_sig == S5B_COMPRESS_EX_ZD ? 2 : 0), ra, st, st_dep, &lay));
        ctx->launches += 1;
        if (ctx->rg_map_n) {
            // in place: `cur` is this lane's own copy of the records (inflated, or the uploaded input)
            CU(launch_rec_rg_remap(const_cast<uint8_t *>(cur), cur_off, ra, n, ctx->d_rg_map, st));
            ctx->launches += 1;
        }
    }
    // ---- signal stage
    const uint8_t *sig_src = nullptr;  // nullptr = pass the stored bytes through
    const uint64_t *sig_src_off = nullptr;
    const uint32_t *sig_src_len = ra.sig_bytes;
    int sig_src_is_samples = 0;
    // raw samples that go straight into the svb-zd encoder are read where they lie inside the records (any byte alignment):
    // no copy to an aligned sample slab in between
    const bool encode_in_place = j.in_sig == S5B_COMPRESS_NONE && j.out_sig == S5B_COMPRESS_SVB_ZD && !j.qts;
    if (j.resig()) {
        if (!encode_in_place) {
            CU(L.sig.reserve(B.sig_samples * 2 + 32));
            StageScope ts(ctx, st, ST_GLUE);
            CU(launch_rec_plan(PLAN_SIG_SAMPLES, n, ra, nullptr, 0, d_tmp, st));
            CU(scan(d_tmp, 8, d_sig_off));
            ctx->launches += 1;
        }
        if (encode_in_place) {
            StageScope ts(ctx, st, ST_GLUE);
            CU(launch_rec_sig_abs(cur_off, ra, n, d_sigabs, st));
            ctx->launches += 1;
        } else if (j.in_sig != S5B_COMPRESS_NONE) {
            StageScope ts(ctx, st, ST_SIG_DEPRESS);
            CU(launch_rec_sig_abs(cur_off, ra, n, d_sigabs, st));
            SvbDecodeArgs da{cur, d_sigabs, ra.sig_bytes, cur_cap, n, static_cast<int16_t *>(L.sig.p), d_sig_off, d_ns2, d_st_sdec,
                             L.d_counter};
            if (j.in_sig == S5B_COMPRESS_SVB_ZD) CU(launch_svbzd_decode(da, ctx->num_sms, ctx->dec_bps, st));
            else CU(launch_exzd_decode(da, ctx->num_sms, ctx->xd_bps, st));
            ctx->launches += 2;
            st_sdec = d_st_sdec;
        } else {
            StageScope ts(ctx, st, ST_SIG_EXTRACT);
            CU(launch_sig_extract(cur, cur_off, ra, n, static_cast<int16_t *>(L.sig.p), d_sig_off, st));
            ctx->launches += 1;
        }
        if (j.qts) {
            // degrade (src/degrade.c:255): the whole sample slab in one streaming pass, its length read on the device
            StageScope ts(ctx, st, ST_SIG_DEGRADE);
            CU(launch_qts_round(static_cast<int16_t *>(L.sig.p), 0, d_sig_off + n, j.qts, ctx->num_sms, st));
            ctx->launches += 1;
        }
        if (j.out_sig == S5B_COMPRESS_NONE) {
            sig_src = static_cast<const uint8_t *>(L.sig.p);
            sig_src_off = d_sig_off;
            sig_src_is_samples = 1;
        } else {  // encode (slow5.c:3973) into worst-case slots
            const bool svb = j.out_sig == S5B_COMPRESS_SVB_ZD;
            CU(L.svb.reserve(B.svb + 32));
            {
                StageScope ts(ctx, st, ST_GLUE);
                CU(launch_rec_plan(svb ? PLAN_SVB_BOUND : PLAN_EXZD_BOUND, n, ra, nullptr, 0, d_tmp, st));
                CU(scan(d_tmp, 16, d_svb_off));
                ctx->launches += 1;
            }
            StageScope ts(ctx, st, ST_SIG_PRESS);
            // a decode failure leaves n_samples as located; the encoder then works on unwritten samples, harmlessly:
            // the chunk is flagged through st_sdec and redone or reported
            SvbEncodeArgs ea{static_cast<const int16_t *>(L.sig.p), d_sig_off, ra.n_samples, n, static_cast<uint8_t *>(L.svb.p),
                             d_svb_off, d_svb_len, d_st_senc, L.d_counter};
            if (encode_in_place) {
                // (a record the locate kernel refused has n_samples = 0 and is flagged through its own status)
                ea.sig = nullptr;
                ea.sig_off = nullptr;
                ea.src_bytes = cur;
                ea.src_byte_off = d_sigabs;
                ea.src_capacity = cur_cap;
            }
            if (svb) CU(launch_svbzd_encode(ea, ctx->num_sms, ctx->enc_bps, st));
            else CU(launch_exzd_encode(ea, ctx->num_sms, ctx->xe_bps, st));
            ctx->launches += 1;
            st_senc = d_st_senc;  // (ex-zd can refuse a read whose stream would outgrow the reference's buffer)
            sig_src = static_cast<const uint8_t *>(L.svb.p);
            sig_src_off = d_svb_off;
            sig_src_len = d_svb_len;
        }
    }
    // ---- pack (slow5.c:3928-4044)
    const uint8_t *fin = cur;
    const uint64_t *fin_off = cur_off;
    const uint32_t *fin_len = cur_len;
    uint64_t fin_cap = cur_cap;
    // when the packed records are the output (no record compression), they are written straight into the file image
    const bool direct_image = j.resig() && j.out_rec == S5B_COMPRESS_NONE;
    const uint32_t *direct_sig_len = nullptr;
    if (direct_image) {
        if (sig_src_is_samples) {
            StageScope ts(ctx, st, ST_PACK);
            CU(launch_rec_plan(PLAN_SIG_BYTES_RAW, n, ra, nullptr, 0, d_svb_len, st));
            sig_src_len = d_svb_len;
            ctx->launches += 1;
        }
        direct_sig_len = sig_src_len;
    } else if (j.resig()) {
        StageScope ts(ctx, st, ST_PACK);
        CU(L.packed.reserve(B.packed + 32));
        if (sig_src_is_samples) {  // raw signal goes into the record: 2 * n_samples bytes
            CU(launch_rec_plan(PLAN_SIG_BYTES_RAW, n, ra, nullptr, 0, d_svb_len, st));
            sig_src_len = d_svb_len;
            ctx->launches += 1;
        }
        CU(launch_rec_plan(PLAN_PACKED_LEN, n, ra, sig_src_len, 0, d_packed_len, st));
        CU(scan(d_packed_len, 16, d_packed_off));
        CU(launch_rec_pack(cur, cur_off, ra, n, sig_src, sig_src_off, sig_src_len, sig_src_is_samples,
                           j.out_sig != S5B_COMPRESS_NONE, static_cast<uint8_t *>(L.packed.p), d_packed_off, st));
        ctx->launches += 2;
        fin = static_cast<const uint8_t *>(L.packed.p);
        fin_off = d_packed_off;
        fin_len = d_packed_len;
        fin_cap = round_up(B.packed, 16);
    }
    // ---- record compression (slow5.c:4050)
    if (j.out_rec == S5B_COMPRESS_ZLIB || j.out_rec == S5B_COMPRESS_ZSTD) {
        if (j.in_rec == j.out_rec && !j.resig() && !j.edits) {
            // nothing changed inside the records: the stored compressed records are the answer
            fin = j.src_dev ? j.src + (c.span0 - (c.span0 & 15u)) : static_cast<const uint8_t *>(L.in.p);
            fin_off = d_rec_off;
            fin_len = d_rec_len;
        } else {
            StageScope ts(ctx, st, ST_REC_PRESS);
            CU(L.z.reserve(B.z + 32));
            CU(launch_rec_plan(PLAN_ZLIB_BOUND, n, ra, fin_len, 0, d_tmp, st));
            CU(scan(d_tmp, 16, d_z_off));
            const uint32_t *split = nullptr;
            if (j.out_sig == S5B_COMPRESS_SVB_ZD) {
                CU(launch_rec_plan(PLAN_SPLIT, n, ra, nullptr, 0, d_split, st));
                split = d_split;
                ctx->launches += 1;
            }
            DeflateArgs za{fin, fin_off, fin_len, fin_cap, split, n, static_cast<uint8_t *>(L.z.p), d_z_off, d_z_len, d_st_z,
                           L.d_counter};
            // PLAN_ZLIB_BOUND slots also cover zstd_encode_bound() (3 bytes of header per block instead of 6)
            if (j.out_rec == S5B_COMPRESS_ZSTD) CU(launch_zstd_encode(za, ctx->num_sms, ctx->ze_bps, st));
            else CU(launch_deflate_ws(L.work, za, ctx->num_sms, ctx->def_bps, st));
            ctx->launches += 9;
            st_z = d_st_z;
            fin = static_cast<const uint8_t *>(L.z.p);
            fin_off = d_z_off;
            fin_len = d_z_len;
        }
    }
    // ---- file image: [u64 size][record] ... (slow5.c:4055-4060) and the chunk's verdict
    {
        StageScope ts(ctx, st, direct_image ? ST_PACK : ST_IMAGE);
        if (direct_image) CU(launch_rec_plan(PLAN_PACKED_IMAGE_LEN, n, ra, direct_sig_len, 0, d_tmp, st));
        else CU(launch_rec_plan(PLAN_IMAGE_LEN, n, ra, fin_len, 0, d_tmp, st));
        CU(scan(d_tmp, 1, d_img_off));
        uint8_t *img;
        const uint64_t *base_ptr = nullptr;
        uint64_t cap = 0;
        uint64_t *abs = nullptr;
        if (j.dst_dev) {
            img = j.dst;
            base_ptr = ctx->d_img_base;
            cap = j.dst_cap;
            abs = j.abs_off ? j.abs_off + c.first : nullptr;
        } else {
            CU(L.img.reserve(B.img + 32));
            img = static_cast<uint8_t *>(L.img.p);
        }
        // statuses: locate (covers the record decompression it was handed), signal decode / encode, record compression
        CU(launch_recode_finish(n, d_img_off, base_ptr, cap, ra.status, st_sdec, st_senc, st_z, L.d_res, st));
        if (direct_image) {
            CU(launch_rec_pack(cur, cur_off, ra, n, sig_src, sig_src_off, direct_sig_len, sig_src_is_samples,
                               j.out_sig != S5B_COMPRESS_NONE, img, d_img_off, st, 1, base_ptr, L.d_res, abs));
        } else {
            CU(launch_image_gather(fin, fin_off, fin_len, n, img, d_img_off, st, base_ptr, L.d_res, abs));
        }
        ctx->launches += 4;
        if (j.dst_dev) {
            CU(launch_recode_advance(ctx->d_img_base, L.d_res, j.d_acc, st));
            ctx->launches += 1;
        } else {
            // the 24 bytes the host waits for go through mapped pinned memory, written by a kernel: as a D2H copy they queued
            // behind the other lanes' payload copies on the copy engine (4 ms each), and the host, which needs the image size
            // before it can queue this chunk's payload, learned it that much later
            CU(launch_recode_publish(L.d_res, L.h_res_dev, st));
            ctx->launches += 1;
            if (j.abs_off) {
                // image offsets relative to the chunk (pinned staging, behind the table that has been consumed by now);
                // finish() adds the chunk's position once it is known
                CU(cudaMemcpyAsync(static_cast<uint8_t *>(L.h_tab.p) + n * 12, d_img_off, n * 8, cudaMemcpyDeviceToHost, st));
            }
        }
    }
    CU(cudaEventRecord(L.front, st));
    return S5B_OK;
}

bool methods_ok(int in_rec, int in_sig, int out_rec, int out_sig) {
    auto rec_ok = [](int m) { return m == S5B_COMPRESS_NONE || m == S5B_COMPRESS_ZLIB || m == S5B_COMPRESS_ZSTD; };
    auto sig_ok = [](int m) { return m == S5B_COMPRESS_NONE || m == S5B_COMPRESS_SVB_ZD || m == S5B_COMPRESS_EX_ZD; };
    return rec_ok(in_rec) && rec_ok(out_rec) && sig_ok(in_sig) && sig_ok(out_sig);
}

// cuts [0, n) into chunks; false when the record table is not laid out in ascending order (then: one chunk = everything)
bool cut_chunks(const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n, uint64_t src_bytes, uint64_t max_records,
                uint64_t max_bytes, std::vector<Chunk> &out) {
    out.clear();
    bool ascending = true;
    for (uint64_t i = 0; i < n; ++i) {
        if (rec_off[i] + rec_len[i] > src_bytes) return false;
        if (i && rec_off[i] < rec_off[i - 1] + rec_len[i - 1]) ascending = false;
    }
    if (!ascending) {
        Chunk c;
        c.first = 0;
        c.count = n;
        c.span0 = 0;
        c.span1 = src_bytes;
        out.push_back(c);
        return true;
    }
    uint64_t first = 0;
    while (first < n) {
        uint64_t last = first;  // inclusive
        const uint64_t s0 = rec_off[first];
        while (last + 1 < n && last + 1 - first < max_records && rec_off[last + 1] + rec_len[last + 1] - s0 <= max_bytes) ++last;
        Chunk c;
        c.first = first;
        c.count = last - first + 1;
        c.span0 = s0;
        c.span1 = rec_off[last] + rec_len[last];
        out.push_back(c);
        first = last + 1;
    }
    return true;
}

}  // namespace

extern "C" {

int s5b_blow5_recode_batch_host(s5b_ctx_t *ctx, int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *h_in,
                                uint64_t in_bytes, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n,
                                uint8_t *h_out, uint64_t out_cap, uint64_t *out_bytes, uint64_t *out_img_off) {
    if (!ctx || !out_bytes) return S5B_ERR_ARG;
    *out_bytes = 0;
    if (out_img_off) out_img_off[0] = 0;
    if (n == 0) return S5B_OK;
    if (!h_in || !rec_off || !rec_len || !h_out) return S5B_ERR_ARG;
    if (!methods_ok(in_rec, in_sig, out_rec, out_sig)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    {
        const int rc = lanes_init(ctx);
        if (rc != S5B_OK) return rc;
    }
    const int NL = ctx->n_lanes;
    std::vector<Chunk> chunks;
    if (!cut_chunks(rec_off, rec_len, n, in_bytes, ctx->recode_chunk_records, ctx->recode_chunk_bytes, chunks)) return S5B_ERR_ARG;
    Job j{in_rec, in_sig, out_rec, out_sig, h_in, false, in_bytes, rec_off, rec_len, n, h_out, false, out_cap, out_img_off, nullptr,
          nullptr, nullptr};
    j.edits = ctx->rg_map_n != 0;
    j.qts = ctx->qts_bits;
    const size_t nc = chunks.size();
    uint64_t pos = 0;           // image bytes placed so far
    bool overflow = false;      // the image outgrew out_cap: keep sizing, stop copying
    int first_err = S5B_OK;
    std::vector<uint64_t> chunk_pos(nc, 0);
    // finish(k): the chunk's verdict is in; queue its payload D2H (or redo it carefully)
    auto finish = [&](size_t k) -> int {
        RecodeLane &L = ctx->lane[k % NL];
        const Chunk &c = chunks[k];
        CU(cudaEventSynchronize(L.front));
        const uint64_t total = L.h_res[0];
        const int32_t err = (int32_t)L.h_res[1];
        chunk_pos[k] = pos;
        if (err != S5B_OK) {
            // careful path for this chunk: exact slots, exact first error
            CU(cudaStreamSynchronize(L.stream));
            std::vector<uint64_t> off(c.count);
            for (uint64_t i = 0; i < c.count; ++i) off[i] = rec_off[c.first + i] - c.span0;
            uint64_t got = 0;
            const uint64_t room = overflow || pos > out_cap ? 0 : out_cap - pos;
            int rc = recode_chunk_sync(ctx, in_rec, in_sig, out_rec, out_sig, h_in + c.span0, c.span1 - c.span0, off.data(),
                                       rec_len + c.first, c.count, h_out + (room ? pos : 0), room, &got);
            if (rc == S5B_ERR_NOSPACE) {
                overflow = true;
                rc = S5B_OK;
            }
            if (rc != S5B_OK) {
                if (first_err == S5B_OK) first_err = rc;
                return rc;
            }
            if (out_img_off && !overflow) {
                // walk the size prefixes of the chunk's image (rare path)
                uint64_t at = 0;
                for (uint64_t i = 0; i < c.count; ++i) {
                    out_img_off[c.first + i] = at;
                    uint64_t sz;
                    memcpy(&sz, h_out + pos + at, 8);
                    at += 8 + sz;
                }
            }
            pos += got;
            return S5B_OK;
        }
        if (out_img_off) {
            const uint64_t *rel = reinterpret_cast<const uint64_t *>(static_cast<const uint8_t *>(L.h_tab.p) + c.count * 12);
            for (uint64_t i = 0; i < c.count; ++i) out_img_off[c.first + i] = rel[i];
        }
        if (pos + total > out_cap) overflow = true;
        if (!overflow) {
            StageScope ts(ctx, L.stream, ST_D2H);
            CU(cudaMemcpyAsync(h_out + pos, L.img.p, total, cudaMemcpyDeviceToHost, L.stream));
        }
        pos += total;
        return S5B_OK;
    };
    for (size_t k = 0; k < nc + (size_t)(NL - 1); ++k) {
        if (k < nc) {
            const int rc = enqueue_chunk(ctx, ctx->lane[k % NL], j, chunks[k]);
            if (rc != S5B_OK) {
                for (int i = 0; i < NLANE; ++i) cudaStreamSynchronize(ctx->lane[i].stream);
                return rc;
            }
        }
        if (k >= (size_t)(NL - 1)) {
            const int rc = finish(k - (size_t)(NL - 1));
            if (rc != S5B_OK) {
                for (int i = 0; i < NLANE; ++i) cudaStreamSynchronize(ctx->lane[i].stream);
                return rc;
            }
        }
    }
    for (int i = 0; i < NLANE; ++i) CU(cudaStreamSynchronize(ctx->lane[i].stream));
    *out_bytes = pos;
    if (overflow) return S5B_ERR_NOSPACE;
    if (out_img_off) {
        for (size_t k = 0; k < nc; ++k) {
            const Chunk &c = chunks[k];
            for (uint64_t i = 0; i < c.count; ++i) out_img_off[c.first + i] += chunk_pos[k];
        }
        out_img_off[n] = pos;
    }
    return first_err;
}

int s5b_blow5_recode_host(s5b_ctx_t *ctx, int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *h_in,
                          uint64_t in_bytes, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n, uint8_t *h_out,
                          uint64_t out_cap, uint64_t *out_bytes) {
    return s5b_blow5_recode_batch_host(ctx, in_rec, in_sig, out_rec, out_sig, h_in, in_bytes, rec_off, rec_len, n, h_out,
                                       out_cap, out_bytes, nullptr);
}

int s5b_blow5_recode_dev(s5b_ctx_t *ctx, int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *d_in,
                         uint64_t in_bytes, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n, uint8_t *d_out,
                         uint64_t out_cap, uint64_t *d_result, uint64_t *d_img_off) {
    if (!ctx || !d_result) return S5B_ERR_ARG;
    if (!methods_ok(in_rec, in_sig, out_rec, out_sig)) return S5B_ERR_ARG;
    if (n && (!d_in || !rec_off || !rec_len || !d_out)) return S5B_ERR_ARG;
    // bulk copies and realigning loads work on the 16-byte granules of the payload: its base must be a granule boundary
    if (n && (reinterpret_cast<uintptr_t>(d_in) & 15u)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    {
        const int rc = lanes_init(ctx);
        if (rc != S5B_OK) return rc;
    }
    RecodeLane &L = ctx->lane[0];
    CU(cudaMemsetAsync(d_result, 0, 16, L.stream));
    CU(cudaMemsetAsync(ctx->d_img_base, 0, 8, L.stream));
    if (n == 0) return S5B_OK;
    std::vector<Chunk> chunks;
    // Device-resident passes take the largest chunks the card has room for: there is no copy to overlap, and every chunk ends
    // with a tail -- the entropy kernels hand out records in rounds of 32 per warp that take the same few milliseconds whether
    // the GPU is full or not, so a 1 M-record pass cut into four chunks pays four partly empty last rounds (measured on the
    // north-star step, tools/dev/dev_chunk_sweep.sh: 262 144-record chunks 84.6 ms, one chunk 78.5 ms).  The workspace a
    // chunk needs follows from its stored bytes (chunk_bounds); 60 % of the memory that is free now, plus what this lane
    // already holds, is the budget.  S5B_RECODE_DEV_CHUNK / S5B_RECODE_DEV_CHUNK_MB fix the caps instead.
    uint64_t max_records = ctx->recode_dev_chunk_records, max_bytes = ctx->recode_dev_chunk_bytes;
    if (!max_records || !max_bytes) {
        Job probe{in_rec, in_sig, out_rec, out_sig, d_in, true, in_bytes, rec_off, rec_len, n, d_out, true, out_cap, d_img_off, d_result,
                  nullptr, nullptr};
        probe.qts = ctx->qts_bits;
        // the stored bytes a single chunk would span (records and whatever lies between them)
        uint64_t lo = ~0ull, hi = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (rec_off[i] < lo) lo = rec_off[i];
            if (rec_off[i] + rec_len[i] > hi) hi = rec_off[i] + rec_len[i];
        }
        const uint64_t stored = hi > lo ? hi - lo : 0;
        const Bounds b = chunk_bounds(probe, stored, n);
        // + the deflate workspace (0.45 B per packed byte) and the per-record tables
        const double need = (double)b.infl + 2.0 * b.sig_samples + b.svb + b.packed + 1.45 * b.z + 256.0 * n;
        uint64_t parts = 1;
        // (a lane that has done this pass before holds the slabs already: nothing to ask the driver)
        const bool holds_it = L.infl.cap >= b.infl && L.sig.cap >= 2 * b.sig_samples + 32 && L.svb.cap >= b.svb + 32 &&
                              L.packed.cap >= b.packed + 32 && L.z.cap >= b.z + 32;
        if (ctx->recode_dev_workspace) {
            if (need > (double)ctx->recode_dev_workspace) parts = (uint64_t)(need / (double)ctx->recode_dev_workspace) + 1;
        } else if (!holds_it) {
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            uint64_t held = 0;
            for (const DevBuf *d : {&L.infl, &L.sig, &L.svb, &L.packed, &L.z, &L.work, &L.iwork, &L.meta}) held += d->cap;
            const double budget = 0.6 * (double)free_b + (double)held;
            if (need > budget) parts = (uint64_t)(need / budget) + 1;
        }
        const uint64_t by_cap = (n + DEV_CHUNK_RECORDS_MAX - 1) / DEV_CHUNK_RECORDS_MAX;  // (per-record tables stay 32-bit friendly)
        if (parts < by_cap) parts = by_cap;
        if (!max_records) max_records = (n + parts - 1) / parts;
        if (!max_bytes) max_bytes = (uint64_t)((double)stored / parts * 1.02) + (64u << 20);
    }
    if (!cut_chunks(rec_off, rec_len, n, in_bytes, max_records, max_bytes, chunks)) return S5B_ERR_ARG;
    // the record table goes up once per call, chunks slice it on the device
    CU(L.tab.reserve(n * 12 + 64));
    uint64_t *d_tab_off = static_cast<uint64_t *>(L.tab.p);
    uint32_t *d_tab_len = reinterpret_cast<uint32_t *>(d_tab_off + n);
    {
        const unsigned t = ctx->dev_tab_next++ & 1u;
        CU(cudaEventSynchronize(ctx->dev_tab_done[t]));  // its upload of two calls ago (a fresh event counts as complete)
        CU(ctx->dev_tab[t].reserve(n * 12));
        memcpy(ctx->dev_tab[t].p, rec_off, n * 8);
        memcpy(static_cast<uint8_t *>(ctx->dev_tab[t].p) + n * 8, rec_len, n * 4);
        CU(cudaMemcpyAsync(d_tab_off, ctx->dev_tab[t].p, n * 12, cudaMemcpyHostToDevice, L.stream));
        CU(cudaEventRecord(ctx->dev_tab_done[t], L.stream));
    }
    if (ctx->rg_map_n) return S5B_ERR_ARG;  // renumbering works on the library's own copy of the records: host form only
    Job j{in_rec, in_sig, out_rec, out_sig, d_in, true, in_bytes, rec_off, rec_len, n, d_out, true, out_cap, d_img_off, d_result,
          d_tab_off, d_tab_len};
    j.qts = ctx->qts_bits;
    for (const Chunk &c : chunks) {
        const int rc = enqueue_chunk(ctx, L, j, c);
        if (rc != S5B_OK) return rc;
    }
    return S5B_OK;
}

int s5b_ctx_sync(s5b_ctx_t *ctx) {
    if (!ctx) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    if (ctx->lanes_ready)
        for (int i = 0; i < NLANE; ++i) CU(cudaStreamSynchronize(ctx->lane[i].stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return S5B_OK;
}

void *s5b_ctx_recode_stream(s5b_ctx_t *ctx) {
    if (!ctx) return nullptr;
    DeviceGuard g(ctx->device);
    if (lanes_init(ctx) != S5B_OK) return nullptr;
    return ctx->lane[0].stream;
}

int s5b_ctx_stage_timing(s5b_ctx_t *ctx, int enable) {
    if (!ctx) return S5B_ERR_ARG;
    ctx->timer.enabled = enable != 0;
    ctx->timer.isolate = getenv("S5B_STAGE_ISOLATE") != nullptr;
    return S5B_OK;
}

int s5b_ctx_stage_report(s5b_ctx_t *ctx, double *ms, uint64_t *count, int reset) {
    if (!ctx) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    StageTimer &t = ctx->timer;
    for (auto &m : t.marks) {
        float f = 0.f;
        if (cudaEventSynchronize(m.b) == cudaSuccess && cudaEventElapsedTime(&f, m.a, m.b) == cudaSuccess) {
            t.ms[m.stage] += f;
            t.count[m.stage] += 1;
        }
        t.pool.push_back(m.a);
        t.pool.push_back(m.b);
    }
    (void)cudaGetLastError();
    t.marks.clear();
    for (int i = 0; i < ST_COUNT; ++i) {
        if (ms) ms[i] = t.ms[i];
        if (count) count[i] = t.count[i];
        if (reset) {
            t.ms[i] = 0;
            t.count[i] = 0;
        }
    }
    return S5B_OK;
}

int s5b_ctx_set_rg_map(s5b_ctx_t *ctx, const uint32_t *map, uint32_t n) {
    if (!ctx || (n && !map)) return S5B_ERR_ARG;
    DeviceGuard g(ctx->device);
    if (ctx->d_rg_map) {
        // transcodes that use the old table have been waited for by their callers (the host form returns when its image is complete)
        cudaFree(ctx->d_rg_map);
        ctx->d_rg_map = nullptr;
    }
    ctx->rg_map_n = 0;
    if (n) {
        CU(cudaMalloc(&ctx->d_rg_map, (size_t)n * 4));
        CU(cudaMemcpy(ctx->d_rg_map, map, (size_t)n * 4, cudaMemcpyHostToDevice));
        ctx->rg_map_n = n;
    }
    return S5B_OK;
}

int s5b_ctx_set_recode_workspace(s5b_ctx_t *ctx, uint64_t max_bytes) {
    if (!ctx) return S5B_ERR_ARG;
    ctx->recode_dev_workspace = max_bytes;
    return S5B_OK;
}

int s5b_ctx_set_degrade(s5b_ctx_t *ctx, int bits, int check_dataset, float digitisation, float sampling_rate) {
    if (!ctx || bits < 0 || bits > 16) return S5B_ERR_ARG;
    ctx->qts_bits = bits;
    ctx->aux_layout.ds_check = bits && check_dataset ? 1u : 0u;
    ctx->aux_layout.ds_digitisation = digitisation;
    ctx->aux_layout.ds_sampling_rate = sampling_rate;
    return S5B_OK;
}

int s5b_ctx_set_aux_layout(s5b_ctx_t *ctx, const uint8_t *elem_size, const uint8_t *is_array, uint32_t n_fields) {
    if (!ctx) return S5B_ERR_ARG;
    s5b::AuxLayout lay;
    lay.ds_check = ctx->aux_layout.ds_check;  // the degrade rule is set on its own (s5b_ctx_set_degrade)
    lay.ds_digitisation = ctx->aux_layout.ds_digitisation;
    lay.ds_sampling_rate = ctx->aux_layout.ds_sampling_rate;
    if (n_fields != s5b::AUX_LAYOUT_UNKNOWN) {
        if (n_fields > (uint32_t)s5b::AUX_LAYOUT_MAX || (n_fields && (!elem_size || !is_array))) return S5B_ERR_ARG;
        lay.n = n_fields;
        for (uint32_t f = 0; f < n_fields; ++f) {
            if (elem_size[f] == 0 || elem_size[f] > 8) return S5B_ERR_ARG;
            lay.size[f] = elem_size[f];
            if (is_array[f]) lay.array_mask |= 1ull << f;
        }
    }
    ctx->aux_layout = lay;
    return S5B_OK;
}

int s5b_stage_count(void) { return ST_COUNT; }
const char *s5b_stage_name(int stage) {
    static const char *names[ST_COUNT] = {"h2d", "record_depress", "glue", "signal_depress", "signal_press", "pack",
                                          "record_press", "image", "d2h", "signal_extract", "signal_degrade"};
    return stage >= 0 && stage < ST_COUNT ? names[stage] : "";
}

}  // extern "C"
