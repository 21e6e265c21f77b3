// s5b_kernels.h -- internal launcher interface between the C-ABI (s5b_capi.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace s5b {

struct SvbEncodeArgs {
    const int16_t *sig;
    const uint64_t *sig_off;    // n_reads + 1, samples, multiples of 8
    const uint32_t *n_samples;  // n_reads
    uint64_t n_reads;
    uint8_t *svb;
    const uint64_t *svb_off;  // n_reads + 1, bytes
    uint32_t *svb_len;        // n_reads
    int32_t *status;          // n_reads
    unsigned long long *work_counter;  // zeroed before launch
    // Alternative source (sig == nullptr): the samples of read r are the 2 * n_samples[r] bytes at src_bytes + src_byte_off[r],
    // any byte alignment -- the raw signal where it lies inside a packed record (record path: no copy to an aligned slab first).
    const uint8_t *src_bytes = nullptr;
    const uint64_t *src_byte_off = nullptr;  // n_reads
    uint64_t src_capacity = 0;               // bytes of src_bytes that may be read (multiple of 16)
};

struct SvbDecodeArgs {
    const uint8_t *svb;
    const uint64_t *svb_off;  // n_reads + 1
    const uint32_t *svb_len;  // n_reads
    uint64_t svb_capacity;    // bytes, multiple of 16
    uint64_t n_reads;
    int16_t *sig;
    const uint64_t *sig_off;  // n_reads + 1
    uint32_t *n_samples;      // out
    int32_t *status;          // out
    unsigned long long *work_counter;
};

struct InflateArgs {
    const uint8_t *in;
    const uint64_t *in_off;  // n_reads + 1
    const uint32_t *in_len;  // n_reads
    uint64_t in_capacity;    // bytes, multiple of 16
    uint64_t n_reads;
    uint8_t *out;
    const uint64_t *out_off;  // n_reads + 1 (slot bounds)
    uint32_t *out_len;        // bytes produced (bytes NEEDED when status is S5B_ERR_NOSPACE)
    int32_t *status;
    unsigned long long *work_counter;
    void *work = nullptr;     // inflate only: device scratch of >= inflate_work_bytes(num_sms) bytes (thread-per-stream rows)
    uint64_t work_bytes = 0;
};

struct DeflateArgs {
    const uint8_t *in;
    const uint64_t *in_off;  // n_reads + 1
    const uint32_t *in_len;  // n_reads
    uint64_t in_capacity;    // bytes, multiple of 16
    const uint32_t *split;   // optional: byte offset inside record r where a new Huffman block should start
    uint64_t n_reads;
    uint8_t *out;
    const uint64_t *out_off;  // n_reads + 1 (slot bounds, >= deflate_bound(len))
    uint32_t *out_len;
    int32_t *status;
    unsigned long long *work_counter;
    void *work = nullptr;        // deflate only: device workspace of >= deflate_work_bytes(in_capacity, n_reads) bytes
    uint64_t work_bytes = 0;
};
int deflate_blocks_per_sm();
size_t deflate_work_bytes(uint64_t in_capacity, uint64_t n_reads);
uint64_t deflate_bound(uint64_t len);
cudaError_t launch_deflate(const DeflateArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st);

// zstd frame encoder: same argument block as deflate (split hint included)
int zstd_encode_blocks_per_sm();
uint64_t zstd_encode_bound(uint64_t len);
cudaError_t launch_zstd_encode(const DeflateArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st);

// zstd frames: same argument block as inflate (exact output slots: every frame carries its content size)
int zstd_decode_blocks_per_sm();
size_t zstd_decode_scratch_bytes(int num_sms, int blocks_per_sm);
cudaError_t launch_zstd_decode(const InflateArgs &a, int num_sms, int blocks_per_sm, void *lit_scratch, cudaStream_t st);

// grid sizing helpers (queried once per context)
int inflate_blocks_per_sm();
cudaError_t launch_inflate(const InflateArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st);
// thread-per-stream decoder for streams of at most max_len bytes (longer ones are left untouched)
size_t inflate_work_bytes(int num_sms);
cudaError_t launch_inflate_threads(const InflateArgs &a, uint32_t max_len, int num_sms, cudaStream_t st);
int svbzd_encode_blocks_per_sm();
int svbzd_decode_blocks_per_sm();

cudaError_t launch_svbzd_encode(const SvbEncodeArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st);
cudaError_t launch_svbzd_decode(const SvbDecodeArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st);
cudaError_t launch_svbzd_peek(const uint8_t *svb, const uint64_t *svb_off, const uint32_t *svb_len, uint64_t n_reads,
                              uint32_t *n_samples, cudaStream_t st);

// ex-zd signal codec (exzd_kernels.cu): same argument blocks as svb-zd (svb = the ex-zd byte slab)
int exzd_encode_blocks_per_sm();
int exzd_decode_blocks_per_sm();
uint64_t exzd_bound(uint32_t n_samples);
cudaError_t launch_exzd_encode(const SvbEncodeArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st);
cudaError_t launch_exzd_decode(const SvbDecodeArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st);

// exclusive scan of len[] rounded up to `align` units -> off[0..n] (n+1 entries); 3 launches
cudaError_t launch_scan(const uint32_t *len, uint64_t n, uint32_t align, uint64_t *off, void *scratch, cudaStream_t st);

// ---- record-level kernels (record_kernels.cu): locate the signal inside packed BLOW5 records, plan the
// layouts of the following stages, pack records, build the file image
struct RecArrays {           // per-record device arrays, n entries each (SoA)
    uint32_t *head_len;      // bytes before the u64 len_raw_signal field (slow5.c:3928-3987)
    uint32_t *n_samples;
    uint32_t *sig_at;        // offset of the stored signal bytes inside the record
    uint32_t *sig_bytes;     // stored signal bytes (input form)
    uint32_t *aux_len;
    int32_t *status;
};
// The auxiliary columns of the file (slow5_aux_meta_t): element size of every field and which ones are arrays (u64 count first).
// n == AUX_LAYOUT_UNKNOWN: the caller did not say, the auxiliary section is taken as it is.
constexpr uint32_t AUX_LAYOUT_UNKNOWN = 0xffffffffu;
constexpr int AUX_LAYOUT_MAX = 64;
struct AuxLayout {
    uint32_t rg_n = 0;   // > 0: read groups the file's header names (a record's read_group must be below it); 0: not checked
    uint32_t n = AUX_LAYOUT_UNKNOWN;
    uint64_t array_mask = 0;
    uint8_t size[AUX_LAYOUT_MAX] = {0};
    // degrade with an automatically chosen bit count (src/degrade.c:195-211, :249-253): every record must carry the dataset's
    // digitisation and sampling rate (compared as the reference does: the record's double against a float), else S5B_ERR_DATASET
    uint32_t ds_check = 0;
    float ds_digitisation = 0.f, ds_sampling_rate = 0.f;
};
// in_status (optional): the status the record decompression left; a failed record is marked and skipped
cudaError_t launch_rec_locate(const uint8_t *rec, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n,
                              int sig_is_svb, RecArrays a, cudaStream_t st, const int32_t *in_status = nullptr,
                              const AuxLayout *aux = nullptr);
// elementwise size planning: mode selects which length is written to out[]
enum RecPlan { PLAN_SIG_SAMPLES = 0, PLAN_SVB_BOUND = 1, PLAN_PACKED_LEN = 2, PLAN_ZLIB_BOUND = 3, PLAN_IMAGE_LEN = 4,
               PLAN_INFLATE_GUESS = 5, PLAN_SPLIT = 6, PLAN_SIG_BYTES_RAW = 7, PLAN_EXZD_BOUND = 8, PLAN_PACKED_IMAGE_LEN = 9 };
cudaError_t launch_rec_plan(int mode, uint64_t n, RecArrays a, const uint32_t *aux_in /*mode dependent*/, uint32_t param,
                            uint32_t *out, cudaStream_t st);
// read ids out of (prefixes of) decompressed records: len[r] = id bytes or 0xFFFFFFFF (prefix too short), src[r] = id start
cudaError_t launch_rec_ids(const uint8_t *rec, const uint64_t *rec_off, const uint32_t *have, const int32_t *status, uint64_t n,
                           uint32_t *len, uint64_t *src, cudaStream_t st);
// out[r] = rec_off[r] + sig_at[r] (absolute offset of the stored signal in the record slab)
cudaError_t launch_rec_sig_abs(const uint64_t *rec_off, RecArrays a, uint64_t n, uint64_t *out, cudaStream_t st);
cudaError_t launch_sig_extract(const uint8_t *rec, const uint64_t *rec_off, RecArrays a, uint64_t n, int16_t *sig,
                               const uint64_t *sig_off, cudaStream_t st);
// packed record r = head (from the input record) + u64 len field + signal bytes + aux (from the input record)
cudaError_t launch_rec_pack(const uint8_t *rec, const uint64_t *rec_off, RecArrays a, uint64_t n, const uint8_t *sig_src,
                            const uint64_t *sig_src_off, const uint32_t *sig_src_len, int sig_src_is_samples,
                            int sig_out_compressed, uint8_t *out, const uint64_t *out_off, cudaStream_t st, int image = 0,
                            const uint64_t *base_ptr = nullptr, const uint64_t *res = nullptr, uint64_t *abs_off = nullptr);
// merge (src/merge.c:52): every good record's read_group renumbered in place through a device table (range checked by rec_locate)
cudaError_t launch_rec_rg_remap(uint8_t *rec, const uint64_t *rec_off, RecArrays a, uint64_t n, const uint32_t *rg_map, cudaStream_t st);
// image != 0: out_off[] is the scan of PLAN_PACKED_IMAGE_LEN (align 1) and every record is written behind its u64 size
// prefix at out + *base_ptr: the packed records are the file image (no separate gather)
// file image: [u64 size][bytes] per record, back to back
// base_ptr (optional, device): added to every image offset; res (optional): the finish kernel's verdict -- an image that
// does not fit is not written; abs_off (optional, n+1 entries): absolute image offset of every record's size prefix
cudaError_t launch_image_gather(const uint8_t *src, const uint64_t *src_off, const uint32_t *len, uint64_t n, uint8_t *img,
                                const uint64_t *img_off, cudaStream_t st, const uint64_t *base_ptr = nullptr,
                                const uint64_t *res = nullptr, uint64_t *abs_off = nullptr);
// the sync-free end of a transcoding pass over one chunk (record_kernels.cu)
// res[0..2] -> mapped pinned host memory (the host reads it after the stream's next event)
cudaError_t launch_recode_publish(const uint64_t *res, uint64_t *host_mapped, cudaStream_t st);
cudaError_t launch_recode_finish(uint64_t n, const uint64_t *img_off, const uint64_t *base_ptr, uint64_t cap, const int32_t *s0,
                                 const int32_t *s1, const int32_t *s2, const int32_t *s3, uint64_t *res, cudaStream_t st);
cudaError_t launch_recode_advance(uint64_t *base_ptr, const uint64_t *res, uint64_t *acc, cudaStream_t st);
cudaError_t launch_rebase_off(uint64_t *dst, const uint64_t *src, uint64_t n, uint64_t sub, cudaStream_t st);
// content sizes of a batch of zstd frames (slow5_press.c:1206-1211): size[r], or status[r] = S5B_ERR_PRESS when the frame
// carries none; sizes above len * mul + add get S5B_ERR_NOSPACE and size 0 (the careful path sizes those exactly)
cudaError_t launch_zstd_sizes(const uint8_t *in, const uint64_t *off, const uint32_t *len, uint64_t n, uint32_t mul,
                              uint32_t add, uint32_t *size, int32_t *status, cudaStream_t st);

// lossy degradation (qts_kernels.cu; slow5_arr_qts_round, slow5_press.c:1991-2005): the `bits` low bits of every sample rounded
// away, in place.  The sample count is n_samples, or *d_n_samples (device) when that is given.
cudaError_t launch_qts_round(int16_t *sig, uint64_t n_samples, const uint64_t *d_n_samples, int bits, int num_sms, cudaStream_t st);

// SLOW5 text raw_signal column (ascii_kernels.cu): sizes, text, and back
cudaError_t launch_ascii_size(const int16_t *sig, const uint64_t *sig_off, const uint32_t *n_samples, uint64_t n_reads,
                              uint32_t *text_len, cudaStream_t st);
cudaError_t launch_ascii_format(const int16_t *sig, const uint64_t *sig_off, const uint32_t *n_samples, uint64_t n_reads,
                                uint8_t *text, const uint64_t *text_off, cudaStream_t st);
cudaError_t launch_ascii_parse(const uint8_t *text, const uint64_t *text_off, const uint32_t *text_len, uint64_t n_reads,
                               int16_t *sig, const uint64_t *sig_off, const uint32_t *expect, uint32_t *n_samples, int32_t *status,
                               cudaStream_t st);

// dense gather: scratch must hold >= compact_scratch_bytes(n_reads)
size_t compact_scratch_bytes(uint64_t n_reads);
cudaError_t launch_compact(const uint8_t *src, const uint64_t *src_off, const uint32_t *len, uint64_t n_reads,
                           uint32_t align, uint8_t *dst, uint64_t *dst_off, void *scratch, cudaStream_t st,
                           int *n_launches);

}  // namespace s5b
