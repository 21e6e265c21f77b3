/* slow5b200.h -- C-ABI of the B200-native BLOW5 per-record codec (libslow5b200.so).
 *
 * This is the drop-in boundary for the reference's per-record compress/decompress path.  Plain
 * pointers and sizes only; no C++/torch types.  Every entry point names the reference interface it
 * replaces (paths relative to the reference tree, slow5tools @ c114858 / slow5lib @ c13c4b8).
 *
 * Three levels, narrow to wide:
 *   1. per-buffer   s5b_ptr_compress_solo / s5b_ptr_depress_solo
 *                   == slow5_ptr_compress_solo / slow5_ptr_depress_solo
 *                      (slow5lib/include/slow5/slow5_press.h:118,122; slow5lib/src/slow5_press.c:330,439)
 *   2. host batch   s5b_*_batch_host: arrays of record buffers in, malloc'd buffers out -- the slot
 *                   `work_db(&core,&db,func)` fills at src/view.c:292 and the body of
 *                   slow5_get_next_batch / slow5_encode_batch (slow5lib/src/slow5_mt.c:336,353)
 *   3. device batch s5b_*_dev: slabs already resident in HBM, used by the batch scheduler and bench.
 *
 * Error convention (slow5lib/include/slow5/slow5_defs.h:137-154): 0 on success, negative on failure;
 * the codes below reuse the reference's values where one exists.
 * There is NO CPU fallback: every compute entry point returns S5B_ERR_DEVICE when no CUDA device /
 * kernel image is usable.
 */
#ifndef SLOW5B200_H
#define SLOW5B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S5B_OK            0
#define S5B_ERR_ARG      (-2)   /* == SLOW5_ERR_ARG   (slow5_defs.h) bad argument                  */
#define S5B_ERR_MEM      (-10)  /* == SLOW5_ERR_MEM   allocation failure (host or device)          */
#define S5B_ERR_PRESS    (-13)  /* == SLOW5_ERR_PRESS stream is not a valid svb-zd / zlib stream   */
#define S5B_ERR_NOSPACE  (-40)  /* output slot smaller than the worst case for that read           */
#define S5B_ERR_DEVICE   (-41)  /* no CUDA device, no sm_100a image, or a CUDA runtime error       */
#define S5B_ERR_DATASET  (-42)  /* degrade: a record's digitisation / sampling rate is not the dataset's (s5b_ctx_set_degrade) */

/* enum slow5_press_method (slow5_press.h:61-67): library enum, NOT the file byte */
#define S5B_COMPRESS_NONE   0
#define S5B_COMPRESS_ZLIB   1
#define S5B_COMPRESS_SVB_ZD 2
#define S5B_COMPRESS_ZSTD   3
#define S5B_COMPRESS_EX_ZD  4

typedef struct s5b_ctx s5b_ctx_t;   /* one per (host thread, GPU): stream, scratch, pinned staging */

/* ---- library / context ------------------------------------------------------------------ */
const char *s5b_version(void);
const char *s5b_strerror(int err);
/* number of CUDA devices visible (0 when none; never fails) */
int  s5b_device_count(void);
/* Replaces slow5_press_init / slow5_init_mt (slow5_press.c:174, slow5_mt.c:257) as the holder of
 * per-worker codec state.  device < 0 selects the current device. */
int  s5b_ctx_create(int device, s5b_ctx_t **ctx);
void s5b_ctx_destroy(s5b_ctx_t *ctx);
/* last CUDA error string seen by this context ("" if none) */
const char *s5b_ctx_last_cuda_error(const s5b_ctx_t *ctx);
/* kernels launched through this context since creation (bench.py's gpu_launches) */
uint64_t s5b_ctx_launch_count(const s5b_ctx_t *ctx);

/* ---- sizes ------------------------------------------------------------------------------- */
/* Worst-case svb-zd bytes for n int16 samples INCLUDING the 4-byte length header:
 * 4 + ceil(n/4) + 3n.  (int16 deltas zigzag to < 2^24, so the 4-byte code of
 * streamvbyte.h:31-37's 4n bound is unreachable from slow5_press.c:1082.) */
uint64_t s5b_svbzd_bound(uint32_t n_samples);
/* s5b_svbzd_bound rounded up to the 16-byte slot granule used by the device layout */
uint64_t s5b_svbzd_slot(uint32_t n_samples);

/* ---- level 3: device-resident batches ------------------------------------------------------
 * All d_* pointers are device pointers; `stream` is a cudaStream_t.  NULL selects the context's own
 * (non-blocking) stream; pass cudaStreamLegacy ((void*)0x1) to run on the legacy default stream.
 * Calls are asynchronous with respect to the host: results are valid once `stream` is synced.
 *
 * Layout contract ("slab + offsets", the device twin of slow5_batch_t's mem_records[]/mem_bytes[],
 * slow5_mt.h:23-33):
 *   d_sig        int16 slab; read r's samples start at sample index d_sig_off[r]; d_sig_off has
 *                n_reads+1 entries, every entry a multiple of 8 samples (16-byte TMA granule), and
 *                d_sig_off[r+1]-d_sig_off[r] >= n_samples[r]; the slab base is 16-byte aligned.
 *   d_svb        byte slab; read r's stream occupies [d_svb_off[r], d_svb_off[r]+d_svb_len[r]);
 *                d_svb_off has n_reads+1 entries (entry r+1 bounds slot r); any byte alignment.
 *                The slab allocation must be a multiple of 16 bytes (svb_capacity says how long).
 */

/* Replaces ptr_compress_svb_zd (slow5_press.c:1082-1115) for a whole batch: int16 -> svb-zd stream
 * [u32 N][keys ceil(N/4)][data].  d_status[r] = 0 or S5B_ERR_*; on error d_svb_len[r] = 0.
 * Requires slot r >= s5b_svbzd_bound(n_samples[r]) else S5B_ERR_NOSPACE for that read. */
int s5b_svbzd_encode_dev(s5b_ctx_t *ctx,
                         const int16_t *d_sig, const uint64_t *d_sig_off, const uint32_t *d_n_samples,
                         uint64_t n_reads,
                         uint8_t *d_svb, const uint64_t *d_svb_off, uint32_t *d_svb_len,
                         int32_t *d_status, void *stream);

/* Replaces ptr_depress_svb_zd (slow5_press.c:1143-1173).  d_n_samples[r] receives the header count,
 * d_status[r] = 0, S5B_ERR_PRESS (stream does not consume exactly d_svb_len[r]-4 bytes,
 * slow5_press.c:1130-1136), or S5B_ERR_ARG / S5B_ERR_NOSPACE (len < 4 / signal slot too small). */
int s5b_svbzd_decode_dev(s5b_ctx_t *ctx,
                         const uint8_t *d_svb, const uint64_t *d_svb_off, const uint32_t *d_svb_len,
                         uint64_t svb_capacity, uint64_t n_reads,
                         int16_t *d_sig, const uint64_t *d_sig_off, uint32_t *d_n_samples,
                         int32_t *d_status, void *stream);

/* Reads only the u32 headers: d_n_samples[r] = N of stream r (0 when len < 4). */
int s5b_svbzd_peek_dev(s5b_ctx_t *ctx, const uint8_t *d_svb, const uint64_t *d_svb_off,
                       const uint32_t *d_svb_len, uint64_t n_reads, uint32_t *d_n_samples, void *stream);

/* ex-zd signal codec (slow5_press.c:1236-1848): QTS shift + 16-bit zigzag-delta + one byte per value with an exception
 * list.  Same slab contract as the svb-zd pair; d_out slots must hold s5b_exzd_bound(n) = 2n + 1024 bytes, the size of
 * the reference's own working buffer (:1728) -- a read whose stream would not fit makes the reference abort
 * (SLOW5_ASSERT) and gets S5B_ERR_PRESS here; an empty read (undefined in the reference) gets S5B_ERR_ARG.
 * Replaces ptr_compress_ex_zd (:1778) / ptr_depress_ex_zd (:1824); bytes identical to the reference's. */
uint64_t s5b_exzd_bound(uint32_t n_samples);
uint64_t s5b_exzd_slot(uint32_t n_samples);
int s5b_exzd_encode_dev(s5b_ctx_t *ctx, const int16_t *d_sig, const uint64_t *d_sig_off, const uint32_t *d_n_samples,
                        uint64_t n_reads, uint8_t *d_out, const uint64_t *d_out_off, uint32_t *d_out_len,
                        int32_t *d_status, void *stream);
/* d_n_samples[r] receives the header's sample count; S5B_ERR_PRESS for an unsupported version, sections that overrun
 * the stream or do not consume their stated length (:1492-1500), exception positions outside the read. */
int s5b_exzd_decode_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                        uint64_t in_capacity, uint64_t n_reads, int16_t *d_sig, const uint64_t *d_sig_off,
                        uint32_t *d_n_samples, int32_t *d_status, void *stream);

/* Replaces ptr_depress_zlib_solo (slow5_press.c:973-1010: inflateInit2(15) + inflate loop) for a batch of
 * independent zlib streams (one per record, slow5.c:4046).  Stream r = d_in[d_in_off[r] .. +d_in_len[r])
 * (any alignment; d_in base 16-byte aligned, in_capacity a multiple of 16); its output goes to the slot
 * [d_out_off[r], d_out_off[r+1]).  d_status[r]: 0; S5B_ERR_PRESS for what zlib reports as Z_DATA_ERROR /
 * Z_NEED_DICT; S5B_ERR_NOSPACE when the slot is too small -- d_out_len[r] then holds the size the stream
 * needs.  As in the reference, input that ends early is not an error: the bytes decoded so far are returned. */
int s5b_zlib_inflate_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                         uint64_t in_capacity, uint64_t n_reads, uint8_t *d_out, const uint64_t *d_out_off,
                         uint32_t *d_out_len, int32_t *d_status, void *stream);

/* Replaces ptr_depress_zstd (slow5_press.c:1205-1230: ZSTD_getFrameContentSize + ZSTD_decompress) for a batch of
 * independent single-frame streams.  Same slab contract as s5b_zlib_inflate_dev.  Every frame must carry its
 * content size (the reference refuses frames that do not, :1206-1211); s5b_zstd_content_size() reads it on the
 * host so slots can be sized exactly.  S5B_ERR_PRESS for anything libzstd rejects; S5B_ERR_NOSPACE (d_out_len[r] =
 * content size) when the slot is too small. */
int s5b_zstd_decode_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                        uint64_t in_capacity, uint64_t n_reads, uint8_t *d_out, const uint64_t *d_out_off,
                        uint32_t *d_out_len, int32_t *d_status, void *stream);
/* content size stored in a frame header; returns 0 and *size, or S5B_ERR_PRESS (not a zstd frame / no size) */
int s5b_zstd_content_size(const void *frame, size_t len, uint64_t *size);

/* Replaces ptr_compress_zlib / ptr_compress_zlib_solo (slow5_press.c:837-913: deflate at level 6 with
 * Z_FINISH) for a batch: record r = d_in[d_in_off[r] .. +d_in_len[r]) becomes one complete zlib stream
 * (78 9C .. Adler-32) in the slot [d_out_off[r], d_out_off[r+1]), which must hold s5b_zlib_bound(len).
 * The bytes differ from zlib's (dynamic Huffman + distance-1 run matches); any zlib inflates them to the
 * exact input.  d_split (optional, may be NULL): byte offset inside record r where a new Huffman block
 * should start -- for BLOW5 records the start of the svb-zd data bytes, whose statistics differ from the
 * header + key bytes before them.
 * Size contract: <= 1.03 x zlib level 6 on records whose signal is svb-zd or ex-zd coded (measured 0.98 x).  Records with
 * an UNCOMPRESSED signal (-s none) are outside that contract: the encoder matches runs only, zlib's LZ77 finds ~13 % more
 * there; the output is still a valid stream. */
uint64_t s5b_zlib_bound(uint64_t len);
int s5b_zlib_deflate_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                         uint64_t in_capacity, const uint32_t *d_split, uint64_t n_reads, uint8_t *d_out,
                         const uint64_t *d_out_off, uint32_t *d_out_len, int32_t *d_status, void *stream);

/* Replaces ptr_compress_zstd (slow5_press.c:1183-1202: ZSTD_compress at level 1) for a batch: record r becomes one
 * single-segment zstd frame with its content size in the header (what ptr_depress_zstd requires, :1206-1211) in
 * the slot [d_out_off[r], d_out_off[r+1]), which must hold s5b_zstd_bound(len).  Same argument contract as
 * s5b_zlib_deflate_dev, split hint included.  The bytes differ from libzstd's (Huffman-coded literals, no
 * sequences); any zstd decoder regenerates the exact input. */
uint64_t s5b_zstd_bound(uint64_t len);
int s5b_zstd_encode_dev(s5b_ctx_t *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len,
                        uint64_t in_capacity, const uint32_t *d_split, uint64_t n_reads, uint8_t *d_out,
                        const uint64_t *d_out_off, uint32_t *d_out_len, int32_t *d_status, void *stream);

/* Gathers slotted streams into a dense slab: d_dst_off[r] (n_reads+1 entries, exclusive scan of
 * len rounded up to `align`, align in {1,16}) and the copied bytes.  d_dst capacity is checked
 * against dst_capacity (S5B_ERR_NOSPACE is reported through the return of the host wrappers). */
int s5b_compact_dev(s5b_ctx_t *ctx, const uint8_t *d_src, const uint64_t *d_src_off,
                    const uint32_t *d_len, uint64_t n_reads, uint32_t align,
                    uint8_t *d_dst, uint64_t *d_dst_off, void *stream);

/* ---- level 2: host batches (the work_db / slow5_mt slot) ----------------------------------- */
/* Slab form: h_sig/h_sig_off/h_n_samples as in the device contract but in host memory (pinned or
 * pageable; pinned avoids a staging copy).  h_svb receives the DENSE streams, stream r at
 * h_svb_off[r] (n_reads+1 entries written by the call, 16-byte granule), h_svb_len[r] bytes.
 * Returns 0, or the first non-zero per-read status / a call-level error. */
int s5b_svbzd_encode_host(s5b_ctx_t *ctx,
                          const int16_t *h_sig, const uint64_t *h_sig_off, const uint32_t *h_n_samples,
                          uint64_t n_reads,
                          uint8_t *h_svb, uint64_t h_svb_capacity, uint64_t *h_svb_off, uint32_t *h_svb_len,
                          int32_t *h_status);
int s5b_svbzd_decode_host(s5b_ctx_t *ctx,
                          const uint8_t *h_svb, const uint64_t *h_svb_off, const uint32_t *h_svb_len,
                          uint64_t n_reads,
                          int16_t *h_sig, uint64_t h_sig_capacity /*samples*/, uint64_t *h_sig_off,
                          uint32_t *h_n_samples, int32_t *h_status);

/* Pointer-array form, the exact shape of db_t / slow5_batch_t: n buffers in, n malloc()'d buffers
 * out (caller free()s each), NULL + out_n[i]=0 for a failed record.  `method` is a
 * S5B_COMPRESS_* value (NONE is a copy; ZLIB, SVB_ZD, ZSTD and EX_ZD run on the GPU). */
int s5b_compress_batch_host(s5b_ctx_t *ctx, int method, const void *const *ptrs, const size_t *counts,
                            size_t n, void **out_ptrs, size_t *out_n);
int s5b_depress_batch_host(s5b_ctx_t *ctx, int method, const void *const *ptrs, const size_t *counts,
                           size_t n, void **out_ptrs, size_t *out_n);
/* Record compression (the slow5_ptr_compress(record_press) call of slow5_rec_to_mem, slow5.c:4050) for a batch of
 * packed records, method S5B_COMPRESS_ZLIB or S5B_COMPRESS_ZSTD, with the per-record Huffman-block split hint of
 * s5b_zlib_deflate_dev / s5b_zstd_encode_dev (splits may be NULL). */
int s5b_compress_records_host(s5b_ctx_t *ctx, int method, const void *const *ptrs, const size_t *counts,
                              const uint32_t *splits, size_t n, void **out_ptrs, size_t *out_n);

/* Whole-batch BLOW5 record transcoding with everything between the two copies on the device: the body of
 * slow5_convert_parallel's batch loop (src/view.c:254-301: work_db over depress_parse_rec_to_mem) for
 * blow5 -> blow5 conversions.  h_in holds n packed records exactly as stored in the file (record i =
 * h_in[rec_off[i] .. +rec_len[i]), size prefixes excluded), compressed with (in_rec, in_sig); h_out receives the
 * output FILE IMAGE -- [u64 size][record bytes] per record, in order -- compressed with (out_rec, out_sig), ready
 * for one fwrite.  Methods: S5B_COMPRESS_NONE / ZLIB / ZSTD for records, NONE / SVB_ZD / EX_ZD for signals.  Pinned h_in / h_out
 * (s5b_host_alloc) avoid staging copies.  Returns 0, the first per-record error, or S5B_ERR_NOSPACE with *out_bytes =
 * bytes needed when out_cap is too small. */
int s5b_blow5_recode_host(s5b_ctx_t *ctx, int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *h_in,
                          uint64_t in_bytes, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n,
                          uint8_t *h_out, uint64_t out_cap, uint64_t *out_bytes);
/* The same transcoding for batches of any size, as a 3-lane CUDA-stream pipeline (the replacement of the work_db pool of
 * src/thread.c:114 around src/view.c:35-57): the batch is cut into chunks of <= 24 Ki records / 256 MiB, and the H2D copy
 * of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 run concurrently; between its two copies a chunk never
 * synchronises with the host.  s5b_blow5_recode_host is this call with out_img_off = NULL.  out_img_off (optional,
 * n+1 entries) receives the offset of every record's u64 size prefix inside h_out (entry n = *out_bytes): record i of the
 * output is h_out[out_img_off[i] + 8 .. out_img_off[i+1]).  The record table must be ascending for the batch to be
 * chunked (a table in any other order is processed as one chunk). */
int s5b_blow5_recode_batch_host(s5b_ctx_t *ctx, int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *h_in,
                                uint64_t in_bytes, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n,
                                uint8_t *h_out, uint64_t out_cap, uint64_t *out_bytes, uint64_t *out_img_off);
/* Device-resident form: payload d_in and image d_out live in HBM (d_in must be 16-byte aligned -- S5B_ERR_ARG otherwise -- and
 * its allocation must extend to in_bytes rounded up to 16), the record table rec_off / rec_len is HOST memory (metadata the caller produced when it laid the batch out).  The
 * call only enqueues work on the context's transcoding stream (s5b_ctx_recode_stream) and returns; after s5b_ctx_sync
 * d_result[0] = image bytes written, d_result[1] = first error as a sign-extended S5B_ERR_* (0 = none; S5B_ERR_NOSPACE also
 * when a record inflates to more than 4x + 1 KiB of its stored size -- the host form retries those, this form cannot --
 * or the image outgrows out_cap).  d_img_off: optional device array of n+1 image offsets as above. */
int s5b_blow5_recode_dev(s5b_ctx_t *ctx, int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *d_in,
                         uint64_t in_bytes, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n, uint8_t *d_out,
                         uint64_t out_cap, uint64_t *d_result, uint64_t *d_img_off);
/* Workspace of the device-resident form.  A pass is cut into chunks, and every chunk ends with a partly idle GPU (the entropy
 * kernels hand out records in rounds of 32 per warp), so s5b_blow5_recode_dev takes the largest chunks its workspace budget
 * allows: by default 60 % of the device memory that is free at the call plus what the context already holds (the slabs stay with
 * the context until s5b_ctx_destroy).  Worst-case slabs are sized from the stored bytes alone: about 65 KB per 4096-sample record
 * in the none -> zlib + svb-zd direction, 46 KB in the opposite one.  max_bytes > 0 fixes the budget instead; 0 = automatic.
 * (The environment variables S5B_RECODE_DEV_CHUNK / S5B_RECODE_DEV_CHUNK_MB, read at s5b_ctx_create, fix the chunk caps.) */
int s5b_ctx_set_recode_workspace(s5b_ctx_t *ctx, uint64_t max_bytes);
/* waits for everything the context has enqueued */
int s5b_ctx_sync(s5b_ctx_t *ctx);
/* the cudaStream_t the device-resident transcoder enqueues on (for callers that order their own work against it) */
void *s5b_ctx_recode_stream(s5b_ctx_t *ctx);
/* Per-stage device timing of the transcoder: when enabled every stage of every chunk is bracketed by CUDA events on its
 * stream; s5b_ctx_stage_report waits for them and returns accumulated milliseconds and launch-group counts per stage
 * (arrays of s5b_stage_count() entries, named by s5b_stage_name), optionally resetting the totals. */
int s5b_ctx_stage_timing(s5b_ctx_t *ctx, int enable);
int s5b_ctx_stage_report(s5b_ctx_t *ctx, double *ms, uint64_t *count, int reset);
int s5b_stage_count(void);
const char *s5b_stage_name(int stage);

/* The auxiliary columns of the file whose records the transcoder is about to see (slow5_aux_meta_t, slow5.h:198-213): element
 * size in bytes (1..8) of every field in header order and whether it is an array (stored as a u64 count and count elements).
 * With a layout set, a record whose auxiliary section is not exactly these fields fails with S5B_ERR_PRESS, like
 * slow5_rec_aux_parse does (slow5.c:3088-3166); without one (the default, or n_fields = 0xffffffff) the section is carried as it
 * is.  Conversions that keep both methods copy the stored records without opening them either way. */
int s5b_ctx_set_aux_layout(s5b_ctx_t *ctx, const uint8_t *elem_size, const uint8_t *is_array, uint32_t n_fields);

/* merge (src/merge.c:52, `read->read_group = list[file][read->read_group]`): while a table is set, s5b_blow5_recode_batch_host /
 * s5b_blow5_recode_host write every record with read_group map[read_group]; a record whose read_group is >= n fails with
 * S5B_ERR_PRESS.  n = 0 clears the table.  Not available to s5b_blow5_recode_dev (the renumbering is done in the library's own
 * copy of the records): S5B_ERR_ARG there. */
int s5b_ctx_set_rg_map(s5b_ctx_t *ctx, const uint32_t *map, uint32_t n);

/* degrade (src/degrade.c:240-263: the view worker with slow5_rec_qts_round between decode and re-encode).  While bits is 1..16,
 * s5b_blow5_recode_batch_host / _host / _dev round the `bits` least significant bits of every sample away (slow5_arr_qts_round,
 * slow5lib/src/slow5_press.c:1965-2005: to the nearest multiple of 2^bits, halves up, stored back as int16) before the signal is
 * re-encoded -- also when the signal method stays the same.  check_dataset != 0 is the reference's `-b auto` rule
 * (slow5_reccmp, src/degrade.c:195-211): a record whose digitisation or sampling_rate differs from the given values fails with
 * S5B_ERR_DATASET.  bits = 0 switches the step off. */
int s5b_ctx_set_degrade(s5b_ctx_t *ctx, int bits, int check_dataset, float digitisation, float sampling_rate);
/* slow5_arr_qts_round (slow5lib/src/slow5_extra.h:152, slow5_press.c:1991-2005) for a sample slab in HBM, in place, enqueued on
 * `stream` (a cudaStream_t; NULL = the context's stream); bits 1..16 (0: nothing to do, like the reference). */
int s5b_qts_round_dev(s5b_ctx_t *ctx, int16_t *d_sig, uint64_t n_samples, int bits, void *stream);
/* The same for a batch of host arrays: array i holds counts[i] BYTES of int16 samples (any alignment); out_ptrs[i] receives a
 * malloc'd array of the same size (the caller frees), out_n[i] its size in bytes.  One H2D, one launch, one D2H. */
int s5b_qts_round_batch_host(s5b_ctx_t *ctx, int bits, const void *const *ptrs, const size_t *counts, size_t n, void **out_ptrs,
                             size_t *out_n);
/* The per-record work of index building (slow5_idx_build, slow5lib/src/slow5_idx.c:283-334) for a batch: the read_id of
 * every stored record.  Records compressed with in_rec (S5B_COMPRESS_NONE / ZLIB / ZSTD) are decompressed on the device --
 * for zlib only their first 256 bytes, like the reference's partial decompression (:290-310), with a full pass for the
 * records whose id does not fit in that prefix (:312-320) -- and only the id bytes come back: record i's id is
 * h_ids[id_off[i] .. id_off[i+1]) (no terminator).  ctx may be NULL for S5B_COMPRESS_NONE. */
int s5b_blow5_read_ids_host(s5b_ctx_t *ctx, int in_rec, const uint8_t *h_in, uint64_t in_bytes, const uint64_t *rec_off,
                            const uint32_t *rec_len, uint64_t n, uint8_t *h_ids, uint64_t ids_cap, uint64_t *id_off);
/* The raw_signal column of SLOW5 text records, the other per-sample loop of `view` (slow5_rec_to_mem's sprintf("%d,") loop,
 * slow5.c:3866-3878, and slow5_rec_parse's strsep + slow5_ato_int16 loop, slow5.c:2754-2778), for a batch:
 *   s5b_signal_to_ascii_batch_host: record i's STORED signal bytes (sig_method S5B_COMPRESS_NONE = raw int16, SVB_ZD or EX_ZD:
 *     decoded on the device first) -> malloc()'d, NUL-terminated "v0,v1,...,vN-1" (out_n[i] = characters, no trailing comma);
 *   s5b_ascii_to_signal_batch_host: text of out_n characters -> malloc()'d int16 samples; expect[i] is the record's
 *     len_raw_signal column, a different count or a token the reference rejects (empty, leading zero, a character other than
 *     digits and '-', outside int16: slow5_misc.c:122-139, :303-319) gives S5B_ERR_ARG for that record (NULL pointer). */
int s5b_signal_to_ascii_batch_host(s5b_ctx_t *ctx, int sig_method, const void *const *ptrs, const size_t *counts, size_t n,
                                   char **out_ptrs, size_t *out_n);
int s5b_ascii_to_signal_batch_host(s5b_ctx_t *ctx, const char *const *ptrs, const size_t *counts, const uint64_t *expect, size_t n,
                                   int16_t **out_ptrs, size_t *out_n);
/* page-locked host memory for the slab entry points (cudaHostAlloc / cudaFreeHost) */
void *s5b_host_alloc(size_t bytes);
void s5b_host_free(void *p);

/* ---- level 1: single buffers ---------------------------------------------------------------
 * Same contract as slow5_ptr_compress_solo / slow5_ptr_depress_solo: returns a malloc()'d buffer,
 * *n its size; NULL and *n = 0 on failure (s5b_last_error() holds the code, the twin of the
 * reference's thread-local slow5_errno).  A batch of one: correct but latency-bound; callers that
 * care about throughput use the batch forms.  Uses a lazily created per-thread context. */
void *s5b_ptr_compress_solo(int method, const void *ptr, size_t count, size_t *n);
void *s5b_ptr_depress_solo(int method, const void *ptr, size_t count, size_t *n);
int   s5b_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SLOW5B200_H */
