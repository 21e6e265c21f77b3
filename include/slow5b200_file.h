/* slow5b200_file.h -- the slice of the slow5lib C API that the per-record codec path lives behind, provided by
 * libslow5b200.so with the codec work done on the GPU.  Names mirror the reference one-to-one (s5b_ for
 * slow5_); `#define S5B_SLOW5_COMPAT` before including maps the slow5_* spellings onto them so code written
 * against slow5lib's low-level API (slow5lib/examples/adv/sequential_read_pthreads.c) compiles unchanged.
 *
 *   s5b_open            slow5_open            slow5lib/include/slow5/slow5.h:345   (slow5.c:295)
 *   s5b_close           slow5_close           slow5.h:354                            (slow5.c:505)
 *   s5b_get_next_mem    slow5_get_next_mem    slow5_extra.h                          (slow5.c:3206)
 *   s5b_get_next_bytes  slow5_get_next_bytes  slow5.h:656                            (slow5.c:3302)
 *   s5b_decode          slow5_decode          slow5.h:658                            (slow5.c:2613)
 *   s5b_encode          slow5_encode          slow5.h:660                            (slow5.c:4083)
 *   s5b_write_bytes     slow5_write_bytes     slow5.h:662                            (slow5.c:3785)
 *   s5b_get_next / s5b_get / s5b_write        slow5_get_next / slow5_get / slow5_write   slow5.h:440 / :423 / :600 (one record each)
 *   s5b_aux_get_* / s5b_hdr_get               slow5_aux_get_* / slow5_hdr_get            slow5.h:469-508 / :396
 *   s5b_rec_free        slow5_rec_free        slow5.h:454
 *   s5b_set_press       slow5_set_press       slow5.h:612
 *   s5b_hdr_write       slow5_hdr_write       slow5.h:586
 *   s5b_decode_batch / s5b_encode_batch       the worker bodies of slow5_get_next_batch / slow5_encode_batch
 *                                             (slow5_mt.c:124-181), one GPU batch instead of a pthread pool
 *
 * Differences a caller can see: slow5_rec_t's aux_map (a khash) is replaced by the record's binary auxiliary
 * section kept as-is (aux / aux_len) -- the slow5_aux_get_* accessors work on it as they do on the map --; the leading fields
 * have the reference's names, types and order (slow5.h:274-287).  Ownership rules are the reference's: s5b_decode frees *mem and replaces it when the
 * record was compressed (slow5.c:2595-2597); s5b_encode output carries the 8-byte size prefix, s5b_get_next_mem
 * output does not; everything returned is malloc()'d.
 */
#ifndef SLOW5B200_FILE_H
#define SLOW5B200_FILE_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include "slow5b200.h"
#include "slow5b200_press.h"
#ifdef __cplusplus
extern "C" {
#endif

#define S5B_ERR_EOF       (-1)   /* == SLOW5_ERR_EOF */
#define S5B_ERR_IO        (-5)   /* == SLOW5_ERR_IO */
#define S5B_ERR_RECPARSE  (-4)   /* == SLOW5_ERR_RECPARSE (slow5_defs.h:140) */
#define S5B_ERR_NOIDX     (-6)   /* == SLOW5_ERR_NOIDX: s5b_get before s5b_idx_load */
#define S5B_ERR_NOTFOUND  (-7)   /* == SLOW5_ERR_NOTFOUND: read id not in the index */
#define S5B_ERR_NOAUX     (-11)  /* == SLOW5_ERR_NOAUX: the record has no auxiliary fields */
#define S5B_ERR_NOFLD     (-12)  /* == SLOW5_ERR_NOFLD: no auxiliary field of that name */
#define S5B_ERR_TYPE      (-17)  /* == SLOW5_ERR_TYPE: the field has another type */

typedef struct s5b_rec {
    uint16_t read_id_len;
    char *read_id;
    uint32_t read_group;
    double digitisation;
    double offset;
    double range;
    double sampling_rate;
    uint64_t len_raw_signal;
    int16_t *raw_signal;
    uint8_t *aux;      /* binary auxiliary section of the record, header order (slow5.c:3993-4044) */
    uint64_t aux_len;
    void *aux_meta;    /* private: where the fields of `aux` lie (set by s5b_decode*; NULL in records a caller builds) */
} s5b_rec_t;

/* The leading fields of the file handle are public and laid out like the reference's slow5_file_t
 * (slow5lib/include/slow5/slow5.h:291-318): callers read fp->compress->record_press->method, fp->format, fp->meta.pathname
 * ... directly (src/view.c:43-49, slow5_mt.c:163).  `header` starts like slow5_hdr_t (version, num_read_groups); the
 * attribute and auxiliary-field tables behind it are private (the reference keeps khash maps there) and are reached through
 * this library's calls only. */
typedef struct s5b_hdr {
    struct { uint8_t major, minor, patch; } version;     /* struct slow5_version, slow5.h:168-172 */
    uint32_t num_read_groups;
} s5b_hdr_t;
typedef struct s5b_file_meta {                           /* slow5_file_meta_t, slow5.h:291-298 */
    const char *pathname;
    int fd;
    uint64_t start_rec_offset;
    char *fread_buffer;                                  /* always NULL here */
    const char *mode;
} s5b_file_meta_t;
typedef struct s5b_file {
    FILE *fp;
    int format;                  /* enum slow5_fmt: 0 unknown, 1 ASCII (SLOW5), 2 binary (BLOW5) */
    s5b_press_t *compress;       /* methods of the file (read) / of the records to be written; NULL for ASCII */
    s5b_hdr_t *header;
    void *index;                 /* NULL: random access goes through `slow5tools-b200 get` */
    s5b_file_meta_t meta;
} s5b_file_t;                    /* the library's private state follows these fields */

extern int s5b_errno_value(void);   /* the twin of the thread-local slow5_errno */

s5b_file_t *s5b_open(const char *pathname, const char *mode /* "r", "w" (.blow5 or .slow5) or "a" (append to a BLOW5 file) */);
int s5b_close(s5b_file_t *fp);                     /* "w": appends the end-of-file marker first */
/* "w" files: take the header (attributes, aux columns, version) from an opened input, choose the methods */
int s5b_hdr_copy(s5b_file_t *dst, const s5b_file_t *src);
int s5b_set_press(s5b_file_t *fp, int rec_press /*S5B_COMPRESS_*/, int sig_press);
int s5b_hdr_write(s5b_file_t *fp);
int s5b_file_record_press(const s5b_file_t *fp);
int s5b_file_signal_press(const s5b_file_t *fp);

void *s5b_get_next_mem(size_t *n, s5b_file_t *fp);
int s5b_get_next_bytes(char **mem, size_t *bytes, s5b_file_t *fp);
int s5b_decode(char **mem, size_t *bytes, s5b_rec_t **read, s5b_file_t *fp);
int s5b_encode(char **mem, size_t *bytes, s5b_rec_t *read, s5b_file_t *fp);
int s5b_write_bytes(char *mem, size_t bytes, s5b_file_t *fp);   /* 0, or negative (slow5.c:3785-3794) */
void s5b_rec_free(s5b_rec_t *read);
/* the single-record conveniences (slow5.h:423, :440, :600): slow5_get_next = get_next_bytes + decode, slow5_get = index lookup
 * (after s5b_idx_load; S5B_ERR_NOIDX / S5B_ERR_NOTFOUND) + pread + decode, slow5_write = encode + write_bytes (bytes written or -1) */
int s5b_get_next(s5b_rec_t **read, s5b_file_t *fp);
int s5b_get(const char *read_id, s5b_rec_t **read, s5b_file_t *fp);
int s5b_write(s5b_rec_t *read, s5b_file_t *fp);

/* Auxiliary fields of a decoded record and header attributes -- the accessors that stand in for slow5_rec_t's aux_map
 * (slow5lib/include/slow5/slow5.h:396, :469-508; slow5.c:1383-1400, :3493-3660).  Same names, arguments and results: a primitive
 * comes back by value (the type's NULL value -- INT8_MAX ... UINT64_MAX, NaN, 0 -- on error), an array as a pointer into the
 * record with *len elements (strings NUL-terminated; NULL with *len = 0 for a value marked missing, which is not an error);
 * *err (optional) and the thread's errno receive 0 or S5B_ERR_ARG / _NOAUX / _NOFLD / _TYPE.  The pointers live as long as the
 * record does; s5b_hdr_get's as long as the file. */
int8_t s5b_aux_get_int8(const s5b_rec_t *read, const char *field, int *err);
int16_t s5b_aux_get_int16(const s5b_rec_t *read, const char *field, int *err);
int32_t s5b_aux_get_int32(const s5b_rec_t *read, const char *field, int *err);
int64_t s5b_aux_get_int64(const s5b_rec_t *read, const char *field, int *err);
uint8_t s5b_aux_get_uint8(const s5b_rec_t *read, const char *field, int *err);
uint16_t s5b_aux_get_uint16(const s5b_rec_t *read, const char *field, int *err);
uint32_t s5b_aux_get_uint32(const s5b_rec_t *read, const char *field, int *err);
uint64_t s5b_aux_get_uint64(const s5b_rec_t *read, const char *field, int *err);
float s5b_aux_get_float(const s5b_rec_t *read, const char *field, int *err);
double s5b_aux_get_double(const s5b_rec_t *read, const char *field, int *err);
char s5b_aux_get_char(const s5b_rec_t *read, const char *field, int *err);
uint8_t s5b_aux_get_enum(const s5b_rec_t *read, const char *field, int *err);
int8_t *s5b_aux_get_int8_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
int16_t *s5b_aux_get_int16_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
int32_t *s5b_aux_get_int32_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
int64_t *s5b_aux_get_int64_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
uint8_t *s5b_aux_get_uint8_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
uint16_t *s5b_aux_get_uint16_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
uint32_t *s5b_aux_get_uint32_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
uint64_t *s5b_aux_get_uint64_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
float *s5b_aux_get_float_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
double *s5b_aux_get_double_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
char *s5b_aux_get_string(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
uint8_t *s5b_aux_get_enum_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err);
char *s5b_hdr_get(const char *attr, uint32_t read_group, const s5b_hdr_t *header);
/* header / index introspection (slow5.h:633-654): the attribute keys (sorted; a malloc()'d array the caller frees, the strings
 * stay the library's), the auxiliary column names and types (enum slow5_aux_type values; the library's arrays), the labels of an
 * enum column, the read ids of a loaded index in index order (S5B_ERR_NOIDX before s5b_idx_load) */
const char **s5b_get_hdr_keys(const s5b_hdr_t *header, uint64_t *len);
char **s5b_get_aux_names(const s5b_hdr_t *header, uint64_t *len);
int *s5b_get_aux_types(const s5b_hdr_t *header, uint64_t *len);
char **s5b_get_aux_enum_labels(const s5b_hdr_t *header, const char *field, uint8_t *n);
char **s5b_get_rids(const s5b_file_t *fp, uint64_t *len);
/* Building a file from scratch (slow5.h:404-447, :510-545), the calls slow5lib/examples/write.c makes: header attributes and
 * auxiliary columns are added to the header of a file opened with "w" before s5b_hdr_write; a record is made with s5b_rec_init,
 * its public fields are filled in by the caller (read_id and raw_signal malloc()'d: s5b_rec_free frees them), its auxiliary values
 * are set one by one -- a column that gets none is written as the type's NULL value / an empty array -- and s5b_write encodes
 * and appends it.  Return values are the reference's: 0, or -1 (bad argument), -2 (attribute / column exists, or no such
 * column), -3 (s5b_aux_add with an enum type: use s5b_aux_add_enum; array / primitive mismatch), -4 (an enum value beyond the
 * column's labels, a label that is not a C identifier). */
s5b_rec_t *s5b_rec_init(void);
int s5b_hdr_add(const char *attr, s5b_hdr_t *header);
int s5b_hdr_set(const char *attr, const char *value, uint32_t read_group, s5b_hdr_t *header);
int64_t s5b_hdr_add_rg(s5b_hdr_t *header);
int s5b_aux_add(const char *field, int type /* enum slow5_aux_type */, s5b_hdr_t *header);
int s5b_aux_add_enum(const char *field, const char **enum_labels, uint8_t num_labels, s5b_hdr_t *header);   /* -4: bad label */
int s5b_aux_set(s5b_rec_t *read, const char *field, const void *data, s5b_hdr_t *header);
int s5b_aux_set_array(s5b_rec_t *read, const char *field, const void *data, size_t len, s5b_hdr_t *header);
int s5b_aux_set_string(s5b_rec_t *read, const char *field, const char *data, s5b_hdr_t *header);

/* n records at once: mems[i]/bytes[i] as returned by s5b_get_next_mem; reads[i] allocated when NULL */
int s5b_decode_batch(s5b_file_t *fp, char **mems, size_t *bytes, size_t n, s5b_rec_t **reads);
int s5b_encode_batch(s5b_file_t *fp, s5b_rec_t **reads, size_t n, char **mems, size_t *bytes);

/* ---- the "easy multi-thread" batch API (slow5lib/include/slow5/slow5_mt.h:23-65, slow5lib/src/slow5_mt.c:202-400), the
 * interface pyslow5 uses.  Same structs, field order and call shapes; `num_thread` keeps its place in the signature but
 * the codec stage of a batch is one pass over the GPU (s5b_decode_batch / s5b_encode_batch), not a fork-join pool.
 *   s5b_get_next_batch : up to num_reads records fetched (serial, like slow5_mt.c:84-107) and decoded into
 *                        batch->slow5_rec[]; returns the number of records, < num_reads at the end of the file
 *   s5b_encode_batch_mt: batch->slow5_rec[] -> batch->mem_records[] / mem_bytes[] (slow5_encode_batch, :353)
 *   s5b_write_batch    : encode + write, returns the number of records written (slow5_write_batch, :359)
 * Errors: a negative S5B_ERR_* return (the reference exits the process instead, slow5_mt.c:131-137). */
typedef struct {
    int32_t n_rec;
    int32_t capacity_rec;
    char **mem_records;
    size_t *mem_bytes;
    s5b_rec_t **slow5_rec;
    char **rid; /* the read ids of the last s5b_get_batch call (caller's array) */
} s5b_batch_t;
typedef struct {
    s5b_file_t *sf;
    int num_thread;
} s5b_mt_t;
s5b_mt_t *s5b_init_mt(int num_thread, s5b_file_t *fp);
s5b_batch_t *s5b_init_batch(int batch_capacity);
int s5b_get_next_batch(s5b_mt_t *mt, s5b_batch_t *batch, int num_reads);
int s5b_encode_batch_mt(s5b_mt_t *mt, s5b_batch_t *batch, int num_reads);
int s5b_write_batch(s5b_mt_t *mt, s5b_batch_t *batch, int num_reads);
/* slow5_idx_load (slow5.h:560) / slow5_get_batch (slow5_mt.h:52, slow5_mt.c:319-333): random access by read id.  The index is
 * FILE.idx as written by `slow5tools-b200 index` (byte-identical to the reference's); a missing one is created first, like the
 * reference does (slow5.c:4152-4169; for a compressed file that is a pass over the GPU), S5B_ERR_IO when that fails.
 * s5b_get_batch fetches the num_rid records with pread() and decodes them as ONE GPU batch into batch->slow5_rec[];
 * returns num_rid, or S5B_ERR_ARG when an id is not in the index (the reference exits the process). */
int s5b_idx_load(s5b_file_t *fp);
void s5b_idx_unload(s5b_file_t *fp);   /* slow5_idx_unload, slow5.h:382 */
int s5b_get_batch(s5b_mt_t *mt, s5b_batch_t *batch, char **rid, int num_rid);
void s5b_free_batch(s5b_batch_t *batch);
void s5b_free_mt(s5b_mt_t *mt);
/* the "lazy" forms pyslow5 calls (slow5_mt.h:59-65): no mt / batch objects to keep; *read receives a malloc()'d array of batch_size
 * (num_rid) record pointers of which the first <return value> are set, to be released with s5b_free_batch_lazy */
int s5b_get_next_batch_lazy(s5b_rec_t ***read, s5b_file_t *fp, int batch_size, int num_threads);
int s5b_get_batch_lazy(s5b_rec_t ***read, s5b_file_t *fp, char **rid, int num_rid, int num_threads);
int s5b_write_batch_lazy(s5b_rec_t **read, s5b_file_t *fp, int batch_size, int num_threads);
void s5b_free_batch_lazy(s5b_rec_t ***read, int num_rec);

#ifdef S5B_SLOW5_COMPAT
#define slow5_press_method_t s5b_press_method_t
#define slow5_press_t s5b_press_t
#define __slow5_press __s5b_press
#define slow5_press s5b_press
#define slow5_press_init s5b_press_init
#define __slow5_press_init __s5b_press_init
#define slow5_press_free s5b_press_free
#define __slow5_press_free __s5b_press_free
#define slow5_ptr_compress s5b_ptr_compress
#define slow5_ptr_depress s5b_ptr_depress
#define slow5_ptr_compress_solo s5b_ptr_compress_solo
#define slow5_ptr_depress_solo s5b_ptr_depress_solo
#define slow5_compress_footer_next s5b_compress_footer_next
#define slow5_batch_t s5b_batch_t
#define slow5_mt_t s5b_mt_t
#define slow5_init_mt s5b_init_mt
#define slow5_init_batch s5b_init_batch
#define slow5_get_next_batch s5b_get_next_batch
#define slow5_encode_batch s5b_encode_batch_mt
#define slow5_write_batch s5b_write_batch
#define slow5_get_batch s5b_get_batch
#define slow5_idx_load s5b_idx_load
#define slow5_idx_unload s5b_idx_unload
#define slow5_free_batch s5b_free_batch
#define slow5_free_mt s5b_free_mt
#define slow5_get_next_batch_lazy s5b_get_next_batch_lazy
#define slow5_get_batch_lazy s5b_get_batch_lazy
#define slow5_write_batch_lazy s5b_write_batch_lazy
#define slow5_free_batch_lazy s5b_free_batch_lazy
#define slow5_file_t s5b_file_t
#define slow5_rec_t s5b_rec_t
#define slow5_open s5b_open
#define slow5_close s5b_close
#define slow5_get_next_mem s5b_get_next_mem
#define slow5_get_next_bytes s5b_get_next_bytes
#define slow5_decode s5b_decode
#define slow5_encode s5b_encode
#define slow5_write_bytes s5b_write_bytes
#define slow5_get_next s5b_get_next
#define slow5_get s5b_get
#define slow5_write s5b_write
#define slow5_rec_free s5b_rec_free
#define slow5_set_press s5b_set_press
#define slow5_hdr_write s5b_hdr_write
#define slow5_aux_get_int8 s5b_aux_get_int8
#define slow5_aux_get_int16 s5b_aux_get_int16
#define slow5_aux_get_int32 s5b_aux_get_int32
#define slow5_aux_get_int64 s5b_aux_get_int64
#define slow5_aux_get_uint8 s5b_aux_get_uint8
#define slow5_aux_get_uint16 s5b_aux_get_uint16
#define slow5_aux_get_uint32 s5b_aux_get_uint32
#define slow5_aux_get_uint64 s5b_aux_get_uint64
#define slow5_aux_get_float s5b_aux_get_float
#define slow5_aux_get_double s5b_aux_get_double
#define slow5_aux_get_char s5b_aux_get_char
#define slow5_aux_get_enum s5b_aux_get_enum
#define slow5_aux_get_int8_array s5b_aux_get_int8_array
#define slow5_aux_get_int16_array s5b_aux_get_int16_array
#define slow5_aux_get_int32_array s5b_aux_get_int32_array
#define slow5_aux_get_int64_array s5b_aux_get_int64_array
#define slow5_aux_get_uint8_array s5b_aux_get_uint8_array
#define slow5_aux_get_uint16_array s5b_aux_get_uint16_array
#define slow5_aux_get_uint32_array s5b_aux_get_uint32_array
#define slow5_aux_get_uint64_array s5b_aux_get_uint64_array
#define slow5_aux_get_float_array s5b_aux_get_float_array
#define slow5_aux_get_double_array s5b_aux_get_double_array
#define slow5_aux_get_string s5b_aux_get_string
#define slow5_aux_get_enum_array s5b_aux_get_enum_array
#define slow5_hdr_get s5b_hdr_get
#define slow5_get_hdr_keys s5b_get_hdr_keys
#define slow5_get_aux_names s5b_get_aux_names
#define slow5_get_aux_types(h, n) ((enum slow5_aux_type *)s5b_get_aux_types(h, n))
#define slow5_get_aux_enum_labels s5b_get_aux_enum_labels
#define slow5_get_rids s5b_get_rids
#define slow5_rec_init s5b_rec_init
#define slow5_hdr_add s5b_hdr_add
#define slow5_hdr_add_attr s5b_hdr_add
#define slow5_hdr_set s5b_hdr_set
#define slow5_hdr_add_rg s5b_hdr_add_rg
#define slow5_aux_add s5b_aux_add
#define slow5_aux_add_enum s5b_aux_add_enum
#define slow5_aux_set s5b_aux_set
#define slow5_aux_set_string s5b_aux_set_string
#define slow5_hdr_t s5b_hdr_t
#define slow5_errno (s5b_errno_value())
#endif

#ifdef __cplusplus
}
#endif
#endif
