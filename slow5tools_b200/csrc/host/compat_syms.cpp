// compat_syms.cpp -> libslow5b200_compat.so: the slow5lib symbol names themselves (slow5_open, slow5_get_next_bytes,
// slow5_decode, slow5_encode, slow5_ptr_compress_solo, slow5_get_next_batch ...), each a direct forward to its s5b_ twin in
// libslow5b200.so.  For programs that were compiled against slow5lib's own headers and only need relinking
// (slow5lib/include/slow5/slow5.h:345-662, slow5_press.h:97-125, slow5_mt.h:49-65); code that can be recompiled uses the
// macro mapping of include/compat/slow5/slow5.h instead.  Kept out of libslow5b200.so so that a process may hold this
// library and the reference's libslow5 side by side (the parity tests do).
#include "../../../include/slow5b200_file.h"

extern "C" {

// slow5_errno is `(*slow5_errno_location())` in the reference (slow5_error.h:120-126)
int *slow5_errno_location(void) {
    static thread_local int value;
    value = s5b_errno_value();
    return &value;
}

s5b_file_t *slow5_open(const char *pathname, const char *mode) { return s5b_open(pathname, mode); }
int slow5_close(s5b_file_t *fp) { return s5b_close(fp); }
void *slow5_get_next_mem(size_t *n, const s5b_file_t *fp) { return s5b_get_next_mem(n, const_cast<s5b_file_t *>(fp)); }
int slow5_get_next_bytes(char **mem, size_t *bytes, s5b_file_t *fp) { return s5b_get_next_bytes(mem, bytes, fp); }
int slow5_decode(char **mem, size_t *bytes, s5b_rec_t **read, s5b_file_t *fp) { return s5b_decode(mem, bytes, read, fp); }
int slow5_encode(char **mem, size_t *bytes, s5b_rec_t *read, s5b_file_t *fp) { return s5b_encode(mem, bytes, read, fp); }
int slow5_write_bytes(char *mem, size_t bytes, s5b_file_t *fp) { return s5b_write_bytes(mem, bytes, fp); }
void slow5_rec_free(s5b_rec_t *read) { s5b_rec_free(read); }
int slow5_get_next(s5b_rec_t **read, s5b_file_t *fp) { return s5b_get_next(read, fp); }
int slow5_get(const char *read_id, s5b_rec_t **read, s5b_file_t *fp) { return s5b_get(read_id, read, fp); }
int slow5_write(s5b_rec_t *read, s5b_file_t *fp) { return s5b_write(read, fp); }
int slow5_set_press(s5b_file_t *fp, int rec_press, int sig_press) { return s5b_set_press(fp, rec_press, sig_press); }
int slow5_hdr_write(s5b_file_t *fp) { return s5b_hdr_write(fp); }

s5b_mt_t *slow5_init_mt(int num_thread, s5b_file_t *fp) { return s5b_init_mt(num_thread, fp); }
s5b_batch_t *slow5_init_batch(int cap) { return s5b_init_batch(cap); }
int slow5_get_next_batch(s5b_mt_t *mt, s5b_batch_t *b, int n) { return s5b_get_next_batch(mt, b, n); }
int slow5_encode_batch(s5b_mt_t *mt, s5b_batch_t *b, int n) { return s5b_encode_batch_mt(mt, b, n); }
int slow5_write_batch(s5b_mt_t *mt, s5b_batch_t *b, int n) { return s5b_write_batch(mt, b, n); }
int slow5_get_batch(s5b_mt_t *mt, s5b_batch_t *b, char **rid, int n) { return s5b_get_batch(mt, b, rid, n); }
int slow5_idx_load(s5b_file_t *fp) { return s5b_idx_load(fp); }
void slow5_idx_unload(s5b_file_t *fp) { s5b_idx_unload(fp); }
void slow5_free_batch(s5b_batch_t *b) { s5b_free_batch(b); }
void slow5_free_mt(s5b_mt_t *mt) { s5b_free_mt(mt); }
int slow5_get_next_batch_lazy(s5b_rec_t ***read, s5b_file_t *fp, int n, int t) { return s5b_get_next_batch_lazy(read, fp, n, t); }
int slow5_get_batch_lazy(s5b_rec_t ***read, s5b_file_t *fp, char **rid, int n, int t) { return s5b_get_batch_lazy(read, fp, rid, n, t); }
int slow5_write_batch_lazy(s5b_rec_t **read, s5b_file_t *fp, int n, int t) { return s5b_write_batch_lazy(read, fp, n, t); }
void slow5_free_batch_lazy(s5b_rec_t ***read, int n) { s5b_free_batch_lazy(read, n); }
int slow5_hdr_add_attr(const char *attr, s5b_hdr_t *h) { return s5b_hdr_add(attr, h); }

s5b_press_t *slow5_press_init(s5b_press_method_t m) { return s5b_press_init(m); }
struct __s5b_press *__slow5_press_init(int method) { return __s5b_press_init(method); }
void slow5_press_free(s5b_press_t *c) { s5b_press_free(c); }
void __slow5_press_free(struct __s5b_press *c) { __s5b_press_free(c); }
void *slow5_ptr_compress(struct __s5b_press *c, const void *p, size_t count, size_t *n) { return s5b_ptr_compress(c, p, count, n); }
void *slow5_ptr_depress(struct __s5b_press *c, const void *p, size_t count, size_t *n) { return s5b_ptr_depress(c, p, count, n); }
void *slow5_ptr_compress_solo(int method, const void *p, size_t count, size_t *n) { return s5b_ptr_compress_solo(method, p, count, n); }
void *slow5_ptr_depress_solo(int method, const void *p, size_t count, size_t *n) { return s5b_ptr_depress_solo(method, p, count, n); }
void slow5_compress_footer_next(struct __s5b_press *c) { s5b_compress_footer_next(c); }


// auxiliary field accessors and header attributes (slow5.h:396, :469-508)
int8_t slow5_aux_get_int8(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_int8(r, f, err); }
int16_t slow5_aux_get_int16(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_int16(r, f, err); }
int32_t slow5_aux_get_int32(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_int32(r, f, err); }
int64_t slow5_aux_get_int64(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_int64(r, f, err); }
uint8_t slow5_aux_get_uint8(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_uint8(r, f, err); }
uint16_t slow5_aux_get_uint16(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_uint16(r, f, err); }
uint32_t slow5_aux_get_uint32(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_uint32(r, f, err); }
uint64_t slow5_aux_get_uint64(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_uint64(r, f, err); }
float slow5_aux_get_float(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_float(r, f, err); }
double slow5_aux_get_double(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_double(r, f, err); }
char slow5_aux_get_char(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_char(r, f, err); }
uint8_t slow5_aux_get_enum(const s5b_rec_t *r, const char *f, int *err) { return s5b_aux_get_enum(r, f, err); }
int8_t *slow5_aux_get_int8_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_int8_array(r, f, len, err); }
int16_t *slow5_aux_get_int16_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_int16_array(r, f, len, err); }
int32_t *slow5_aux_get_int32_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_int32_array(r, f, len, err); }
int64_t *slow5_aux_get_int64_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_int64_array(r, f, len, err); }
uint8_t *slow5_aux_get_uint8_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_uint8_array(r, f, len, err); }
uint16_t *slow5_aux_get_uint16_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_uint16_array(r, f, len, err); }
uint32_t *slow5_aux_get_uint32_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_uint32_array(r, f, len, err); }
uint64_t *slow5_aux_get_uint64_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_uint64_array(r, f, len, err); }
float *slow5_aux_get_float_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_float_array(r, f, len, err); }
double *slow5_aux_get_double_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_double_array(r, f, len, err); }
char *slow5_aux_get_string(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_string(r, f, len, err); }
uint8_t *slow5_aux_get_enum_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return s5b_aux_get_enum_array(r, f, len, err); }
char *slow5_hdr_get(const char *attr, uint32_t rg, const s5b_hdr_t *h) { return s5b_hdr_get(attr, rg, h); }
const char **slow5_get_hdr_keys(const s5b_hdr_t *h, uint64_t *len) { return s5b_get_hdr_keys(h, len); }
char **slow5_get_aux_names(const s5b_hdr_t *h, uint64_t *len) { return s5b_get_aux_names(h, len); }
int *slow5_get_aux_types(const s5b_hdr_t *h, uint64_t *len) { return s5b_get_aux_types(h, len); }
char **slow5_get_aux_enum_labels(const s5b_hdr_t *h, const char *field, uint8_t *n) { return s5b_get_aux_enum_labels(h, field, n); }
char **slow5_get_rids(const s5b_file_t *fp, uint64_t *len) { return s5b_get_rids(fp, len); }
s5b_rec_t *slow5_rec_init(void) { return s5b_rec_init(); }
int slow5_hdr_add(const char *attr, s5b_hdr_t *h) { return s5b_hdr_add(attr, h); }
int slow5_hdr_set(const char *attr, const char *value, uint32_t rg, s5b_hdr_t *h) { return s5b_hdr_set(attr, value, rg, h); }
int64_t slow5_hdr_add_rg(s5b_hdr_t *h) { return s5b_hdr_add_rg(h); }
int slow5_aux_add(const char *field, int type, s5b_hdr_t *h) { return s5b_aux_add(field, type, h); }
int slow5_aux_add_enum(const char *field, const char **labels, uint8_t n, s5b_hdr_t *h) { return s5b_aux_add_enum(field, labels, n, h); }
int slow5_aux_set(s5b_rec_t *r, const char *field, const void *data, s5b_hdr_t *h) { return s5b_aux_set(r, field, data, h); }
int slow5_aux_set_string(s5b_rec_t *r, const char *field, const char *data, s5b_hdr_t *h) { return s5b_aux_set_string(r, field, data, h); }

}  // extern "C"
