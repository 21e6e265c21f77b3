// test-only stubs of the GPU entry points the host objects reference (none is reached on uncompressed files)
#include <cstdlib>
#include <cstring>
#include "/root/repo/include/slow5b200.h"
extern "C" {
int s5b_ctx_create(int, s5b_ctx_t **o) { *o = nullptr; return S5B_ERR_DEVICE; }
void s5b_ctx_destroy(s5b_ctx_t *) {}
const char *s5b_strerror(int) { return "stub"; }
void *s5b_host_alloc(size_t n) { return malloc(n); }
void s5b_host_free(void *p) { free(p); }
int s5b_depress_batch_host(s5b_ctx_t *, int, const void *const *, const size_t *, size_t, void **, size_t *) { return S5B_ERR_DEVICE; }
int s5b_compress_batch_host(s5b_ctx_t *, int, const void *const *, const size_t *, size_t, void **, size_t *) { return S5B_ERR_DEVICE; }
int s5b_compress_records_host(s5b_ctx_t *, int, const void *const *, const size_t *, const uint32_t *, size_t, void **, size_t *) { return S5B_ERR_DEVICE; }
void *s5b_ptr_compress_solo(int, const void *, size_t, size_t *n) { *n = 0; return nullptr; }
void *s5b_ptr_depress_solo(int, const void *, size_t, size_t *n) { *n = 0; return nullptr; }
int s5b_blow5_read_ids_host(s5b_ctx_t *, int in_rec, const uint8_t *h_in, uint64_t, const uint64_t *rec_off, const uint32_t *rec_len,
                            uint64_t n, uint8_t *h_ids, uint64_t ids_cap, uint64_t *id_off) {
    if (in_rec != 0) return S5B_ERR_DEVICE;   // uncompressed records: the id is right behind its u16 length
    uint64_t at = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint16_t l;
        memcpy(&l, h_in + rec_off[i], 2);
        (void)rec_len;
        if (at + l > ids_cap) return S5B_ERR_NOSPACE;
        id_off[i] = at;
        memcpy(h_ids + at, h_in + rec_off[i] + 2, l);
        at += l;
    }
    id_off[n] = at;
    return S5B_OK;
}
}
extern "C" {
int s5b_blow5_recode_host(s5b_ctx_t *, int, int, int, int, const uint8_t *, uint64_t, const uint64_t *, const uint32_t *, uint64_t, uint8_t *, uint64_t, uint64_t *) { return S5B_ERR_DEVICE; }
int s5b_ascii_to_signal_batch_host(s5b_ctx_t *, const char *const *, const size_t *, const uint64_t *, size_t, int16_t **, size_t *) { return S5B_ERR_DEVICE; }
int s5b_signal_to_ascii_batch_host(s5b_ctx_t *, int, const void *const *, const size_t *, size_t, char **, size_t *) { return S5B_ERR_DEVICE; }
int s5b_qts_round_batch_host(s5b_ctx_t *, int, const void *const *, const size_t *, size_t, void **, size_t *) { return S5B_ERR_DEVICE; }
int s5b_ctx_set_aux_layout(s5b_ctx_t *, const uint8_t *, const uint8_t *, uint32_t) { return 0; }
int s5b_ctx_set_rg_map(s5b_ctx_t *, const uint32_t *, uint32_t) { return 0; }
int s5b_ctx_set_degrade(s5b_ctx_t *, int, int, float, float) { return 0; }
const char *s5b_version(void) { return "asan-stub"; }
}
