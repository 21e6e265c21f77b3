/* oracle/blow5_oracle.c -- TEST / BASELINE INFRASTRUCTURE (see oracle.h).
 *
 * CPU restatement of the per-record transcoding step of `slow5tools view` for binary records:
 *     slow5_rec_depress_parse   slow5lib/src/slow5.c:2580-2611   (record decompression + parse)
 *     slow5_rec_parse (binary)  slow5lib/src/slow5.c:2811-2950   (field walk, signal decompression :2913-2925)
 *     slow5_rec_to_mem (binary) slow5lib/src/slow5.c:3928-4074   (signal compression :3973-3990, record
 *                                                                  compression :4046-4052, u64 size prefix :4055-4060)
 * for records = [u16 id_len][id][u32 read_group][f64 x4][u64 len_raw_signal][signal][aux bytes].
 * Auxiliary fields are carried over byte for byte (binary aux fields are position independent, slow5.c:3993-4044).
 * Signal methods: none / svb-zd / ex-zd through this oracle's own restatements; record methods: none / zlib through
 * the SYSTEM zlib (the same library the reference links: level 6 = Z_DEFAULT_COMPRESSION, wbits 15, memLevel 8,
 * slow5_press.c:816-827), zstd is not restated here (libzstd has no header in this image; the tests use ctypes).
 *
 * Pinned by tests/test_oracle_blow5.py against the compiled reference (oracle/_ref) run on the same records and
 * against records cut from the reference's own BLOW5 fixtures (tests/golden/).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define M_NONE 0
#define M_ZLIB 1
#define M_SVB_ZD 2
#define M_EX_ZD 4

static uint64_t rd_u64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* slow5_press.c:876-913 (ptr_compress_zlib_solo): one complete zlib stream */
static uint8_t *zlib_pack(const uint8_t *in, size_t n, size_t *out_n) {
    uLongf cap = compressBound((uLong) n);
    uint8_t *out = (uint8_t *) malloc(cap ? cap : 1);
    if (!out) return NULL;
    if (compress2(out, &cap, in, (uLong) n, Z_DEFAULT_COMPRESSION) != Z_OK) { free(out); return NULL; }
    *out_n = cap;
    return out;
}

/* slow5_press.c:973-1010 (ptr_depress_zlib_solo): inflate until the stream ends; input that stops early is not an
 * error there (Z_BUF_ERROR falls through), the bytes decoded so far are returned */
static uint8_t *zlib_unpack(const uint8_t *in, size_t n, size_t *out_n) {
    z_stream s;
    memset(&s, 0, sizeof s);
    if (inflateInit2(&s, 15) != Z_OK) return NULL;
    size_t cap = n * 4 + 1024, have = 0;
    uint8_t *out = (uint8_t *) malloc(cap);
    if (!out) { inflateEnd(&s); return NULL; }
    s.next_in = (Bytef *) in;
    s.avail_in = (uInt) n;
    for (;;) {
        s.next_out = out + have;
        s.avail_out = (uInt) (cap - have);
        int rc = inflate(&s, Z_NO_FLUSH);
        have = cap - s.avail_out;
        if (rc == Z_STREAM_END) break;
        if (rc == Z_STREAM_ERROR || rc == Z_DATA_ERROR || rc == Z_NEED_DICT || rc == Z_MEM_ERROR) {
            free(out); inflateEnd(&s); return NULL;
        }
        if (s.avail_out != 0) break;              /* no progress possible: truncated input */
        cap *= 2;
        uint8_t *bigger = (uint8_t *) realloc(out, cap);
        if (!bigger) { free(out); inflateEnd(&s); return NULL; }
        out = bigger;
    }
    inflateEnd(&s);
    *out_n = have;
    return out;
}

/* One stored record (size prefix excluded) in, one output record INCLUDING its u64 size prefix out (malloc'd).
 * Returns 0, -13 (SLOW5_ERR_PRESS), -4 (SLOW5_ERR_RECPARSE, slow5_defs.h:140), -10 (SLOW5_ERR_MEM) or -2 (unsupported method). */
int orc_blow5_recode_record(int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *in, size_t in_len,
                            uint8_t **out, size_t *out_len) {
    *out = NULL;
    *out_len = 0;
    if ((in_rec != M_NONE && in_rec != M_ZLIB) || (out_rec != M_NONE && out_rec != M_ZLIB)) return -2;
    uint8_t *plain = NULL;
    const uint8_t *rec = in;
    size_t len = in_len;
    if (in_rec == M_ZLIB) {                                        /* slow5.c:2583-2597 */
        plain = zlib_unpack(in, in_len, &len);
        if (!plain || len == 0) { free(plain); return -13; }
        rec = plain;
    }
    int rc = 0;
    uint8_t *sig_new = NULL, *packed = NULL, *final = NULL;
    int16_t *samples = NULL;
    /* ---- field walk, slow5.c:2811-2927 */
    if (len < 2) { rc = -4; goto done; }
    uint16_t idlen; memcpy(&idlen, rec, 2);
    size_t head = 2 + (size_t) idlen + 4 + 32;
    if (head + 8 > len) { rc = -4; goto done; }
    uint64_t lrs = rd_u64(rec + head);
    size_t sig_at = head + 8;
    uint64_t sig_bytes = in_sig == M_NONE ? lrs * 2 : lrs;
    if (sig_bytes > len - sig_at) { rc = -4; goto done; }
    const uint8_t *aux = rec + sig_at + sig_bytes;
    size_t aux_len = len - sig_at - (size_t) sig_bytes;
    /* ---- signal: decode (slow5.c:2913-2925) */
    uint64_t ns = 0;
    const uint8_t *sig_out = rec + sig_at;
    uint64_t sig_out_bytes = sig_bytes, lrs_out = lrs;
    if (in_sig != out_sig) {
        const int16_t *raw;
        if (in_sig == M_NONE) {
            ns = lrs;
            samples = (int16_t *) malloc((size_t) ns * 2 + 2);
            if (!samples) { rc = -10; goto done; }
            memcpy(samples, rec + sig_at, (size_t) ns * 2);     /* unaligned in the record */
        } else if (in_sig == M_SVB_ZD) {
            uint32_t n32 = 0;
            if (sig_bytes < 4) { rc = -13; goto done; }
            memcpy(&n32, rec + sig_at, 4);
            samples = (int16_t *) malloc((size_t) n32 * 2 + 2);
            if (!samples) { rc = -10; goto done; }
            if (orc_svbzd_depress(rec + sig_at, (size_t) sig_bytes, samples, n32, &n32) != 0) { rc = -13; goto done; }
            ns = n32;
        } else if (in_sig == M_EX_ZD) {
            if (sig_bytes < 9) { rc = -13; goto done; }
            uint64_t nin = rd_u64(rec + sig_at + 1);
            if (nin > (1ull << 32)) { rc = -13; goto done; }
            samples = (int16_t *) malloc((size_t) nin * 2 + 2);
            if (!samples) { rc = -10; goto done; }
            if (orc_exzd_depress(rec + sig_at, (size_t) sig_bytes, samples, (size_t) nin, &nin) != 0) { rc = -13; goto done; }
            ns = nin;
        } else { rc = -2; goto done; }
        raw = samples;
        /* ---- signal: encode (slow5.c:3973-3990); len_raw_signal becomes the compressed byte count */
        if (out_sig == M_NONE) {
            sig_out = (const uint8_t *) raw;
            sig_out_bytes = ns * 2;
            lrs_out = ns;
        } else if (out_sig == M_SVB_ZD) {
            sig_new = (uint8_t *) malloc(orc_svbzd_bound((uint32_t) ns) + 16);
            if (!sig_new) { rc = -10; goto done; }
            sig_out_bytes = orc_svbzd_compress(raw, (size_t) ns * 2, sig_new);
            sig_out = sig_new;
            lrs_out = sig_out_bytes;
        } else if (out_sig == M_EX_ZD) {
            sig_new = (uint8_t *) malloc(orc_exzd_bound(ns) + 16);
            if (!sig_new) { rc = -10; goto done; }
            sig_out_bytes = orc_exzd_compress(raw, (size_t) ns * 2, sig_new);
            if (sig_out_bytes == 0) { rc = -13; goto done; }
            sig_out = sig_new;
            lrs_out = sig_out_bytes;
        } else { rc = -2; goto done; }
    }
    /* ---- pack (slow5.c:3928-4044) */
    size_t packed_len = head + 8 + (size_t) sig_out_bytes + aux_len;
    packed = (uint8_t *) malloc(packed_len + 8);
    if (!packed) { rc = -10; goto done; }
    memcpy(packed + 8, rec, head);
    memcpy(packed + 8 + head, &lrs_out, 8);
    memcpy(packed + 8 + head + 8, sig_out, (size_t) sig_out_bytes);
    memcpy(packed + 8 + head + 8 + sig_out_bytes, aux, aux_len);
    /* ---- record compression + size prefix (slow5.c:4046-4060) */
    if (out_rec == M_ZLIB) {
        size_t zn = 0;
        uint8_t *z = zlib_pack(packed + 8, packed_len, &zn);
        if (!z) { rc = -13; goto done; }
        final = (uint8_t *) malloc(zn + 8);
        if (!final) { free(z); rc = -10; goto done; }
        uint64_t zn64 = zn;
        memcpy(final, &zn64, 8);
        memcpy(final + 8, z, zn);
        free(z);
        *out = final;
        *out_len = zn + 8;
    } else {
        uint64_t pl = packed_len;
        memcpy(packed, &pl, 8);
        *out = packed;
        *out_len = packed_len + 8;
        packed = NULL;
    }
done:
    free(plain); free(sig_new); free(samples); free(packed);
    return rc;
}
