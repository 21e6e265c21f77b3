/* oracle/exzd_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  Plain-C restatement of the reference's
 * "ex-zd" signal codec (slow5lib/src/slow5_press.c:1236-1848): QTS shift, 16-bit zigzag-delta, one byte per
 * value with an exception list for values above 255.  One pass per stage, no allocation tricks; every function
 * cites the reference lines it follows.  Pinned by tests/test_oracle_exzd.py against the compiled reference
 * (oracle/_ref) and the reference's own ex-zd BLOW5 golden (test/data/exp/one_fast5/exp_1_lossless_zlib_ex_zd.blow5,
 * carried as tests/golden/exzd_ref_vectors.npz).
 *
 * Stream layout (little endian, slow5_press.c:1721-1776, :1596-1628, :1263-1424):
 *   u8  version = 0
 *   u64 nin                      number of samples
 *   u8  q                        QTS: low bits shifted out of every sample (0..5)
 *   u16 zd[0]                    zigzag-delta of the first (shifted) sample, prev = 0
 *   u32 nex                      number of exceptions among zd[1..nin)  (value > 255)
 *   nex > 1 : u32 len, svb(pos[0], pos[i]-pos[i-1]-1 ...) ; u32 len, svb(zd - 256 ...)   (plain StreamVByte)
 *   nex == 1: u32 pos, u32 (zd - 256)
 *   one byte per non-exception value, in order
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* slow5_press.c:1567-1570 (int16 argument: the delta wraps mod 2^16 before the zigzag) */
static uint16_t zz16(int16_t x) { return (uint16_t)((x + x) ^ (x >> 15)); }
/* slow5_press.c:1630-1633 */
static int16_t unzz16(uint16_t x) { return (int16_t)((x >> 1) ^ -(x & 1)); }

/* slow5_press.c:1675-1698 : number of low zero bits shared by all samples, at most 5 (5 for an empty array) */
static uint8_t find_qts(const int16_t *s, uint64_t n) {
    uint8_t q = 5;
    for (uint64_t i = 0; i < n && q; ++i)
        while (q && (s[i] & ((1 << q) - 1))) --q;
    return q;
}

size_t orc_exzd_bound(uint64_t n) {
    /* header 10 + zd0 2 + nex 4 + two (len + svb worst case) + bytes: generous */
    return 32 + 2 * (8 + (size_t)(n + 3) / 4 + 4 * (size_t)n) + (size_t)n;
}

/* ptr_compress_ex_zd_v0, slow5_press.c:1721-1776.  Returns bytes written, 0 on failure (nin == 0 is undefined
 * behaviour in the reference -- zigdelta of an empty array is dereferenced, :1612 -- and is refused here). */
size_t orc_exzd_compress(const int16_t *in, size_t count_bytes, uint8_t *out) {
    const uint64_t nin = count_bytes / 2;
    if (nin == 0) return 0;
    size_t off = 0;
    out[off++] = 0;
    memcpy(out + off, &nin, 8);
    off += 8;
    const uint8_t q = find_qts(in, nin);
    out[off++] = q;
    /* zigdelta_16_u16 on the shifted samples (do_qts :1700-1711 is an arithmetic shift) */
    uint16_t *zd = (uint16_t *)malloc(nin * sizeof *zd);
    int16_t prev = 0;
    for (uint64_t i = 0; i < nin; ++i) {
        const int16_t s = (int16_t)(in[i] >> q);
        zd[i] = zz16((int16_t)(s - prev));
        prev = s;
    }
    memcpy(out + off, zd, 2);
    off += 2;
    /* ex_press(zd + 1, nin - 1), :1263-1424 */
    const uint32_t m = (uint32_t)(nin - 1);
    const uint16_t *v = zd + 1;
    uint32_t nex = 0;
    for (uint32_t i = 0; i < m; ++i) nex += v[i] > 255;
    memcpy(out + off, &nex, 4);
    off += 4;
    if (nex > 0) {
        uint32_t *pos = (uint32_t *)malloc(nex * sizeof *pos), *ex = (uint32_t *)malloc(nex * sizeof *ex);
        uint32_t k = 0;
        for (uint32_t i = 0; i < m; ++i)
            if (v[i] > 255) {
                pos[k] = i;
                ex[k] = (uint32_t)v[i] - 256;
                ++k;
            }
        if (nex > 1) {
            uint32_t *d = (uint32_t *)malloc(nex * sizeof *d);
            d[0] = pos[0];
            for (uint32_t i = 1; i < nex; ++i) d[i] = pos[i] - pos[i - 1] - 1; /* delta_increasing_u32 :1236-1260 */
            uint32_t len = (uint32_t)orc_svb_encode(d, nex, out + off + 4);
            memcpy(out + off, &len, 4);
            off += 4 + len;
            len = (uint32_t)orc_svb_encode(ex, nex, out + off + 4);
            memcpy(out + off, &len, 4);
            off += 4 + len;
            free(d);
        } else {
            memcpy(out + off, pos, 4);
            memcpy(out + off + 4, ex, 4);
            off += 8;
        }
        free(pos);
        free(ex);
    }
    for (uint32_t i = 0; i < m; ++i)
        if (v[i] <= 255) out[off++] = (uint8_t)v[i];
    free(zd);
    return off;
}

/* ptr_depress_ex_zd, slow5_press.c:1787-1848 + ex_zd_depress_16 :1646-1673 + ex_depress :1441-1561.
 * Returns 0 and *n_samples, or -13 (SLOW5_ERR_PRESS) for an unsupported version / an svb section that does not
 * consume its stated length, -2 when the output is too small or the header is cut short.  (The reference trusts the
 * rest of the stream; inputs that would make it read or write out of bounds are refused here with -13.) */
int orc_exzd_depress(const uint8_t *in, size_t count, int16_t *out, size_t out_cap_samples, uint64_t *n_samples) {
    if (count < 16) return -2;
    if (in[0] != 0) return -13;
    uint64_t nin;
    memcpy(&nin, in + 1, 8);
    const uint8_t q = in[9];
    *n_samples = nin;
    if (nin == 0 || nin > out_cap_samples) return -2;
    if (q > 5) return -13;
    uint16_t *zd = (uint16_t *)calloc(nin, sizeof *zd);
    size_t off = 10;
    memcpy(zd, in + off, 2);
    off += 2;
    uint32_t nex;
    memcpy(&nex, in + off, 4);
    off += 4;
    const uint64_t m = nin - 1;
    int rc = 0;
    uint32_t *pos = NULL, *ex = NULL;
    if (nex > m) rc = -13;
    if (!rc && nex > 0) {
        pos = (uint32_t *)malloc((size_t)nex * sizeof *pos);
        ex = (uint32_t *)malloc((size_t)nex * sizeof *ex);
        if (nex > 1) {
            uint32_t len;
            if (off + 4 > count) rc = -13;
            if (!rc) {
                memcpy(&len, in + off, 4);
                off += 4;
                if (off + len > count || len < (nex + 3) / 4 || orc_svb_decode(in + off, pos, nex) != len) rc = -13;
                off += len;
            }
            if (!rc) {
                for (uint32_t i = 1; i < nex; ++i) pos[i] += pos[i - 1] + 1; /* undelta :1427-1438 */
                if (off + 4 > count) rc = -13;
            }
            if (!rc) {
                memcpy(&len, in + off, 4);
                off += 4;
                if (off + len > count || len < (nex + 3) / 4 || orc_svb_decode(in + off, ex, nex) != len) rc = -13;
                off += len;
            }
        } else {
            if (off + 8 > count) rc = -13;
            if (!rc) {
                memcpy(pos, in + off, 4);
                memcpy(ex, in + off + 4, 4);
                off += 8;
            }
        }
        if (!rc)
            for (uint32_t i = 0; i < nex; ++i) {
                if (pos[i] >= m || (i && pos[i] <= pos[i - 1])) {
                    rc = -13;
                    break;
                }
                zd[1 + pos[i]] = (uint16_t)(ex[i] + 256);
            }
    }
    if (!rc && count - off != m - nex) rc = -13; /* one byte per remaining value, nothing else */
    if (!rc) {
        uint32_t j = 0;
        for (uint64_t i = 0; i < m; ++i) {
            if (j < nex && i == pos[j]) ++j;
            else zd[1 + i] = in[off++];
        }
        int16_t prev = 0; /* unzigdelta_u16_16 :1635-1645, then do_rev_qts_inplace :1713-1718 */
        for (uint64_t i = 0; i < nin; ++i) {
            prev = (int16_t)(prev + unzz16(zd[i]));
            out[i] = (int16_t)((uint16_t)prev << q);
        }
    }
    free(pos);
    free(ex);
    free(zd);
    return rc;
}
