// zstd_core.h -- Zstandard frame decoder (RFC 8878) written for one CUDA thread per serial dependency chain.
//
// Replaces, together with zstd_kernels.cu, ptr_depress_zstd (slow5lib/src/slow5_press.c:1205-1230:
// ZSTD_getFrameContentSize + ZSTD_decompress from system libzstd >= 1.3).  The algorithm is libzstd's (an
// unvendored dependency of the reference); this is a restatement of the published format: frame header, raw /
// RLE / compressed blocks, Huffman-coded literals (1 or 4 streams, direct or FSE-compressed weights), FSE-coded
// sequences (predefined / RLE / compressed / repeat modes), repeat offsets, optional XXH64 content checksum.
//
// The functions are plain serial code over byte pointers, marked host+device so the same source is exercised on
// the CPU by tests/zstd_host_check.cpp during development; the product only ever runs them inside the kernel.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define ZHD __host__ __device__ __forceinline__
#define ZHDN __host__ __device__
#else
#define ZHD inline
#define ZHDN
#endif

// development aid: the host harness records where a frame was rejected
#if defined(S5BZ_TRACE) && !defined(__CUDA_ARCH__)
static int s5bz_fail_line = 0;
#define ZF(code) (s5bz_fail_line = __LINE__, (code))
#else
#define ZF(code) (code)
#endif

namespace s5bz {

enum { Z_OK = 0, Z_ERR_CORRUPT = -13 /* S5B_ERR_PRESS */, Z_ERR_NOSPACE = -40, Z_ERR_UNSUPPORTED = -13 };

struct FseEntry {
    uint16_t base;
    uint8_t sym;
    uint8_t nbits;
};

constexpr int HUF_MAX_BITS = 11;
constexpr int LL_MAX_AL = 9, OF_MAX_AL = 8, ML_MAX_AL = 9, WT_MAX_AL = 6;

// per-stream decoding tables (shared memory in the kernel, stack/heap on the host)
struct Tables {
    uint16_t huf[1 << HUF_MAX_BITS];  // (symbol << 4) | nbits
    FseEntry ll[1 << LL_MAX_AL];
    FseEntry of[1 << OF_MAX_AL];
    FseEntry ml[1 << ML_MAX_AL];
    FseEntry wt[1 << WT_MAX_AL];
    uint8_t weights[256];
    int16_t norm[64];      // normalised counts while a table is being built
    uint8_t spread[512];   // symbol spreading scratch
    uint16_t next[64];     // per-symbol next-state scratch
    int huf_bits;          // 0 = no Huffman table yet
    int ll_al, of_al, ml_al;  // accuracy logs of the current sequence tables (-1 = none yet)
};

ZHD int highest_bit(uint32_t v) {
#ifdef __CUDA_ARCH__
    return 31 - __clz(v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// ---- forward bit reader (FSE table descriptions) ---------------------------------------------------------
struct FwdBits {
    const uint8_t *p;
    uint64_t nbits_total;
    uint64_t pos;  // bit position
    ZHD uint32_t read(int n) {
        uint32_t v = 0;
        for (int i = 0; i < n; ++i) {
            const uint64_t b = pos + i;
            if (b < nbits_total) v |= (uint32_t)((p[b >> 3] >> (b & 7)) & 1u) << i;
        }
        pos += n;
        return v;
    }
};

// ---- backward bit reader (Huffman streams, FSE streams) --------------------------------------------------
// `off` is the number of unread bits (counted from the start of the stream); reads take the n bits just below
// it.  Bits below the start of the stream read as zero and drive `off` negative (the callers check for that).
struct BackBits {
    const uint8_t *p;
    int64_t off;
    ZHD bool init(const uint8_t *src, uint32_t len) {
        p = src;
        if (len == 0 || src[len - 1] == 0) return false;
        off = (int64_t)len * 8 - (8 - highest_bit(src[len - 1]));
        return true;
    }
    ZHD uint32_t read(int n) {
        off -= n;
        if (n == 0) return 0;
        int64_t lo = off;
        int shift = 0;
        int take = n;
        if (lo < 0) {  // part (or all) of the request lies before the stream start: those bits are zero
            shift = (int)(-lo < n ? -lo : n);
            take = n - shift;
            lo = 0;
        }
        if (take <= 0) return 0;
        uint64_t w = 0;
        const int64_t byte0 = lo >> 3;
        const int nb = (int)(((lo & 7) + take + 7) >> 3);
        for (int i = 0; i < nb; ++i) w |= (uint64_t)p[byte0 + i] << (8 * i);
        const uint32_t v = (uint32_t)((w >> (lo & 7)) & ((1ull << take) - 1ull));
        return v << shift;
    }
};

// ---- FSE ----------------------------------------------------------------------------------------------------
// Reads a normalised-count table description (forward bits) into t.norm; returns bytes consumed or < 0.
ZHDN inline int fse_read_ncount(Tables &t, const uint8_t *src, uint32_t len, int max_al, int max_sym, int *al_out, int *nsym_out) {
    FwdBits fb{src, (uint64_t)len * 8, 0};
    if (len < 1) return ZF(Z_ERR_CORRUPT);
    const int al = 5 + (int)fb.read(4);
    if (al > max_al) return ZF(Z_ERR_CORRUPT);
    int remaining = 1 << al;
    int sym = 0;
    while (remaining > 0 && sym <= max_sym) {
        const int bits = highest_bit((uint32_t)(remaining + 1)) + 1;
        uint32_t val = fb.read(bits);
        const uint32_t lower_mask = (1u << (bits - 1)) - 1u;
        const uint32_t threshold = (1u << bits) - 1u - (uint32_t)(remaining + 1);
        if ((val & lower_mask) < threshold) {
            fb.pos -= 1;
            val &= lower_mask;
        } else if (val > lower_mask) {
            val -= threshold;
        }
        const int proba = (int)val - 1;  // -1 = "less than one"
        remaining -= proba < 0 ? -proba : proba;
        t.norm[sym++] = (int16_t)proba;
        if (proba == 0) {
            uint32_t rep = fb.read(2);
            for (;;) {
                for (uint32_t i = 0; i < rep && sym <= max_sym; ++i) t.norm[sym++] = 0;
                if (rep == 3) rep = fb.read(2);
                else break;
            }
        }
        if (fb.pos > fb.nbits_total + 16) return ZF(Z_ERR_CORRUPT);
    }
    if (remaining != 0 || sym > max_sym + 1) return ZF(Z_ERR_CORRUPT);
    if (fb.pos > fb.nbits_total) return ZF(Z_ERR_CORRUPT);
    *al_out = al;
    *nsym_out = sym;
    return (int)((fb.pos + 7) >> 3);
}

// Builds the decoding table from t.norm[0..nsym) (RFC 8878 4.1.1)
ZHDN inline int fse_build(Tables &t, FseEntry *tab, int al, int nsym) {
    const int size = 1 << al;
    int high = size;
    for (int s = 0; s < nsym; ++s) {
        if (t.norm[s] == -1) {
            tab[--high].sym = (uint8_t)s;
            t.next[s] = 1;
        }
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < nsym; ++s) {
        if (t.norm[s] <= 0) continue;
        t.next[s] = (uint16_t)t.norm[s];
        for (int i = 0; i < t.norm[s]; ++i) {
            tab[pos].sym = (uint8_t)s;
            do {
                pos = (pos + step) & mask;
            } while (pos >= high);
        }
    }
    if (pos != 0) return ZF(Z_ERR_CORRUPT);
    for (int i = 0; i < size; ++i) {
        const int s = tab[i].sym;
        const uint32_t nx = t.next[s]++;
        const int nb = al - highest_bit(nx);
        tab[i].nbits = (uint8_t)nb;
        tab[i].base = (uint16_t)((nx << nb) - size);
    }
    return Z_OK;
}

ZHD void fse_rle(FseEntry *tab, uint8_t sym) {
    tab[0].sym = sym;
    tab[0].nbits = 0;
    tab[0].base = 0;
}

// ---- Huffman -----------------------------------------------------------------------------------------------
// weights[0..n) given (the last symbol's weight is implied); builds t.huf
ZHDN inline int huf_build(Tables &t, int n) {
    uint32_t sum = 0;
    for (int i = 0; i < n; ++i) {
        if (t.weights[i] > HUF_MAX_BITS) return ZF(Z_ERR_CORRUPT);
        sum += t.weights[i] ? 1u << (t.weights[i] - 1) : 0;
    }
    if (sum == 0) return ZF(Z_ERR_CORRUPT);
    const int max_bits = highest_bit(sum) + 1;
    if (max_bits > HUF_MAX_BITS) return ZF(Z_ERR_CORRUPT);
    const uint32_t left = (1u << max_bits) - sum;
    if (left & (left - 1)) return ZF(Z_ERR_CORRUPT);  // must be a power of two
    if (n >= 256) return ZF(Z_ERR_CORRUPT);
    t.weights[n] = (uint8_t)(highest_bit(left) + 1);
    const int nsym = n + 1;
    // code length = max_bits + 1 - weight; table filled from the longest codes (smallest weights) upwards
    uint32_t rank_count[HUF_MAX_BITS + 2];
    for (int b = 0; b <= HUF_MAX_BITS + 1; ++b) rank_count[b] = 0;
    for (int i = 0; i < nsym; ++i) rank_count[t.weights[i]]++;
    // position of the first entry of each weight class: weight w occupies (1 << (w-1)) entries per symbol
    uint32_t start[HUF_MAX_BITS + 2];
    uint32_t pos = 0;
    for (int w = 1; w <= max_bits; ++w) {
        start[w] = pos;
        pos += rank_count[w] << (w - 1);
    }
    if (pos != (1u << max_bits)) return ZF(Z_ERR_CORRUPT);
    for (int i = 0; i < nsym; ++i) {
        const int w = t.weights[i];
        if (!w) continue;
        const uint32_t len = 1u << (w - 1);
        const uint16_t e = (uint16_t)((i << 4) | (max_bits + 1 - w));
        for (uint32_t k = 0; k < len; ++k) t.huf[start[w] + k] = e;
        start[w] += len;
    }
    t.huf_bits = max_bits;
    return Z_OK;
}

// Huffman tree description: direct 4-bit weights or FSE-compressed weights.  Returns bytes consumed or < 0.
ZHDN inline int huf_read_tree(Tables &t, const uint8_t *src, uint32_t len) {
    if (len < 1) return ZF(Z_ERR_CORRUPT);
    const uint32_t hb = src[0];
    int n = 0;
    uint32_t used;
    if (hb >= 128) {
        n = (int)hb - 127;
        const uint32_t bytes = (uint32_t)(n + 1) / 2;
        if (1 + bytes > len) return ZF(Z_ERR_CORRUPT);
        for (int i = 0; i < n; ++i) {
            const uint8_t b = src[1 + i / 2];
            t.weights[i] = (i & 1) ? (b & 15) : (b >> 4);
        }
        used = 1 + bytes;
    } else {
        if (1 + hb > len || hb == 0) return ZF(Z_ERR_CORRUPT);
        int al, nsym;
        const int hdr = fse_read_ncount(t, src + 1, hb, WT_MAX_AL, 12, &al, &nsym);
        if (hdr < 0) return hdr;
        if (fse_build(t, t.wt, al, nsym) != Z_OK) return ZF(Z_ERR_CORRUPT);
        BackBits bb;
        if ((uint32_t)hdr >= hb || !bb.init(src + 1 + hdr, hb - hdr)) return ZF(Z_ERR_CORRUPT);
        uint32_t s1 = bb.read(al), s2 = bb.read(al);
        if (bb.off < 0) return ZF(Z_ERR_CORRUPT);
        for (;;) {
            if (n >= 254) return ZF(Z_ERR_CORRUPT);
            t.weights[n++] = t.wt[s1].sym;
            s1 = t.wt[s1].base + bb.read(t.wt[s1].nbits);
            if (bb.off < 0) {
                t.weights[n++] = t.wt[s2].sym;
                break;
            }
            if (n >= 254) return ZF(Z_ERR_CORRUPT);
            t.weights[n++] = t.wt[s2].sym;
            s2 = t.wt[s2].base + bb.read(t.wt[s2].nbits);
            if (bb.off < 0) {
                t.weights[n++] = t.wt[s1].sym;
                break;
            }
        }
        used = 1 + hb;
    }
    const int rc = huf_build(t, n);
    if (rc != Z_OK) return rc;
    return (int)used;
}

// one Huffman stream -> exactly `regen` symbols
ZHDN inline int huf_decode_stream(const Tables &t, const uint8_t *src, uint32_t len, uint8_t *dst, uint32_t regen) {
    BackBits bb;
    if (!bb.init(src, len)) return ZF(Z_ERR_CORRUPT);
    const int mb = t.huf_bits;
    const uint32_t mask = (1u << mb) - 1u;
    uint32_t state = bb.read(mb);
    uint32_t n = 0;
    while (bb.off > -(int64_t)mb) {
        if (n >= regen) return ZF(Z_ERR_CORRUPT);
        const uint16_t e = t.huf[state];
        dst[n++] = (uint8_t)(e >> 4);
        const int nb = e & 15;
        state = ((state << nb) + bb.read(nb)) & mask;
    }
    if (bb.off != -(int64_t)mb || n != regen) return ZF(Z_ERR_CORRUPT);
    return Z_OK;
}

// ---- sequences ------------------------------------------------------------------------------------------------
struct SeqCodes {
    uint32_t ll_base[36];
    uint8_t ll_bits[36];
    uint32_t ml_base[53];
    uint8_t ml_bits[53];
};
ZHD uint32_t ll_base_of(int c) {
    if (c < 16) return (uint32_t)c;
    const uint32_t tab[20] = {16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
    return tab[c - 16];
}
ZHD int ll_bits_of(int c) {
    if (c < 16) return 0;
    const uint8_t tab[20] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    return tab[c - 16];
}
ZHD uint32_t ml_base_of(int c) {
    if (c < 32) return (uint32_t)c + 3;
    const uint32_t tab[21] = {35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
    return tab[c - 32];
}
ZHD int ml_bits_of(int c) {
    if (c < 32) return 0;
    const uint8_t tab[21] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    return tab[c - 32];
}

ZHDN inline void load_default_norm(Tables &t, int which, int *al, int *nsym) {
    const int8_t ll[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
    const int8_t ml[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                           1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    const int8_t of[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    if (which == 0) {
        for (int i = 0; i < 36; ++i) t.norm[i] = ll[i];
        *al = 6;
        *nsym = 36;
    } else if (which == 1) {
        for (int i = 0; i < 29; ++i) t.norm[i] = of[i];
        *al = 5;
        *nsym = 29;
    } else {
        for (int i = 0; i < 53; ++i) t.norm[i] = ml[i];
        *al = 6;
        *nsym = 53;
    }
}

// Sets up one of the three sequence tables according to its mode.  which: 0 LL, 1 OF, 2 ML.
// Returns bytes consumed from src or < 0.
ZHDN inline int seq_table(Tables &t, int which, int mode, const uint8_t *src, uint32_t len) {
    FseEntry *tab = which == 0 ? t.ll : which == 1 ? t.of : t.ml;
    int *alp = which == 0 ? &t.ll_al : which == 1 ? &t.of_al : &t.ml_al;
    const int max_al = which == 0 ? LL_MAX_AL : which == 1 ? OF_MAX_AL : ML_MAX_AL;
    const int max_sym = which == 0 ? 35 : which == 1 ? 31 : 52;
    int al, nsym;
    switch (mode) {
        case 0:
            load_default_norm(t, which, &al, &nsym);
            if (fse_build(t, tab, al, nsym) != Z_OK) return ZF(Z_ERR_CORRUPT);
            *alp = al;
            return 0;
        case 1:
            if (len < 1 || src[0] > max_sym) return ZF(Z_ERR_CORRUPT);
            fse_rle(tab, src[0]);
            *alp = 0;
            return 1;
        case 2: {
            const int used = fse_read_ncount(t, src, len, max_al, max_sym, &al, &nsym);
            if (used < 0) return used;
            if (fse_build(t, tab, al, nsym) != Z_OK) return ZF(Z_ERR_CORRUPT);
            *alp = al;
            return used;
        }
        default:  // repeat
            if (*alp < 0) return ZF(Z_ERR_CORRUPT);
            return 0;
    }
}

// ---- frame ---------------------------------------------------------------------------------------------------
struct FrameInfo {
    uint64_t content_size;
    bool has_content_size;
    bool checksum;
    uint32_t header_bytes;
    uint64_t window;
};

ZHDN inline int parse_frame_header(const uint8_t *src, uint64_t len, FrameInfo &fi) {
    if (len < 6) return ZF(Z_ERR_CORRUPT);
    if (!(src[0] == 0x28 && src[1] == 0xB5 && src[2] == 0x2F && src[3] == 0xFD)) return ZF(Z_ERR_CORRUPT);
    const uint32_t fhd = src[4];
    const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, dict_flag = fhd & 3;
    if (fhd & 0x08) return ZF(Z_ERR_CORRUPT);  // reserved bit
    fi.checksum = (fhd >> 2) & 1;
    uint32_t pos = 5;
    fi.window = 0;
    if (!single) {
        if (pos >= len) return ZF(Z_ERR_CORRUPT);
        const uint32_t wd = src[pos++];
        const uint32_t e = wd >> 3, m = wd & 7;
        const uint64_t base = 1ull << (10 + e);
        fi.window = base + (base >> 3) * m;
    }
    const uint32_t dict_bytes = dict_flag == 3 ? 4 : dict_flag;
    if (dict_bytes) {
        // a dictionary id of 0 means "none"; anything else cannot be honoured
        uint32_t id = 0;
        if (pos + dict_bytes > len) return ZF(Z_ERR_CORRUPT);
        for (uint32_t i = 0; i < dict_bytes; ++i) id |= (uint32_t)src[pos + i] << (8 * i);
        pos += dict_bytes;
        if (id != 0) return Z_ERR_UNSUPPORTED;
    }
    uint32_t fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    fi.has_content_size = fcs_bytes != 0;
    fi.content_size = 0;
    if (pos + fcs_bytes > len) return ZF(Z_ERR_CORRUPT);
    for (uint32_t i = 0; i < fcs_bytes; ++i) fi.content_size |= (uint64_t)src[pos + i] << (8 * i);
    if (fcs_bytes == 2) fi.content_size += 256;
    pos += fcs_bytes;
    if (single) fi.window = fi.content_size;
    fi.header_bytes = pos;
    return Z_OK;
}

// XXH64 (content checksum, low 32 bits stored) ---------------------------------------------------------------
ZHD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
ZHD uint64_t rd64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}
ZHD uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
ZHDN inline uint64_t xxh64(const uint8_t *p, uint64_t len) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P3 = 1609587929392839161ULL,
                   P4 = 9650029242287828579ULL, P5 = 2870177450012600261ULL;
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        do {
            v1 = rotl64(v1 + rd64(p) * P2, 31) * P1;
            v2 = rotl64(v2 + rd64(p + 8) * P2, 31) * P1;
            v3 = rotl64(v3 + rd64(p + 16) * P2, 31) * P1;
            v4 = rotl64(v4 + rd64(p + 24) * P2, 31) * P1;
            p += 32;
        } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = (h ^ (rotl64(v1 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (rotl64(v2 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (rotl64(v3 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (rotl64(v4 * P2, 31) * P1)) * P1 + P4;
    } else {
        h = P5;
    }
    h += len;
    while (p + 8 <= end) {
        h ^= rotl64(rd64(p) * P2, 31) * P1;
        h = rotl64(h, 27) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)rd32(p) * P1;
        h = rotl64(h, 23) * P2 + P3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * P5;
        h = rotl64(h, 11) * P1;
        ++p;
    }
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
}

// Decodes one compressed block's literals section.  lit: scratch for regenerated literals (>= 128 KiB).
// Returns bytes consumed or < 0; *lit_ptr / *lit_len describe the literals (raw literals point into src).
ZHDN inline int decode_literals(Tables &t, const uint8_t *src, uint32_t len, uint8_t *lit, uint32_t lit_cap, const uint8_t **lit_ptr,
                                uint32_t *lit_len, int lane_count = 1) {
    (void)lane_count;
    if (len < 1) return ZF(Z_ERR_CORRUPT);
    const uint32_t b0 = src[0];
    const uint32_t type = b0 & 3, sf = (b0 >> 2) & 3;
    uint32_t regen, comp = 0, hdr, streams = 1;
    if (type < 2) {
        if ((sf & 1) == 0) {
            regen = b0 >> 3;
            hdr = 1;
        } else if (sf == 1) {
            if (len < 2) return ZF(Z_ERR_CORRUPT);
            regen = (b0 >> 4) | ((uint32_t)src[1] << 4);
            hdr = 2;
        } else {
            if (len < 3) return ZF(Z_ERR_CORRUPT);
            regen = (b0 >> 4) | ((uint32_t)src[1] << 4) | ((uint32_t)src[2] << 12);
            hdr = 3;
        }
        if (type == 0) {
            if (hdr + regen > len) return ZF(Z_ERR_CORRUPT);
            *lit_ptr = src + hdr;
            *lit_len = regen;
            return (int)(hdr + regen);
        }
        if (hdr + 1 > len || regen > lit_cap) return ZF(Z_ERR_CORRUPT);
        memset(lit, src[hdr], regen);
        *lit_ptr = lit;
        *lit_len = regen;
        return (int)(hdr + 1);
    }
    if (sf == 0 || sf == 1) {
        if (len < 3) return ZF(Z_ERR_CORRUPT);
        const uint32_t v = b0 | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16);
        regen = (v >> 4) & 0x3FF;
        comp = (v >> 14) & 0x3FF;
        hdr = 3;
        streams = sf == 0 ? 1 : 4;
    } else if (sf == 2) {
        if (len < 4) return ZF(Z_ERR_CORRUPT);
        const uint32_t v = rd32(src);
        regen = (v >> 4) & 0x3FFF;
        comp = v >> 18;
        hdr = 4;
        streams = 4;
    } else {
        if (len < 5) return ZF(Z_ERR_CORRUPT);
        const uint64_t v = (uint64_t)rd32(src) | ((uint64_t)src[4] << 32);
        regen = (uint32_t)(v >> 4) & 0x3FFFF;
        comp = (uint32_t)(v >> 22) & 0x3FFFF;
        hdr = 5;
        streams = 4;
    }
    if (hdr + comp > len || regen > lit_cap) return ZF(Z_ERR_CORRUPT);
    const uint8_t *p = src + hdr;
    uint32_t left = comp;
    if (type == 2) {
        const int used = huf_read_tree(t, p, left);
        if (used < 0) return used;
        p += used;
        left -= used;
    } else if (t.huf_bits == 0) {
        return ZF(Z_ERR_CORRUPT);  // treeless block without a previous tree
    }
    if (streams == 1) {
        const int rc = huf_decode_stream(t, p, left, lit, regen);
        if (rc != Z_OK) return rc;
    } else {
        if (left < 6) return ZF(Z_ERR_CORRUPT);
        const uint32_t s1 = p[0] | (p[1] << 8), s2 = p[2] | (p[3] << 8), s3 = p[4] | (p[5] << 8);
        if (6ull + s1 + s2 + s3 > left) return ZF(Z_ERR_CORRUPT);
        const uint32_t s4 = left - 6 - s1 - s2 - s3;
        const uint32_t q = (regen + 3) / 4;
        if (3ull * q > regen) return ZF(Z_ERR_CORRUPT);
        const uint8_t *b = p + 6;
        int rc = huf_decode_stream(t, b, s1, lit, q);
        if (rc == Z_OK) rc = huf_decode_stream(t, b + s1, s2, lit + q, q);
        if (rc == Z_OK) rc = huf_decode_stream(t, b + s1 + s2, s3, lit + 2 * q, q);
        if (rc == Z_OK) rc = huf_decode_stream(t, b + s1 + s2 + s3, s4, lit + 3 * q, regen - 3 * q);
        if (rc != Z_OK) return rc;
    }
    *lit_ptr = lit;
    *lit_len = regen;
    return (int)(hdr + comp);
}

struct FrameState {
    uint64_t rep[3];
};

// Decodes the sequences section of a block and executes it.  dst/dst_pos: output so far (the window is the whole
// output, frames here are far smaller than any window).  Returns Z_OK or < 0.
ZHDN inline int decode_sequences(Tables &t, FrameState &fs, const uint8_t *src, uint32_t len, const uint8_t *lit, uint32_t lit_len,
                                 uint8_t *dst, uint64_t dst_cap, uint64_t *dst_pos) {
    if (len < 1) return ZF(Z_ERR_CORRUPT);
    uint32_t pos = 0;
    uint32_t nseq = src[pos++];
    if (nseq >= 128) {
        if (nseq == 255) {
            if (len < 3) return ZF(Z_ERR_CORRUPT);
            nseq = src[1] + ((uint32_t)src[2] << 8) + 0x7F00;
            pos = 3;
        } else {
            if (len < 2) return ZF(Z_ERR_CORRUPT);
            nseq = ((nseq - 128) << 8) + src[1];
            pos = 2;
        }
    }
    uint64_t out = *dst_pos;
    uint32_t lp = 0;
    if (nseq) {
        if (pos >= len) return ZF(Z_ERR_CORRUPT);
        const uint32_t modes = src[pos++];
        // (the two reserved bits of the modes byte are ignored, as libzstd does)
        int used = seq_table(t, 0, (modes >> 6) & 3, src + pos, len - pos);
        if (used < 0) return used;
        pos += used;
        used = seq_table(t, 1, (modes >> 4) & 3, src + pos, len - pos);
        if (used < 0) return used;
        pos += used;
        used = seq_table(t, 2, (modes >> 2) & 3, src + pos, len - pos);
        if (used < 0) return used;
        pos += used;
        BackBits bb;
        if (pos >= len || !bb.init(src + pos, len - pos)) return ZF(Z_ERR_CORRUPT);
        uint32_t sl = bb.read(t.ll_al), so = bb.read(t.of_al), sm = bb.read(t.ml_al);
        if (bb.off < 0) return ZF(Z_ERR_CORRUPT);
        for (uint32_t i = 0; i < nseq; ++i) {
            const int of_code = t.of[so].sym, ml_code = t.ml[sm].sym, ll_code = t.ll[sl].sym;
            if (of_code > 31 || ml_code > 52 || ll_code > 35) return ZF(Z_ERR_CORRUPT);
            const uint64_t ofv = (1ull << of_code) + bb.read(of_code);
            const uint32_t mlen = ml_base_of(ml_code) + bb.read(ml_bits_of(ml_code));
            const uint32_t llen = ll_base_of(ll_code) + bb.read(ll_bits_of(ll_code));
            if (i + 1 < nseq) {
                sl = t.ll[sl].base + bb.read(t.ll[sl].nbits);
                sm = t.ml[sm].base + bb.read(t.ml[sm].nbits);
                so = t.of[so].base + bb.read(t.of[so].nbits);
            }
            if (bb.off < 0) return ZF(Z_ERR_CORRUPT);
            uint64_t offset;
            if (ofv > 3) {
                offset = ofv - 3;
                fs.rep[2] = fs.rep[1];
                fs.rep[1] = fs.rep[0];
                fs.rep[0] = offset;
            } else {
                uint32_t idx = (uint32_t)ofv - 1;
                if (llen == 0) idx++;
                if (idx == 0) {
                    offset = fs.rep[0];
                } else {
                    offset = idx < 3 ? fs.rep[idx] : fs.rep[0] - 1;
                    if (idx > 1) fs.rep[2] = fs.rep[1];
                    fs.rep[1] = fs.rep[0];
                    fs.rep[0] = offset;
                }
            }
            if (llen > lit_len - lp || out + llen + mlen > dst_cap) return llen > lit_len - lp ? Z_ERR_CORRUPT : Z_ERR_NOSPACE;
            for (uint32_t k = 0; k < llen; ++k) dst[out + k] = lit[lp + k];
            out += llen;
            lp += llen;
            if (offset == 0 || offset > out) return ZF(Z_ERR_CORRUPT);
            for (uint32_t k = 0; k < mlen; ++k) dst[out + k] = dst[out - offset + k];
            out += mlen;
        }
        if (bb.off != 0) return ZF(Z_ERR_CORRUPT);
    } else if (pos != len) {
        return ZF(Z_ERR_CORRUPT);
    }
    const uint32_t rest = lit_len - lp;
    if (out + rest > dst_cap) return Z_ERR_NOSPACE;
    for (uint32_t k = 0; k < rest; ++k) dst[out + k] = lit[lp + k];
    out += rest;
    *dst_pos = out;
    return Z_OK;
}

// Whole frame.  lit: scratch >= 128 KiB.  *out_len receives the bytes produced.
ZHDN inline int decode_frame(Tables &t, const uint8_t *src, uint64_t len, uint8_t *dst, uint64_t dst_cap, uint8_t *lit, uint32_t lit_cap,
                             uint64_t *out_len) {
    FrameInfo fi;
    int rc = parse_frame_header(src, len, fi);
    if (rc != Z_OK) return rc;
    // the reference refuses frames without a content size (slow5_press.c:1206-1211)
    if (!fi.has_content_size) return ZF(Z_ERR_CORRUPT);
    if (fi.content_size > dst_cap) {
        *out_len = fi.content_size;
        return Z_ERR_NOSPACE;
    }
    t.huf_bits = 0;
    t.ll_al = t.of_al = t.ml_al = -1;
    FrameState fs;
    fs.rep[0] = 1;
    fs.rep[1] = 4;
    fs.rep[2] = 8;
    uint64_t pos = fi.header_bytes, out = 0;
    for (;;) {
        if (pos + 3 > len) return ZF(Z_ERR_CORRUPT);
        const uint32_t bh = src[pos] | ((uint32_t)src[pos + 1] << 8) | ((uint32_t)src[pos + 2] << 16);
        pos += 3;
        const bool last = bh & 1;
        const uint32_t type = (bh >> 1) & 3, bsize = bh >> 3;
        if (type == 0) {
            if (pos + bsize > len || out + bsize > fi.content_size) return ZF(Z_ERR_CORRUPT);
            for (uint32_t k = 0; k < bsize; ++k) dst[out + k] = src[pos + k];
            out += bsize;
            pos += bsize;
        } else if (type == 1) {
            if (pos + 1 > len || out + bsize > fi.content_size) return ZF(Z_ERR_CORRUPT);
            for (uint32_t k = 0; k < bsize; ++k) dst[out + k] = src[pos];
            out += bsize;
            pos += 1;
        } else if (type == 2) {
            if (pos + bsize > len || bsize > (128u << 10)) return ZF(Z_ERR_CORRUPT);
            const uint8_t *lp;
            uint32_t ll;
            const int used = decode_literals(t, src + pos, bsize, lit, lit_cap, &lp, &ll);
            if (used < 0) return used;
            rc = decode_sequences(t, fs, src + pos + used, bsize - used, lp, ll, dst, fi.content_size, &out);
            if (rc != Z_OK) return rc == Z_ERR_NOSPACE ? Z_ERR_CORRUPT : rc;  // more output than the header promised
            pos += bsize;
        } else {
            return ZF(Z_ERR_CORRUPT);
        }
        if (last) break;
    }
    if (out != fi.content_size) return ZF(Z_ERR_CORRUPT);
    if (fi.checksum) {
        if (pos + 4 > len) return ZF(Z_ERR_CORRUPT);
        if ((uint32_t)xxh64(dst, out) != rd32(src + pos)) return ZF(Z_ERR_CORRUPT);
        pos += 4;
    }
    // ZSTD_decompress would go on to decode further frames into the same buffer; with the buffer sized for the
    // first frame (slow5_press.c:1214) anything but trailing nothing is an error there
    if (pos != len) return ZF(Z_ERR_CORRUPT);
    *out_len = out;
    return Z_OK;
}

}  // namespace s5bz
