// zstd_enc_core.h -- the serial pieces of the Zstandard frame ENCODER (RFC 8878): frame header, block and
// literals-section headers, and the Huffman tree description (direct 4-bit weights or FSE-compressed weights).
//
// Replaces, together with zstd_encode_kernels.cu, ptr_compress_zstd (slow5lib/src/slow5_press.c:1183-1202:
// ZSTD_compress(dst, ZSTD_compressBound(n), src, n, 1) from system libzstd >= 1.3).  The compressed BYTES of that
// call are not pinned by the reference (its encode goldens are commented out, test/test_view.sh:204-214); the
// contract is the format: any conforming decoder (libzstd, the reference binary) regenerates the exact input.
// Frames written here hold one Huffman-coded literals section per block and no sequences.
//
// Plain serial code over byte pointers, host+device like zstd_core.h, so tests/dev/zstd_enc_check.cpp can run it
// on the CPU against libzstd during development; the product only ever runs it inside the kernel (lane 0).
#pragma once
#include "zstd_core.h"

namespace s5bz {

// ---- frame header (RFC 8878 3.1.1.1): magic, single-segment descriptor, frame content size ------------------
// dst must hold 12 bytes.  Returns bytes written.
ZHD int write_frame_header(uint8_t *dst, uint64_t content_size) {
    dst[0] = 0x28;
    dst[1] = 0xB5;
    dst[2] = 0x2F;
    dst[3] = 0xFD;
    if (content_size < 256) {
        dst[4] = 0x20;  // FCS flag 0 + Single_Segment: one byte of content size
        dst[5] = (uint8_t)content_size;
        return 6;
    }
    if (content_size < 65536 + 256) {
        const uint32_t v = (uint32_t)content_size - 256;
        dst[4] = 0x60;  // FCS flag 1: two bytes holding size - 256
        dst[5] = (uint8_t)v;
        dst[6] = (uint8_t)(v >> 8);
        return 7;
    }
    dst[4] = 0xA0;  // FCS flag 2: four bytes
    for (int i = 0; i < 4; ++i) dst[5 + i] = (uint8_t)(content_size >> (8 * i));
    return 9;
}

// Block_Header (3.1.1.2): last (1) | type (2) | size (21), little endian
ZHD uint32_t block_header(bool last, uint32_t type, uint32_t size) { return (last ? 1u : 0u) | (type << 1) | (size << 3); }

// Literals_Section_Header for Huffman-coded literals (3.1.1.3.1.1).  type: 2 = with tree, 3 = treeless.
// Returns the header length (3..5) and the header value in *v (little endian, low `len` bytes).
ZHD int literals_header(uint32_t type, bool four_streams, uint32_t regen, uint32_t comp, uint64_t *v) {
    if (!four_streams) {  // size format 0: single stream, 10 + 10 bits
        *v = type | (0u << 2) | ((uint64_t)regen << 4) | ((uint64_t)comp << 14);
        return 3;
    }
    if (regen < 1024 && comp < 1024) {
        *v = type | (1u << 2) | ((uint64_t)regen << 4) | ((uint64_t)comp << 14);
        return 3;
    }
    if (regen < 16384 && comp < 16384) {
        *v = type | (2u << 2) | ((uint64_t)regen << 4) | ((uint64_t)comp << 18);
        return 4;
    }
    *v = type | (3u << 2) | ((uint64_t)regen << 4) | ((uint64_t)comp << 22);
    return 5;
}
ZHD int literals_header_len(bool four_streams, uint32_t regen, uint32_t comp) {
    if (!four_streams || (regen < 1024 && comp < 1024)) return 3;
    return (regen < 16384 && comp < 16384) ? 4 : 5;
}

// ---- forward little-endian bit writer over a small byte buffer (must be zeroed by the caller) --------------
struct FwdBitWriter {
    uint8_t *p;
    uint32_t cap_bits;
    uint32_t pos;
    bool overflow;
    ZHD void put(uint32_t v, int n) {
        for (int i = 0; i < n; ++i) {
            if (pos >= cap_bits) {
                overflow = true;
                return;
            }
            if ((v >> i) & 1u) p[pos >> 3] |= (uint8_t)(1u << (pos & 7));
            ++pos;
        }
    }
};

// scratch of the weight coder (shared memory in the kernel)
struct WeightEnc {
    int16_t norm[16];
    uint16_t cumul[17];
    uint8_t table_sym[64];
    uint16_t state_tab[64];  // next-state values, grouped by symbol
    int32_t delta_find[16];
    uint32_t delta_nb[16];
    uint32_t count[16];
};

constexpr int HUF_TREE_MAX_BYTES = 160;  // buffer the tree description is built in (header byte + at most 127/128)

// FSE-compressed weights (4.2.1.1 / 4.1): normalised counts at accuracy log 5 or 6, two interleaved states
// (state 1 <-> even positions), written back to front.  Returns the byte count (without the leading header byte)
// or 0 when FSE cannot code this input (one distinct weight) or it does not fit.
ZHDN inline int fse_compress_weights(WeightEnc &e, const uint8_t *w, int n, uint8_t *dst, int cap) {
    if (n < 2) return 0;
    int max_w = 0;
    for (int s = 0; s < 16; ++s) e.count[s] = 0;
    for (int i = 0; i < n; ++i) {
        e.count[w[i]]++;
        if (w[i] > max_w) max_w = w[i];
    }
    int distinct = 0;
    for (int s = 0; s <= max_w; ++s) distinct += e.count[s] != 0;
    if (distinct < 2) return 0;
    const int al = n > 64 ? 6 : 5;
    const int size = 1 << al;
    // normalise: every present weight gets at least one slot, the largest class absorbs the rounding error
    int sum = 0, big = 0;
    for (int s = 0; s <= max_w; ++s) {
        int v = 0;
        if (e.count[s]) {
            v = (int)((e.count[s] * (uint32_t)size + (uint32_t)n / 2) / (uint32_t)n);
            if (v < 1) v = 1;
        }
        e.norm[s] = (int16_t)v;
        sum += v;
        if (e.count[s] > e.count[big]) big = s;
    }
    while (sum != size) {
        if (sum < size) {
            e.norm[big] = (int16_t)(e.norm[big] + (size - sum));
            sum = size;
        } else {  // take from the currently largest slot count, never below one
            int m = 0;
            for (int s = 1; s <= max_w; ++s)
                if (e.norm[s] > e.norm[m]) m = s;
            if (e.norm[m] <= 1) return 0;
            e.norm[m]--;
            --sum;
        }
    }
    // ---- normalised-count header (forward bits, 4.1.1)
    for (int i = 0; i < cap; ++i) dst[i] = 0;
    FwdBitWriter bw{dst, (uint32_t)cap * 8, 0, false};
    bw.put((uint32_t)(al - 5), 4);
    {
        int remaining = size + 1, threshold = size, nb = al + 1, sym = 0;
        bool prev0 = false;
        const int alphabet = max_w + 1;
        while (sym < alphabet && remaining > 1) {
            if (prev0) {
                int start = sym;
                while (sym < alphabet && e.norm[sym] == 0) ++sym;
                if (sym == alphabet) return 0;
                while (sym >= start + 3) {
                    start += 3;
                    bw.put(3, 2);
                }
                bw.put((uint32_t)(sym - start), 2);
            }
            int count = e.norm[sym++];
            const int mx = (2 * threshold - 1) - remaining;
            remaining -= count;
            ++count;
            if (count >= threshold) count += mx;
            bw.put((uint32_t)count, count < mx ? nb - 1 : nb);
            prev0 = count == 1;
            if (remaining < 1) return 0;
            while (remaining < threshold) {
                --nb;
                threshold >>= 1;
            }
        }
        if (remaining != 1) return 0;
    }
    const int hdr_bytes = (int)((bw.pos + 7) >> 3);
    // ---- compression table: symbols spread over the states, next-state table grouped by symbol
    {
        e.cumul[0] = 0;
        for (int s = 0; s <= max_w; ++s) e.cumul[s + 1] = (uint16_t)(e.cumul[s] + e.norm[s]);
        const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
        int pos = 0;
        for (int s = 0; s <= max_w; ++s)
            for (int i = 0; i < e.norm[s]; ++i) {
                e.table_sym[pos] = (uint8_t)s;
                pos = (pos + step) & mask;
            }
        uint16_t fill[16];
        for (int s = 0; s <= max_w; ++s) fill[s] = e.cumul[s];
        for (int u = 0; u < size; ++u) e.state_tab[fill[e.table_sym[u]]++] = (uint16_t)(size + u);
        int total = 0;
        for (int s = 0; s <= max_w; ++s) {
            const int c = e.norm[s];
            if (c == 0) {
                e.delta_nb[s] = ((uint32_t)(al + 1) << 16) - (uint32_t)size;
                e.delta_find[s] = 0;
            } else if (c == 1) {
                e.delta_nb[s] = ((uint32_t)al << 16) - (uint32_t)size;
                e.delta_find[s] = total - 1;
                total += 1;
            } else {
                const int max_bits_out = al - highest_bit((uint32_t)(c - 1));
                const uint32_t min_state_plus = (uint32_t)c << max_bits_out;
                e.delta_nb[s] = ((uint32_t)max_bits_out << 16) - min_state_plus;
                e.delta_find[s] = total - c;
                total += c;
            }
        }
    }
    // ---- payload: positions n-1 .. 0, state 1 codes the even positions, state 2 the odd ones; the first symbol
    // a state sees only selects its starting value (the smallest state of that symbol, so its update costs bits
    // and the decoder runs dry exactly there)
    FwdBitWriter pw{dst + hdr_bytes, (uint32_t)(cap - hdr_bytes) * 8, 0, false};
    uint32_t st[2] = {0, 0};
    bool live[2] = {false, false};
    for (int i = n - 1; i >= 0; --i) {
        const int k = i & 1;
        const int s = w[i];
        if (!live[k]) {
            const uint32_t nb_out = (e.delta_nb[s] + (1u << 15)) >> 16;
            const uint32_t v = (nb_out << 16) - e.delta_nb[s];
            st[k] = e.state_tab[(int)(v >> nb_out) + e.delta_find[s]];
            live[k] = true;
        } else {
            const uint32_t nb_out = (st[k] + e.delta_nb[s]) >> 16;
            pw.put(st[k] & ((1u << nb_out) - 1u), (int)nb_out);
            st[k] = e.state_tab[(int)(st[k] >> nb_out) + e.delta_find[s]];
        }
    }
    pw.put(st[1] & (uint32_t)(size - 1), al);  // state 2 first: the decoder reads state 1 first, from the top
    pw.put(st[0] & (uint32_t)(size - 1), al);
    pw.put(1, 1);  // end mark
    if (bw.overflow || pw.overflow) return 0;
    return hdr_bytes + (int)((pw.pos + 7) >> 3);
}

// Huffman_Tree_Description (4.2.1) for weights[0 .. nsym) with weights[nsym-1] != 0: the last weight is implied.
// dst must hold HUF_TREE_MAX_BYTES.  Returns the byte count, or 0 when the tree cannot be described (more than
// 128 transmitted weights and an FSE stream that does not fit below 128 bytes).
ZHDN inline int huf_write_tree(WeightEnc &e, const uint8_t *weights, int nsym, uint8_t *dst) {
    const int n = nsym - 1;  // transmitted weights
    if (n < 1) return 0;
    const int fse = fse_compress_weights(e, weights, n, dst + 1, 127);
    const int direct = n <= 128 ? 1 + (n + 1) / 2 : 0;
    if (fse > 0 && fse < 128 && (direct == 0 || 1 + fse < direct)) {
        dst[0] = (uint8_t)fse;
        return 1 + fse;
    }
    if (direct == 0) return 0;
    dst[0] = (uint8_t)(127 + n);
    for (int i = 0; i < n; i += 2) dst[1 + i / 2] = (uint8_t)((weights[i] << 4) | (i + 1 < n ? weights[i + 1] : 0));
    return direct;
}

}  // namespace s5bz

namespace s5bz {

// Code lengths (0 = unused, at most HUF_MAX_BITS) -> zstd weights and codes.  zstd numbers its codes from the
// longest (weight 1) upwards, symbols in increasing order inside one weight class (4.2.1.3), the mirror image of
// huf_build() in zstd_core.h.  Returns max_bits (0 when fewer than two symbols are coded); *nsym_out = index of
// the last coded symbol + 1.  code[s] is the value to append with its LSB at the current bit position.
ZHDN inline int huf_codes_from_lengths(const uint8_t *len, uint8_t *weights, uint16_t *code, int *nsym_out) {
    int max_bits = 0, nsym = 0, used = 0;
    for (int s = 0; s < 256; ++s)
        if (len[s]) {
            if (len[s] > max_bits) max_bits = len[s];
            nsym = s + 1;
            ++used;
        }
    *nsym_out = nsym;
    if (used < 2 || max_bits > HUF_MAX_BITS) return 0;
    uint32_t rank_count[HUF_MAX_BITS + 2];
    for (int b = 0; b <= HUF_MAX_BITS + 1; ++b) rank_count[b] = 0;
    for (int s = 0; s < 256; ++s) {
        weights[s] = len[s] ? (uint8_t)(max_bits + 1 - len[s]) : 0;
        rank_count[weights[s]]++;
    }
    uint32_t start[HUF_MAX_BITS + 2];
    uint32_t pos = 0;
    for (int w = 1; w <= max_bits; ++w) {
        start[w] = pos;
        pos += rank_count[w] << (w - 1);
    }
    if (pos != (1u << max_bits)) return 0;  // not a complete code
    for (int s = 0; s < 256; ++s) {
        const int w = weights[s];
        if (!w) continue;
        code[s] = (uint16_t)(start[w] >> (w - 1));
        start[w] += 1u << (w - 1);
    }
    return max_bits;
}

}  // namespace s5bz
