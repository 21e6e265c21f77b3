// huff_common.cuh -- pieces shared by the entropy encoders (deflate_kernels.cu, zstd_encode_kernels.cu): warp
// bitonic sort, length-limited Huffman code construction, and the shared-memory output bit buffer.
#pragma once
#include <cstdint>
#include "s5b_ptx.cuh"

namespace s5b {

constexpr int HC_OUT = 2048;       // output bit buffer bytes (multiple of 16)
constexpr int HC_OUT_SLACK = 96;

// ---- bitonic sort of n (power of two, <= 512) u32 keys in shared memory, ascending -------------
__device__ inline void warp_sort(uint32_t *a, int n, int lane) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < n / 2; t += 32) {
                // t-th compare-exchange pair of this stage
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = (i & k) == 0;
                const uint32_t x = a[i], y = a[p];
                if ((x > y) == up) {
                    a[i] = y;
                    a[p] = x;
                }
            }
            __syncwarp();
        }
    }
}

// ---- Huffman code lengths for `n` symbols with frequencies hist[] (0 = unused), limited to `limit` bits.
// Writes len[0..n).  Guarantees at least two coded symbols (like zlib's build_tree) so the code is complete.
// sortbuf: >= 512 entries (>= 32 for n <= 32), weight/parent: >= 2*n entries.  Whole warp calls it.
// Parallel parts: compaction of the used symbols, bitonic sort of just those, leaf depths (every lane walks its
// leaves up to the root), Kraft sum, length hand-out; lane 0 runs the two-queue merge and, only when a depth
// exceeds the limit, the repair loop.
__device__ inline void huffman_lengths(uint32_t *hist, int n, int limit, uint8_t *len, uint32_t *sortbuf, uint32_t *weight,
                                uint16_t *parent, uint16_t *bl_count, int lane) {
    // zlib forces two codes of non-zero frequency; mimic that so a lone symbol still gets a 1-bit code
    int used = 0;
    for (int s = lane; s < n; s += 32) used += hist[s] != 0;
#pragma unroll
    for (int d = 16; d; d >>= 1) used += __shfl_xor_sync(FULL, used, d);
    if (used < 2 && lane == 0) {
        for (int s = 0; s < n && used < 2; ++s)
            if (hist[s] == 0) {
                hist[s] = 1;
                ++used;
            }
    }
    used = max(used, 2);
    __syncwarp();
    // compact (frequency, symbol) keys of the used symbols, pad to a power of two, sort ascending
    int npad = 32;
    while (npad < used) npad <<= 1;
    {
        int base = 0;
        for (int s0 = 0; s0 < n; s0 += 32) {
            const int s = s0 + lane;
            const uint32_t f = s < n ? hist[s] : 0;
            if (s < n) len[s] = 0;
            const uint32_t m = __ballot_sync(FULL, f != 0);
            if (f) sortbuf[base + __popc(m & ((1u << lane) - 1u))] = (min(f, 0x7fffffu) << 9) | (uint32_t)s;
            base += __popc(m);
        }
        for (int i = used + lane; i < npad; i += 32) sortbuf[i] = 0xffffffffu;
    }
    __syncwarp();
    warp_sort(sortbuf, npad, lane);
    // leaves 0..used-1 in ascending weight; internal nodes used..2*used-2 are created in ascending weight too
    for (int i = lane; i < used; i += 32) weight[i] = sortbuf[i] >> 9;
    __syncwarp();
    if (lane == 0) {
        int li = 0, ii = used, next = used;
        for (int j = 0; j < used - 1; ++j) {
            int pick[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (li < used && (ii >= next || weight[li] <= weight[ii])) pick[t] = li++;
                else pick[t] = ii++;
            }
            weight[next] = weight[pick[0]] + weight[pick[1]];
            parent[pick[0]] = (uint16_t)next;
            parent[pick[1]] = (uint16_t)next;
            ++next;
        }
    }
    __syncwarp();
    // leaf depths: walk up to the root (node 2*used-2); clamp to the limit and sum the Kraft terms in 2^-limit units
    const int root = 2 * used - 2;
    if (lane < 16) bl_count[lane] = 0;
    __syncwarp();
    uint32_t kraft = 0;
    int over = 0;
    for (int i = lane; i < used; i += 32) {
        int d = 0;
        for (int v = i; v != root; v = parent[v]) ++d;
        if (d > limit) {
            d = limit;
            over = 1;
        }
        kraft += 1u << (limit - d);
        weight[i] = (uint32_t)d;  // (the leaf weights are no longer needed)
        atomicAdd(reinterpret_cast<unsigned int *>(bl_count) + (d >> 1), (d & 1) ? 0x10000u : 1u);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        kraft += __shfl_xor_sync(FULL, kraft, d);
        over |= __shfl_xor_sync(FULL, over, d);
    }
    __syncwarp();
    if (over) {
        // length limiting: the clamped depths over-subscribe the code by `excess` units of 2^-limit; repair it one
        // unit at a time the way zlib's gen_bitlen does (push a leaf from the deepest level above the limit one
        // level down and hang one clamped leaf next to it), then hand the lengths out again: longest codes to the
        // rarest symbols (leaves are sorted by weight)
        if (lane == 0) {
            int excess = (int)kraft - (1 << limit);
            while (excess > 0) {
                int bits = limit - 1;
                while (bl_count[bits] == 0) --bits;
                bl_count[bits]--;
                bl_count[bits + 1] += 2;
                bl_count[limit]--;
                --excess;
            }
            int i = 0;
            for (int bits = limit; bits >= 1; --bits)
                for (int c = bl_count[bits]; c > 0; --c) weight[i++] = (uint32_t)bits;
        }
        __syncwarp();
    }
    for (int i = lane; i < used; i += 32) len[sortbuf[i] & 511u] = (uint8_t)weight[i];
    __syncwarp();
}

// canonical codes (RFC 1951 3.2.2) for len[0..n), n <= 288, lengths <= 15; bit-reversed for LSB-first packing when
// `reversed`.  bl_count is a 16-entry scratch.  Whole warp: per-length counts by shared-memory atomics, the
// first code of every length in registers, then 32 symbols per round ranked inside their length class with
// __match_any_sync (symbols of one length get consecutive codes in symbol order).
__device__ inline void canonical_codes(const uint8_t *len, int n, uint16_t *code, uint16_t *bl_count, int lane,
                                       bool reversed = true) {
    if (lane < 16) bl_count[lane] = 0;
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const uint32_t l = len[s];
        if (l) atomicAdd(reinterpret_cast<unsigned int *>(bl_count) + (l >> 1), (l & 1) ? 0x10000u : 1u);
    }
    __syncwarp();
    // lane b (1..15) keeps the next code of length b
    uint32_t next = 0;
    {
        uint32_t c = 0;
#pragma unroll
        for (int b = 1; b <= 15; ++b) {
            c = (c + (b > 1 ? bl_count[b - 1] : 0)) << 1;
            if (lane == b) next = c;
        }
    }
    for (int s0 = 0; s0 < n; s0 += 32) {
        const int s = s0 + lane;
        const uint32_t l = s < n ? len[s] : 0;
        const unsigned same = __match_any_sync(FULL, l);
        const uint32_t base = __shfl_sync(FULL, next, (int)l & 15);  // l == 0 reads lane 0's unused value
        if (s < n) {
            uint32_t cw = 0;
            if (l) {
                cw = base + __popc(same & ((1u << lane) - 1u));
                if (reversed) cw = __brev(cw) >> (32 - l);
            }
            code[s] = (uint16_t)cw;
        }
        // the lane that owns length b advances its counter by the number of symbols of that length in this round
#pragma unroll
        for (int b = 1; b <= 15; ++b) {
            const unsigned mb = __ballot_sync(FULL, l == (uint32_t)b);
            if (lane == b) next += __popc(mb);
        }
    }
    __syncwarp();
}

// ---- output bit buffer (warp-uniform bookkeeping; lanes OR their bits in with shared-memory atomics) ----
struct BitOut {
    uint32_t *buf;    // smem words; bit 0 of buf[0] <-> bit 0 of the byte at gbase
    uint8_t *gbase;   // 16-byte aligned global address of buf[0]
    uint32_t bitpos;  // next free bit
    uint32_t head;    // first valid byte of buf (stream start not 16-byte aligned), only before the first flush
    uint64_t written; // bytes already stored (excluding head padding)

    __device__ __forceinline__ void put(uint32_t pos, uint32_t bits, uint32_t nbits) const {
        if (nbits == 0) return;
        const uint32_t w = pos >> 5, sh = pos & 31u;
        atomicOr(&buf[w], bits << sh);
        if (sh + nbits > 32) atomicOr(&buf[w + 1], bits >> (32 - sh));
    }
    // store whole 16-byte segments, carry the rest (including the partly filled last byte) to the front
    __device__ inline void flush(int lane, bool final) {
        __syncwarp();
        const uint32_t nbytes = final ? (bitpos + 7) >> 3 : bitpos >> 3;
        const uint32_t wseg = final ? (nbytes + 15) >> 4 : nbytes >> 4;
        const uint8_t *b8 = reinterpret_cast<const uint8_t *>(buf);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(buf);
        uint4 *g4 = reinterpret_cast<uint4 *>(gbase);
        for (uint32_t seg = lane; seg < wseg; seg += 32) {
            const uint32_t lo = seg * 16, hi = lo + 16;
            if (lo >= head && hi <= nbytes) g4[seg] = s4[seg];
        }
        if (wseg == 0) return;
        // the ragged ends leave a byte per lane: the front of segment 0 when the stream does not start on a 16-byte
        // boundary, the back of the last segment at the end of the stream
        if (head) {
            const uint32_t i = head + lane;
            if (i < min(16u, nbytes)) gbase[i] = b8[i];
        }
        if (final && (nbytes & 15u) && (wseg > 1 || !head)) {
            const uint32_t i = (wseg - 1) * 16 + lane;
            if (i < nbytes) gbase[i] = b8[i];
        }
        const uint32_t full = final ? nbytes : wseg * 16;
        written += full - head;
        head = 0;
        // carry: bytes [full, ceil(bitpos/8)) move to the front, everything else becomes zero
        const uint32_t tail_bytes = final ? 0 : ((bitpos + 7) >> 3) - full;
        uint32_t keep[1];
        // tail < 16 bytes + slack: at most 4 words plus the slack words (one per lane is plenty)
        const uint32_t tail_words = (tail_bytes + 3) >> 2;
        keep[0] = lane < (int)tail_words ? buf[(full >> 2) + lane] : 0u;
        __syncwarp();
        for (uint32_t i = lane; i < (HC_OUT + HC_OUT_SLACK) / 4; i += 32) buf[i] = 0;
        __syncwarp();
        if (lane < (int)tail_words) buf[lane] = keep[0];
        gbase += full;
        bitpos -= full * 8;
        __syncwarp();
    }
};


}  // namespace s5b
