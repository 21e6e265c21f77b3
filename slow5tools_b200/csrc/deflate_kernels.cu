// deflate_kernels.cu -- sm_100a zlib/DEFLATE encoder for batches of records, four kernels.
//
// Replaces, for whole batches, ptr_compress_zlib / ptr_compress_zlib_solo (slow5lib/src/slow5_press.c:
// 837-913: deflateInit2(level 6, wbits 15, memLevel 8, default strategy) + deflate(Z_FINISH)), producing
// one complete zlib stream per record (slow5.c:4046).  The bytes are NOT zlib's bytes -- no encoder other
// than zlib can reproduce its match choices -- the contract is: the reference's inflate returns exactly the
// input (Adler-32 included), and the compressed size is within the tolerance stated in DESIGN.md of zlib
// level 6 on BLOW5 records.
//
// Encoder shape, chosen from what the data looks like (on svb-zd records Huffman coding is
// where the gain is; LZ77 matching buys < 2 % and only in the zero-filled key bytes):
//   * a record is cut into blocks: at a caller-supplied split point (the boundary between the header + key
//     bytes and the data bytes of the svb stream), then every DEF_BLOCK bytes; every block gets its own
//     dynamic Huffman code;
//   * a block is coded in one of two modes.  LITERAL: no matches, four bytes per lane and iteration straight
//     from global memory.  RUN: literals plus distance-1 run matches found with warp ballots inside 32-byte
//     strips -- used for the part before the split point and for any block in which at least half the bytes
//     repeat their predecessor.  The part before the split point (record header fields, sample count, svb-zd control
//     bytes: the same kind of bytes in every record of every file) does not get a code of its own: it is coded RUN-style
//     under a code fixed ahead of time (tools/gen_deflate_canned.py), still as a dynamic block whose header is the same
//     bits every time -- no histogram, sort, tree, canonical codes or header construction for it (that cost as much as
//     the 4x larger data part), +0.6 % in size;
//   * the work is split by the shape of its parallelism:
//       deflate_count_kernel  one warp per record: Adler-32, byte / length-symbol histogram of every block,
//                             compaction of the used symbols, bitonic sort in registers -> sorted
//                             (frequency, symbol) list per block in the workspace;
//       deflate_tree_kernel   one THREAD per block: the inherently serial part of Huffman construction
//                             (Moffat-Katajainen in-place two-queue merge, depths, 15-bit length limiting) on a
//                             shared-memory row per thread, 32 independent trees per warp instead of one lane
//                             working while 31 wait;
//       deflate_header_kernel one THREAD per block: canonical codes (the table the token loops look up), run-length
//                             coding of the code lengths, the 19-symbol code-length code, the header bits -- into the
//                             workspace (a warp used to do this for one block at a time inside the emit kernel:
//                             4 400 warp instructions per block);
//       deflate_emit_kernel   one warp per record: block header copied from the workspace (or the canned constant),
//                             token bits placed by a warp prefix scan and OR-ed into a shared-memory bit buffer that
//                             leaves as 128-bit stores; stored blocks when a block would not shrink (canned blocks are
//                             measured after they have been coded and taken back if they did not); Adler-32 trailer.
//     (the round-1 single kernel spent 35 % of its issue slots in lane 0's merge loop and the shared-memory sort)
#include <cstdlib>
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "huff_common.cuh"
#include "../../include/slow5b200.h"

namespace s5b {

namespace defl {

constexpr int DEF_BLOCK = 6144;  // max input bytes per deflate block (multiple of 128)
constexpr int DEF_ROW = 288;     // workspace row: entries per block
constexpr int DEF_OUT = HC_OUT;
constexpr int DEF_OUT_SLACK = HC_OUT_SLACK;
constexpr uint32_t ADLER_MOD = 65521u;
constexpr int CNT_WARPS = 4, EMIT_WARPS = 4, TREE_THREADS = 64;
constexpr int CANNED_MAX_BLOCK = 1152;  // canned blocks are coded first and measured afterwards: header + 12 bits per byte must
                                       // fit the emit kernel's bit buffer without a flush in between
constexpr int HDR_WORDS = 132;    // 14 + 19 * 3 + 287 * 14 bits at most
constexpr int TREE_STRIDE = 290;  // u16 entries per row: 145 words, odd -> lanes on the same index hit different banks
enum : uint32_t { MODE_LIT = 0, MODE_RUN = 1, MODE_CANNED = 2 };

#include "deflate_canned.inc"
static_assert(CANNED_MAX_LEN <= 12 && CANNED_HDR_BITS <= 512, "deflate_emit_kernel reserves 1.5 bytes per byte + 160 for a canned block");

struct DefWork {       // views into the caller's workspace
    uint32_t *nblk;    // [n]      blocks per record (scanned into blk_off)
    uint64_t *blk_off; // [n + 1]
    uint32_t *adler;   // [n]
    uint4 *binfo;      // [blocks] x: used | mode << 16 | has_match << 24, y: token bits (tree kernel)
    uint32_t *keys;    // [blocks][DEF_ROW] ascending (frequency << 9 | symbol), `used` entries
    uint8_t *lens;     // [blocks][DEF_ROW] code length of symbol s (tree kernel)
    uint32_t *tab;     // [blocks][DEF_ROW] code (bit-reversed) | length << 16 of symbol s (header kernel)
    uint32_t *hdr;     // [blocks][HDR_WORDS] the dynamic block header after BFINAL / BTYPE, binfo.z bits of it (header kernel)
    void *scan_scratch;
    uint64_t max_blocks;
    uint32_t canned;   // blocks before the split point use the canned code (S5B_DEFLATE_CANNED=0 turns it off)
    uint32_t hdr_by_warp;  // S5B_DEFLATE_HDR=warp: block headers built inside the emit kernel (the older way, kept for A/B runs)
};

__constant__ uint8_t c_def_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__host__ __device__ inline uint32_t def_nblocks(uint32_t len, uint32_t split) {
    const uint32_t na = split ? (split + DEF_BLOCK - 1) / DEF_BLOCK : 0;
    const uint32_t nb = (len - split + DEF_BLOCK - 1) / DEF_BLOCK;
    return na + nb ? na + nb : 1;  // an empty record still gets one (final, empty) block
}
__device__ __forceinline__ void def_block_range(uint32_t len, uint32_t split, uint32_t k, uint32_t &b0, uint32_t &b1) {
    const uint32_t na = split ? (split + DEF_BLOCK - 1) / DEF_BLOCK : 0;
    if (k < na) {
        b0 = k * DEF_BLOCK;
        b1 = min(b0 + DEF_BLOCK, split);
    } else {
        b0 = split + (k - na) * DEF_BLOCK;
        b1 = min(b0 + DEF_BLOCK, len);
    }
}

// match length 3..32 -> (code - 257, extra bits, extra value) per RFC 1951 3.2.5
__device__ __forceinline__ void len_code(uint32_t m, uint32_t &sym, uint32_t &xbits, uint32_t &xval) {
    if (m <= 10) {
        sym = m - 3;
        xbits = 0;
        xval = 0;
    } else if (m <= 18) {
        sym = 8 + ((m - 11) >> 1);
        xbits = 1;
        xval = (m - 11) & 1u;
    } else {  // 19..34
        sym = 12 + ((m - 19) >> 2);
        xbits = 2;
        xval = (m - 19) & 3u;
    }
}

// Token of this lane in a 32-byte strip (one byte per lane): a literal, the start of a distance-1 run match of length
// m (3..32, entirely inside the strip), or nothing (covered by a match).  b = the lane's byte (0x100 past the end),
// prev0 = the byte before the strip (0x200 when there is none).
struct Token {
    uint32_t kind;  // 0 none, 1 literal, 2 match
    uint32_t mlen;
};
__device__ __forceinline__ Token strip_token(uint32_t b, uint32_t prev0, int lane) {
    Token tk;
    uint32_t prev = __shfl_up_sync(FULL, b, 1);
    if (lane == 0) prev = prev0;
    const bool valid = b < 0x100u;
    const bool eq = valid && b == prev;
    const uint32_t mask = __ballot_sync(FULL, eq);
    tk.kind = valid ? 1u : 0u;
    tk.mlen = 0;
    if (eq) {
        const uint32_t below = ~mask & ((1u << lane) - 1u);
        const uint32_t s = below ? 32u - __clz(below) : 0u;            // first lane of this run of equal bytes
        const uint32_t above = ~mask & ~((2u << lane) - 1u);
        const uint32_t e = above ? (uint32_t)__ffs(above) - 2u : 31u;  // last lane of the run
        const uint32_t m = e - s + 1;
        if (m >= 3) {
            tk.kind = (uint32_t)lane == s ? 2u : 0u;
            tk.mlen = m;
        }
    }
    return tk;
}

// ---- reading a block four bytes per lane from global memory, any alignment --------------------------------------
// Strip t of the block = bytes [128 t, 128 t + 128); lane l owns bytes 4 l .. 4 l + 3 of it.  Words are loaded aligned
// and shifted into place with the neighbour lane's word; lane 31's neighbour is lane 0 of the NEXT strip, which is
// loaded one iteration ahead anyway.
struct WordReader {
    const uint32_t *w;   // aligned word at or below the block's first byte
    const uint32_t *end; // no loads at or beyond (the slab's capacity)
    uint32_t sh;         // (first byte address & 3) * 8
    uint32_t cur, nxt;
    __device__ __forceinline__ uint32_t ld(uint32_t i) const {
        const uint32_t *p = w + i;
        return p < end ? __ldg(p) : 0u;
    }
    __device__ __forceinline__ void start(const uint8_t *first, const uint8_t *slab_end, int lane) {
        const uint32_t sk = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 3u);
        w = reinterpret_cast<const uint32_t *>(first - sk);
        end = reinterpret_cast<const uint32_t *>(slab_end);
        sh = sk * 8u;
        cur = ld(lane);
        nxt = ld(32 + lane);
    }
    // the lane's four bytes of strip t (call with t = 0, 1, 2, ... in order)
    __device__ __forceinline__ uint32_t next(uint32_t t, int lane) {
        uint32_t up = __shfl_down_sync(FULL, cur, 1);
        const uint32_t n0 = __shfl_sync(FULL, nxt, 0);
        if (lane == 31) up = n0;
        const uint32_t v = __funnelshift_r(cur, up, sh);
        cur = nxt;
        nxt = ld(32 * (t + 2) + lane);
        return v;
    }
};

// ---- bitonic sort of 32 * R keys held R per lane (element i = register i / 32 of lane i % 32), ascending ----------
template <int R>
__device__ __forceinline__ void reg_sort(uint32_t (&v)[R], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & jr) == 0) {
                        const bool up = (((r << 5) | lane) & k) == 0;
                        const uint32_t a = v[r], b = v[r | jr];
                        const bool sw = (a > b) == up;
                        v[r] = sw ? b : a;
                        v[r | jr] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t o = __shfl_xor_sync(FULL, v[r], j);
                    const bool up = (((r << 5) | lane) & k) == 0;
                    const bool lower = (lane & j) == 0;
                    v[r] = (lower == up) ? min(v[r], o) : max(v[r], o);
                }
            }
        }
    }
}

struct __align__(16) CntWarpSmem {
    uint32_t hist[4][DEF_ROW];  // four copies (lane & 3) so that common bytes do not serialise the atomics
    uint32_t keys[512];
};

// ---- plan: blocks per record, argument checks -------------------------------------------------------------------
__global__ void deflate_plan_kernel(const DeflateArgs a, DefWork w) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_reads) return;
    const uint64_t ioff = a.in_off[r];
    const uint32_t ilen = a.in_len[r];
    const uint64_t ocap = a.out_off[r + 1] - a.out_off[r];
    uint32_t split = a.split ? a.split[r] : 0;
    if (split >= ilen) split = 0;
    // worst case: stored blocks (5 bytes each) + zlib header + Adler-32 + one byte of bit padding per block
    const uint64_t nblocks_max = (uint64_t)ilen / DEF_BLOCK + 2;
    int32_t st = S5B_OK;
    if (ioff + ilen > a.in_capacity) st = S5B_ERR_ARG;
    else if (ocap < (uint64_t)ilen + 6 * nblocks_max + 8) st = S5B_ERR_NOSPACE;
    a.status[r] = st;
    if (st != S5B_OK) a.out_len[r] = 0;
    w.nblk[r] = st == S5B_OK ? def_nblocks(ilen, split) : 0u;
}

// ---- kernel 1: histograms -> sorted symbol lists ------------------------------------------------------------------
__global__ void __launch_bounds__(CNT_WARPS * 32) deflate_count_kernel(const DeflateArgs a, DefWork wk) {
    __shared__ CntWarpSmem smem[CNT_WARPS];
    CntWarpSmem &ws = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint8_t *slab_end = a.in + a.in_capacity;
    for (;;) {
        unsigned long long r = 0;
        if (lane == 0) r = atomicAdd(a.work_counter, 1ULL);
        r = __shfl_sync(FULL, r, 0);
        if (r >= a.n_reads) break;
        if (a.status[r] != S5B_OK) continue;
        const uint8_t *src = a.in + a.in_off[r];
        const uint32_t ilen = a.in_len[r];
        uint32_t split = a.split ? a.split[r] : 0;
        if (split >= ilen) split = 0;
        const uint32_t nb = def_nblocks(ilen, split);
        const uint64_t blk0 = wk.blk_off[r];
        uint32_t ad_a = 1, ad_b = 0;
        for (uint32_t k = 0; k < nb; ++k) {
            uint32_t b0, b1;
            def_block_range(ilen, split, k, b0, b1);
            const uint32_t blen = b1 - b0;
            uint32_t mode = (split && b1 <= split) ? MODE_RUN : MODE_LIT;
            uint32_t nmatch = 0;
            if (mode == MODE_RUN && wk.canned && blen <= (uint32_t)CANNED_MAX_BLOCK) {
                // canned code: nothing to count.  Adler-32 four bytes per lane; whether the block shrinks is seen when it is
                // coded (the emit kernel measures it and falls back to a stored block)
                WordReader rd;
                rd.start(src + b0, slab_end, lane);
                uint32_t s1 = 0, s2 = 0;
                const uint32_t nstrip = (blen + 127) >> 7;
                for (uint32_t t = 0; t < nstrip; ++t) {
                    uint32_t v = rd.next(t, lane);
                    const uint32_t pos = (t << 7) + 4u * lane;
                    const uint32_t nv = pos >= blen ? 0u : min(4u, blen - pos);
                    if (nv < 4) v &= nv ? (1u << (8 * nv)) - 1u : 0u;
                    const uint32_t sum = __dp4a(v, 0x01010101u, 0u), wsum = __dp4a(v, 0x03020100u, 0u);
                    s1 += sum;
                    s2 += (blen - pos) * sum - wsum;
                }
                s2 %= ADLER_MOD;
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    s1 += __shfl_xor_sync(FULL, s1, d);
                    s2 += __shfl_xor_sync(FULL, s2, d);
                }
                ad_b = (uint32_t)((ad_b + (uint64_t)blen * ad_a + s2) % ADLER_MOD);
                ad_a = (ad_a + s1) % ADLER_MOD;
                if (lane == 0) wk.binfo[blk0 + k] = make_uint4((MODE_CANNED << 16) | (1u << 24), 0u, 0u, 0u);
                continue;
            }
            for (int pass = 0; pass < 2; ++pass) {
                WordReader rd;
                if (mode == MODE_LIT) rd.start(src + b0, slab_end, lane);  // (in flight while the histogram is cleared)
                for (int i = lane; i < 4 * DEF_ROW; i += 32) (&ws.hist[0][0])[i] = 0;
                __syncwarp();
                uint32_t s1 = 0, s2 = 0, eqc = 0;
                nmatch = 0;
                if (mode == MODE_LIT) {
                    uint32_t *h = ws.hist[lane & 3];
                    uint32_t carry = 0;  // last byte of the previous strip (for the repeat count only)
                    const uint32_t nstrip = (blen + 127) >> 7;
                    for (uint32_t t = 0; t < nstrip; ++t) {
                        uint32_t v = rd.next(t, lane);
                        const uint32_t pos = (t << 7) + 4u * lane;
                        const uint32_t nv = pos >= blen ? 0u : min(4u, blen - pos);
                        if (nv < 4) v &= nv ? (1u << (8 * nv)) - 1u : 0u;
                        // Adler-32 partial sums: s1 += sum d, s2 += sum (blen - position) d
                        const uint32_t sum = __dp4a(v, 0x01010101u, 0u), wsum = __dp4a(v, 0x03020100u, 0u);
                        s1 += sum;
                        s2 += (blen - pos) * sum - wsum;
                        if (nv == 4) {
                            atomicAdd(&h[v & 255u], 1u);
                            atomicAdd(&h[(v >> 8) & 255u], 1u);
                            atomicAdd(&h[(v >> 16) & 255u], 1u);
                            atomicAdd(&h[v >> 24], 1u);
                        } else {
                            for (uint32_t q = 0; q < nv; ++q) atomicAdd(&h[(v >> (8 * q)) & 255u], 1u);
                        }
                        // bytes equal to their predecessor (mode choice)
                        uint32_t pv = __shfl_up_sync(FULL, v, 1);
                        if (lane == 0) pv = carry;
                        carry = __shfl_sync(FULL, v, 31);
                        const uint32_t x = v ^ __funnelshift_l(pv, v, 8);
                        eqc += nv == 4 ? (uint32_t)__popc(__vcmpeq4(x, 0u)) >> 3 : 0u;
                    }
                } else {
                    uint32_t prev0 = b0 ? (uint32_t)src[b0 - 1] : 0x200u;  // the byte before the block is matchable
                    for (uint32_t t0 = 0; t0 < blen; t0 += 32) {
                        const uint32_t i = t0 + lane;
                        const uint32_t b = i < blen ? (uint32_t)src[b0 + i] : 0x100u;
                        if (i < blen) {
                            s1 += b;
                            s2 += (blen - i) * b;
                        }
                        const Token tk = strip_token(b, prev0, lane);
                        prev0 = __shfl_sync(FULL, b, 31);
                        if (tk.kind == 1) {
                            atomicAdd(&ws.hist[0][b], 1u);
                        } else if (tk.kind == 2) {
                            uint32_t sym, xb, xv;
                            len_code(tk.mlen, sym, xb, xv);
                            atomicAdd(&ws.hist[0][257 + sym], 1u);
                            ++nmatch;
                        }
                    }
                }
                s2 %= ADLER_MOD;
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    s1 += __shfl_xor_sync(FULL, s1, d);
                    s2 += __shfl_xor_sync(FULL, s2, d);
                    eqc += __shfl_xor_sync(FULL, eqc, d);
                    nmatch += __shfl_xor_sync(FULL, nmatch, d);
                }
                if (pass == 0) {
                    ad_b = (uint32_t)((ad_b + (uint64_t)blen * ad_a + s2) % ADLER_MOD);
                    ad_a = (ad_a + s1) % ADLER_MOD;
                }
                __syncwarp();
                // a block that mostly repeats its previous byte is better off with run matches: count it again
                if (pass == 0 && mode == MODE_LIT && blen >= 64 && eqc * 2 > blen) {
                    mode = MODE_RUN;
                    continue;
                }
                break;
            }
            // ---- compact the used symbols (end of block included), sort by (frequency, symbol)
            int used = 0;
            for (int s0 = 0; s0 < DEF_ROW; s0 += 32) {
                const int s = s0 + lane;
                uint32_t f = ws.hist[0][s] + ws.hist[1][s] + ws.hist[2][s] + ws.hist[3][s];
                if (s == 256) f = 1;
                if (s >= 286) f = 0;
                const uint32_t m = __ballot_sync(FULL, f != 0);
                if (f) ws.keys[used + __popc(m & ((1u << lane) - 1u))] = (min(f, 0x7fffffu) << 9) | (uint32_t)s;
                used += __popc(m);
            }
            __syncwarp();
            if (used < 2) {
                // zlib forces two codes so that the code is complete: a second symbol of frequency 1 (used == 1 means
                // only the end-of-block symbol is there: an empty block)
                if (lane == 0) ws.keys[1] = (1u << 9) | (ws.keys[0] == ((1u << 9) | 256u) ? 0u : 255u);
                used = 2;
                __syncwarp();
            }
            uint32_t *row = wk.keys + (blk0 + k) * DEF_ROW;
            if (used <= 32) {
                uint32_t v[1] = {lane < used ? ws.keys[lane] : 0xffffffffu};
                reg_sort<1>(v, lane);
                if (lane < used) row[lane] = v[0];
            } else if (used <= 64) {
                uint32_t v[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) v[q] = q * 32 + lane < used ? ws.keys[q * 32 + lane] : 0xffffffffu;
                reg_sort<2>(v, lane);
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (q * 32 + lane < used) row[q * 32 + lane] = v[q];
            } else if (used <= 128) {
                uint32_t v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = q * 32 + lane < used ? ws.keys[q * 32 + lane] : 0xffffffffu;
                reg_sort<4>(v, lane);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q * 32 + lane < used) row[q * 32 + lane] = v[q];
            } else {
                const int npad = used <= 256 ? 256 : 512;
                for (int i = used + lane; i < npad; i += 32) ws.keys[i] = 0xffffffffu;
                __syncwarp();
                warp_sort(ws.keys, npad, lane);
                for (int i = lane; i < used; i += 32) row[i] = ws.keys[i];
            }
            if (lane == 0) wk.binfo[blk0 + k] = make_uint4((uint32_t)used | (mode << 16) | (nmatch ? 1u << 24 : 0u), 0u, 0u, 0u);
            __syncwarp();
        }
        if (lane == 0) wk.adler[r] = (ad_b << 16) | ad_a;
    }
}

// ---- kernel 2: one thread per block: code lengths from the sorted frequencies ----------------------------------------
// Moffat & Katajainen's in-place construction on a row of 16-bit words: (1) two-queue merge leaving parent indices,
// (2) parent indices -> internal node depths, (3) internal depths -> leaf depths (= code lengths, non-increasing along
// the ascending frequencies).  Then zlib-style repair when a length exceeds `limit`.
__device__ __forceinline__ void mk_lengths(uint16_t *A, int n, int limit) {
    if (n == 2) {
        A[0] = A[1] = 1;
        return;
    }
    A[0] = (uint16_t)(A[0] + A[1]);
    int root = 0, leaf = 2;
    for (int next = 1; next < n - 1; ++next) {
        uint32_t wsum;
        if (leaf >= n || A[root] < A[leaf]) {
            wsum = A[root];
            A[root++] = (uint16_t)next;
        } else {
            wsum = A[leaf++];
        }
        if (leaf >= n || (root < next && A[root] < A[leaf])) {
            wsum += A[root];
            A[root++] = (uint16_t)next;
        } else {
            wsum += A[leaf++];
        }
        A[next] = (uint16_t)wsum;
    }
    A[n - 2] = 0;
    for (int next = n - 3; next >= 0; --next) A[next] = (uint16_t)(A[A[next]] + 1);
    int avbl = 1, used = 0, dpth = 0;
    root = n - 2;
    int next = n - 1;
    while (avbl > 0) {
        while (root >= 0 && A[root] == dpth) {
            ++used;
            --root;
        }
        while (avbl > used) {
            A[next--] = (uint16_t)dpth;
            --avbl;
        }
        avbl = 2 * used;
        ++dpth;
        used = 0;
    }
    if (A[0] <= limit) return;
    // length limiting (rare): clamp, then repair the over-subscribed code one unit of 2^-limit at a time the way zlib's
    // gen_bitlen does, and hand the lengths out again, longest codes to the rarest symbols
    uint16_t cnt[16];
    for (int b = 0; b < 16; ++b) cnt[b] = 0;
    uint32_t kraft = 0;
    for (int i = 0; i < n; ++i) {
        const int l = min((int)A[i], limit);
        ++cnt[l];
        kraft += 1u << (limit - l);
    }
    int excess = (int)kraft - (1 << limit);
    while (excess > 0) {
        int bits = limit - 1;
        while (cnt[bits] == 0) --bits;
        --cnt[bits];
        cnt[bits + 1] += 2;
        --cnt[limit];
        --excess;
    }
    int i = 0;
    for (int bits = limit; bits >= 1; --bits)
        for (int c = cnt[bits]; c > 0; --c) A[i++] = (uint16_t)bits;
}

// Blocks coded under the canned code need neither a tree nor a header, and they sit between the others (block 0 of every
// record): a warp that took 32 consecutive blocks would run with half its lanes idle.  A warp takes a window of 64 consecutive
// blocks instead and hands the ones that need work to its lanes, 32 per round (list[]: 64 entries of warp-private shared memory).
// Returns the number of blocks to do; round r, lane l: block first + list[32 r + l].
__device__ __forceinline__ int window_blocks(const DefWork &wk, uint64_t first, uint64_t nblocks, int lane, uint8_t *list) {
    int total = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint64_t b = first + 32 * h + lane;
        bool need = false;
        if (b < nblocks) {
            const uint32_t x = wk.binfo[b].x;
            need = ((x >> 16) & 0xffu) != MODE_CANNED && (x & 0xffffu) >= 2;
        }
        const uint32_t m = __ballot_sync(FULL, need);
        if (need) list[total + __popc(m & ((1u << lane) - 1u))] = (uint8_t)(32 * h + lane);
        total += __popc(m);
    }
    __syncwarp();
    return total;
}

__global__ void __launch_bounds__(TREE_THREADS) deflate_tree_kernel(DefWork wk, const uint64_t *n_blocks_ptr) {
    __shared__ uint16_t rows[TREE_THREADS * TREE_STRIDE];
    __shared__ uint8_t lists[TREE_THREADS / 32][64];
    const uint64_t nblocks = *n_blocks_ptr;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp_first = (((uint64_t)blockIdx.x * (TREE_THREADS / 32)) + (uint64_t)wib) * 64;
    if (warp_first >= nblocks) return;
    uint16_t *wrows = rows + (size_t)wib * 32 * TREE_STRIDE;
    const int todo = window_blocks(wk, warp_first, nblocks, lane, lists[wib]);
    for (int r0 = 0; r0 < todo; r0 += 32) {
        const bool live = r0 + lane < todo;
        const uint64_t t = warp_first + (live ? lists[wib][r0 + lane] : 0);
        const uint32_t info = live ? wk.binfo[t].x : 0u;
        const int used = (int)(info & 0xffffu);
        // stage the frequencies of the round's rows (coalesced)
        for (int rr = 0; rr < 32; ++rr) {
            const int u = __shfl_sync(FULL, used, rr);
            const uint64_t tr = __shfl_sync(FULL, t, rr);
            const uint32_t *src = wk.keys + tr * DEF_ROW;
            for (int i = lane; i < u; i += 32) wrows[rr * TREE_STRIDE + i] = (uint16_t)min(src[i] >> 9, 0xffffu);
        }
        __syncwarp();
        if (live && used >= 2) mk_lengths(wrows + lane * TREE_STRIDE, used, 15);
        __syncwarp();
        // lengths out, by symbol, and the block's token bits: sum f * (len + extra bits [+ the 1-bit distance code])
        uint32_t my_bits = 0;
        for (int rr = 0; rr < 32; ++rr) {
            const int u = __shfl_sync(FULL, used, rr);
            const uint64_t tr = __shfl_sync(FULL, t, rr);
            const uint32_t dist_len = (__shfl_sync(FULL, info, rr) >> 24) & 1u;
            const uint32_t *src = wk.keys + tr * DEF_ROW;
            uint8_t *dst = wk.lens + tr * DEF_ROW;
            uint32_t bits = 0;
            if (u) {
                for (int i = lane; i < DEF_ROW / 4; i += 32) reinterpret_cast<uint32_t *>(dst)[i] = 0;
                __syncwarp();
            }
            for (int i = lane; i < u; i += 32) {
                const uint32_t key = src[i], l = wrows[rr * TREE_STRIDE + i];
                const uint32_t s = key & 511u, f = key >> 9;
                dst[s] = (uint8_t)l;
                uint32_t xb = 0;
                if (s >= 265 && s < 285) xb = (s - 261) >> 2;
                bits += f * (l + xb + (s > 256 ? dist_len : 0u));
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) bits += __shfl_xor_sync(FULL, bits, d);
            if (lane == rr) my_bits = bits;
        }
        if (live) wk.binfo[t].y = my_bits;
        __syncwarp();
    }
}

// ---- kernel 2b: one thread per block: canonical codes and the dynamic block header -------------------------------------
// What a warp did for one block at a time inside the emit kernel (canonical codes with match/ballot rounds, run-length
// coding of the code lengths, the 19-symbol code-length code built by lane 0, header bits placed by prefix scans: 4 400 warp
// instructions per block, more than coding the block's 4 KiB of literals) is serial work on ~300 small numbers: here every
// thread does it for its own block.  Results: tab[] (what the token loops look up) and the header bits, both in the
// workspace.  The symbol sequence and the code-length code are the ones the warp version produced (same greedy run
// splitting as zlib's send_tree, same two-queue merge and tie-breaking), so the streams did not change.
struct HdrBits {
    uint32_t *out;
    uint64_t acc;
    uint32_t n, words;
    __device__ __forceinline__ void put(uint32_t bits, uint32_t nbits) {
        acc |= (uint64_t)bits << n;
        n += nbits;
        if (n >= 32) {
            out[words++] = (uint32_t)acc;
            acc >>= 32;
            n -= 32;
        }
    }
};

__device__ __forceinline__ void block_header(const DefWork &wk, uint64_t t);

__global__ void __launch_bounds__(128) deflate_header_kernel(DefWork wk, const uint64_t *n_blocks_ptr) {
    __shared__ uint8_t lists[4][64];
    const uint64_t nblocks = *n_blocks_ptr;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp_first = (((uint64_t)blockIdx.x * 4) + (uint64_t)wib) * 64;
    if (warp_first >= nblocks) return;
    const int todo = window_blocks(wk, warp_first, nblocks, lane, lists[wib]);
    for (int r0 = 0; r0 < todo; r0 += 32)
        if (r0 + lane < todo) block_header(wk, warp_first + lists[wib][r0 + lane]);
}

__device__ __forceinline__ void block_header(const DefWork &wk, const uint64_t t) {
    const uint4 info = wk.binfo[t];
    const uint32_t dist_len = (info.x >> 24) & 1u;
    const uint32_t *len4 = reinterpret_cast<const uint32_t *>(wk.lens + t * DEF_ROW);
    uint4 *tab4 = reinterpret_cast<uint4 *>(wk.tab + t * DEF_ROW);
    uint16_t *cl = reinterpret_cast<uint16_t *>(wk.keys + t * DEF_ROW);  // the sorted keys are done with: run-length symbols
    // ---- per-length counts, first code of every length (RFC 1951 3.2.2), number of literal/length codes
    uint32_t next[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) next[b] = 0;
    int hlit = 257;
    for (int w = 0; w < DEF_ROW / 4; ++w) {
        const uint32_t v = len4[w];
        if (v == 0) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t l = (v >> (8 * b)) & 0xffu;
            if (l) {
                ++next[l];
                if (4 * w + b >= 257) hlit = 4 * w + b + 1;
            }
        }
    }
    {
        uint32_t code = 0, prev = 0;
#pragma unroll
        for (int b = 1; b <= 15; ++b) {
            code = (code + prev) << 1;
            prev = next[b];
            next[b] = code;
        }
    }
    // ---- codes in symbol order; the run-length coded length sequence (hlit lengths + the one distance length) on the way
    uint32_t clfreq[19];
#pragma unroll
    for (int q = 0; q < 19; ++q) clfreq[q] = 0;
    int ncl = 0;
    auto emit = [&](uint32_t sym, uint32_t ext) {
        cl[ncl++] = (uint16_t)(sym | (ext << 8));
        ++clfreq[sym];
    };
    auto flush_run = [&](uint32_t v, uint32_t left) {
        if (v == 0) {
            for (; left >= 138; left -= 138) emit(18, 138 - 11);
            if (left >= 11) {
                emit(18, left - 11);
                left = 0;
            } else if (left >= 3) {
                emit(17, left - 3);
                left = 0;
            }
            for (; left; --left) emit(0, 0);
        } else {
            emit(v, 0);
            --left;
            for (; left >= 6; left -= 6) emit(16, 6 - 3);
            if (left >= 3) {
                emit(16, left - 3);
                left = 0;
            }
            for (; left; --left) emit(v, 0);
        }
    };
    uint32_t run_v = 0, run_n = 0;
    for (int w = 0; w < DEF_ROW / 4; ++w) {
        const uint32_t v = len4[w];
        if (v == 0 && run_n && run_v == 0 && 4 * w + 3 < hlit) {  // four more unused symbols inside a run of them (most words)
            run_n += 4;
            tab4[w] = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        uint32_t e[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t l = (v >> (8 * b)) & 0xffu;
            const int s = 4 * w + b;
            uint32_t cw = 0;
            if (l) cw = (__brev(next[l]++) >> (32 - l)) | (l << 16);
            e[b] = cw;
            if (s <= hlit) {  // entry hlit of the sequence is the distance length
                const uint32_t x = s < hlit ? l : dist_len;
                if (run_n && x == run_v) {
                    ++run_n;
                } else {
                    if (run_n) flush_run(run_v, run_n);
                    run_v = x;
                    run_n = 1;
                }
            }
        }
        tab4[w] = make_uint4(e[0], e[1], e[2], e[3]);
    }
    flush_run(run_v, run_n);
    // ---- the code-length code: 19 symbols, at most 7 bits (two-queue merge over the sorted frequencies, zlib's repair when a
    // depth exceeds the limit)
    uint8_t cllen[19];
    {
        int used = 0;
#pragma unroll
        for (int q = 0; q < 19; ++q) used += clfreq[q] != 0;
        for (int q = 0; q < 19 && used < 2; ++q)
            if (clfreq[q] == 0) {
                clfreq[q] = 1;
                ++used;
            }
        uint32_t key[19], weight[38];
        uint8_t parent[38];
        int n = 0;
        for (int q = 0; q < 19; ++q) {
            cllen[q] = 0;
            if (clfreq[q]) {  // insertion sort by (frequency, symbol)
                const uint32_t k = (clfreq[q] << 9) | (uint32_t)q;
                int i = n++;
                for (; i > 0 && key[i - 1] > k; --i) key[i] = key[i - 1];
                key[i] = k;
            }
        }
        for (int i = 0; i < n; ++i) weight[i] = key[i] >> 9;
        int li = 0, ii = n, nx = n;
        for (int j = 0; j < n - 1; ++j) {
            int pick[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (li < n && (ii >= nx || weight[li] <= weight[ii])) pick[k] = li++;
                else pick[k] = ii++;
            }
            weight[nx] = weight[pick[0]] + weight[pick[1]];
            parent[pick[0]] = (uint8_t)nx;
            parent[pick[1]] = (uint8_t)nx;
            ++nx;
        }
        const int root = 2 * n - 2, limit = 7;
        uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t kraft = 0;
        bool over = false;
        for (int i = 0; i < n; ++i) {
            int d = 0;
            for (int v = i; v != root; v = parent[v]) ++d;
            if (d > limit) {
                d = limit;
                over = true;
            }
            kraft += 1u << (limit - d);
            weight[i] = (uint32_t)d;
            ++cnt[d];
        }
        if (over) {
            int excess = (int)kraft - (1 << limit);
            while (excess > 0) {
                int bits = limit - 1;
                while (cnt[bits] == 0) --bits;
                --cnt[bits];
                cnt[bits + 1] += 2;
                --cnt[limit];
                --excess;
            }
            int i = 0;
            for (int bits = limit; bits >= 1; --bits)
                for (uint32_t c = cnt[bits]; c > 0; --c) weight[i++] = (uint32_t)bits;
        }
        for (int i = 0; i < n; ++i) cllen[key[i] & 511u] = (uint8_t)weight[i];
    }
    uint32_t clcode[19];
    {
        uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nxt[8];
        for (int q = 0; q < 19; ++q) ++cnt[cllen[q]];
        cnt[0] = 0;
        uint32_t code = 0;
        nxt[0] = 0;
        for (int b = 1; b <= 7; ++b) {
            code = (code + cnt[b - 1]) << 1;
            nxt[b] = code;
        }
        for (int q = 0; q < 19; ++q) {
            const uint32_t l = cllen[q];
            clcode[q] = l ? __brev(nxt[l]++) >> (32 - l) : 0u;
        }
    }
    int hclen = 19;
    while (hclen > 4 && cllen[c_def_cl_order[hclen - 1]] == 0) --hclen;
    // ---- the header bits (RFC 1951 3.2.7), after BFINAL / BTYPE
    HdrBits hb{wk.hdr + t * HDR_WORDS, 0ull, 0u, 0u};
    hb.put((uint32_t)(hlit - 257), 5);
    hb.put(0u, 5);  // one distance code
    hb.put((uint32_t)(hclen - 4), 4);
    for (int q = 0; q < hclen; ++q) hb.put(cllen[c_def_cl_order[q]], 3);
    for (int q = 0; q < ncl; ++q) {
        const uint32_t sym = cl[q] & 0xffu, ext = cl[q] >> 8;
        const uint32_t l = cllen[sym];
        hb.put(clcode[sym] | (ext << l), l + (sym == 16 ? 2u : sym == 17 ? 3u : sym == 18 ? 7u : 0u));
    }
    const uint32_t total = hb.words * 32 + hb.n;
    if (hb.n) hb.out[hb.words] = (uint32_t)hb.acc;
    wk.binfo[t].z = total;
}

// ---- kernel 3: emit ------------------------------------------------------------------------------------------------
struct DefTreeScratch {  // code-length code construction (19 symbols)
    uint32_t sortbuf[32];
    uint32_t weight[40];
    uint16_t parent[40];
    uint16_t run_start[DEF_ROW + 2];
};
struct __align__(128) EmitWarpSmem {
    uint32_t out[(DEF_OUT + DEF_OUT_SLACK) / 4];
    uint32_t tab[DEF_ROW];  // code (bit-reversed, LSB-first ready) | length << 16
    uint16_t code[DEF_ROW];
    uint8_t len[DEF_ROW];
    uint8_t clsym[320];  // run-length coded code lengths: symbol 0..18
    uint8_t clext[320];  // and its extra-bits value
    uint32_t clhist[32];
    uint16_t clcode[20];
    uint8_t cllen[20];
    alignas(4) uint16_t bl_count[16];
    DefTreeScratch k;
};

__device__ __forceinline__ void put64(const BitOut &bo, uint32_t pos, uint64_t bits, uint32_t nbits) {
    if (nbits == 0) return;
    const uint32_t w = pos >> 5, sh = pos & 31u;
    const uint32_t lo = (uint32_t)bits, hi = (uint32_t)(bits >> 32);
    atomicOr(&bo.buf[w], lo << sh);
    if (sh + nbits > 32) atomicOr(&bo.buf[w + 1], __funnelshift_l(lo, hi, sh));
    if (sh + nbits > 64) atomicOr(&bo.buf[w + 2], __funnelshift_l(hi, 0u, sh));
}

__global__ void __launch_bounds__(EMIT_WARPS * 32) deflate_emit_kernel(const DeflateArgs a, DefWork wk) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    EmitWarpSmem &ws = reinterpret_cast<EmitWarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint8_t *slab_end = a.in + a.in_capacity;
    for (;;) {
        unsigned long long r = 0;
        if (lane == 0) r = atomicAdd(a.work_counter, 1ULL);
        r = __shfl_sync(FULL, r, 0);
        if (r >= a.n_reads) break;
        if (a.status[r] != S5B_OK) continue;
        const uint8_t *src = a.in + a.in_off[r];
        const uint32_t ilen = a.in_len[r];
        uint8_t *dst = a.out + a.out_off[r];
        uint32_t split = a.split ? a.split[r] : 0;
        if (split >= ilen) split = 0;
        const uint32_t nb = def_nblocks(ilen, split);
        const uint64_t blk0 = wk.blk_off[r];

        BitOut bo;
        bo.buf = ws.out;
        bo.head = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
        bo.gbase = dst - bo.head;
        bo.bitpos = bo.head * 8;
        bo.written = 0;
        for (uint32_t i = lane; i < (DEF_OUT + DEF_OUT_SLACK) / 4; i += 32) ws.out[i] = 0;
        __syncwarp();
        if (lane == 0) bo.put(bo.bitpos, 0x9c78u, 16);  // CMF/FLG: deflate, 32 KiB window, default level (78 9C)
        bo.bitpos += 16;

        for (uint32_t k = 0; k < nb; ++k) {
            uint32_t b0, b1;
            def_block_range(ilen, split, k, b0, b1);
            const uint32_t blen = b1 - b0;
            const bool last = k + 1 == nb;
            const uint4 info = wk.binfo[blk0 + k];
            const uint32_t mode = (info.x >> 16) & 0xffu;
            const uint32_t dist_len = (info.x >> 24) & 1u;
            const uint32_t tok_bits = info.y;
            const bool canned = mode == MODE_CANNED;
            const uint8_t *order = c_def_cl_order;
            const int hdist = 1;
            int hlit = 286, ncl = 0, hclen = 19;
            uint32_t dyn_bits;
            // the first loads of the block are issued here, so that they are in flight while the header is being written
            WordReader rd;
            uint32_t run_next = 0x100u;
            if (mode == MODE_LIT) rd.start(src + b0, slab_end, lane);
            else run_next = (uint32_t)lane < blen ? (uint32_t)src[b0 + lane] : 0x100u;
            const bool prebuilt = !canned && !wk.hdr_by_warp;  // table and header bits come from the header kernel
            const uint32_t pre_bits = info.z;
            if (canned) {
                // the code fixed ahead of time: its table and header are constants
                for (int s = lane; s < DEF_ROW; s += 32) ws.tab[s] = g_canned_tab[s];
                dyn_bits = 0;  // (measured after coding)
                __syncwarp();
            } else if (prebuilt) {
                const uint32_t *tab = wk.tab + (blk0 + k) * DEF_ROW;
                for (int s = lane; s < DEF_ROW; s += 32) ws.tab[s] = tab[s];
                dyn_bits = tok_bits + 3 + pre_bits;
                __syncwarp();
            } else {
                // ---- the block's code lengths, in symbol order
                {
                    const uint8_t *lens = wk.lens + (blk0 + k) * DEF_ROW;
                    for (int s = lane; s < DEF_ROW; s += 32) ws.len[s] = lens[s];
                }
                __syncwarp();
                canonical_codes(ws.len, 286, ws.code, ws.bl_count, lane);
                for (int s = lane; s < DEF_ROW; s += 32) ws.tab[s] = (uint32_t)ws.code[s] | ((uint32_t)ws.len[s] << 16);
                hlit = 286;
                while (hlit > 257 && ws.len[hlit - 1] == 0) --hlit;
                // ---- header: run-length code the hlit + hdist code lengths (RFC 1951 3.2.7), whole warp: find the runs of
                // equal lengths, turn every run into its symbols (16: repeat previous 3-6, 17: zeros 3-10, 18: zeros
                // 11-138, greedy like zlib's send_tree), place them with a prefix scan over the runs
                ncl = 0;
                if (lane < 19) ws.clhist[lane] = 0;
                __syncwarp();
                {
                    const int total = hlit + hdist;
                    auto length_at = [&](int q) -> uint32_t { return q < hlit ? ws.len[q] : dist_len; };
                    uint16_t *run_start = ws.k.run_start;
                    int nruns = 0;
                    for (int k0 = 0; k0 < total; k0 += 32) {
                        const int q = k0 + lane;
                        const bool st = q < total && (q == 0 || length_at(q) != length_at(q - 1));
                        const uint32_t m = __ballot_sync(FULL, st);
                        if (st) run_start[nruns + __popc(m & ((1u << lane) - 1u))] = (uint16_t)q;
                        nruns += __popc(m);
                    }
                    __syncwarp();
                    for (int r0 = 0; r0 < nruns; r0 += 32) {
                        const int rr = r0 + lane;
                        uint32_t v = 0, n18 = 0, n17 = 0, n16 = 0, nlit = 0, tail = 0;
                        // n18 / n16: full-size repeat symbols; tail: size of one more, smaller repeat symbol (0 = none)
                        if (rr < nruns) {
                            const int s0 = run_start[rr];
                            const int s1 = rr + 1 < nruns ? run_start[rr + 1] : total;
                            uint32_t left = (uint32_t)(s1 - s0);
                            v = length_at(s0);
                            if (v == 0) {
                                n18 = left / 138u;
                                left -= n18 * 138u;
                                if (left >= 11) {
                                    tail = left;
                                    ++n18;
                                    left = 0;
                                } else if (left >= 3) {
                                    tail = left;
                                    n17 = 1;
                                    left = 0;
                                }
                                nlit = left;
                            } else {
                                nlit = 1;  // the value itself first, repeats refer back to it
                                --left;
                                n16 = left / 6u;
                                left -= n16 * 6u;
                                if (left >= 3) {
                                    tail = left;
                                    ++n16;
                                    left = 0;
                                }
                                nlit += left;
                            }
                        }
                        const uint32_t cnt = n18 + n17 + n16 + nlit;
                        uint32_t incl = cnt;
    #pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t t = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += t;
                        }
                        int at = ncl + (int)(incl - cnt);
                        if (rr < nruns) {
                            if (v == 0) {
                                const uint32_t full18 = tail >= 11 ? n18 - 1 : n18;
                                for (uint32_t i = 0; i < full18; ++i) {
                                    ws.clsym[at] = 18;
                                    ws.clext[at++] = 138 - 11;
                                }
                                if (tail >= 11) {
                                    ws.clsym[at] = 18;
                                    ws.clext[at++] = (uint8_t)(tail - 11);
                                } else if (tail >= 3) {
                                    ws.clsym[at] = 17;
                                    ws.clext[at++] = (uint8_t)(tail - 3);
                                }
                                if (n18) atomicAdd(&ws.clhist[18], n18);
                                if (n17) atomicAdd(&ws.clhist[17], n17);
                                for (uint32_t i = 0; i < nlit; ++i) {
                                    ws.clsym[at] = 0;
                                    ws.clext[at++] = 0;
                                }
                                if (nlit) atomicAdd(&ws.clhist[0], nlit);
                            } else {
                                ws.clsym[at] = (uint8_t)v;
                                ws.clext[at++] = 0;
                                const uint32_t full16 = tail ? n16 - 1 : n16;
                                for (uint32_t i = 0; i < full16; ++i) {
                                    ws.clsym[at] = 16;
                                    ws.clext[at++] = 6 - 3;
                                }
                                if (tail) {
                                    ws.clsym[at] = 16;
                                    ws.clext[at++] = (uint8_t)(tail - 3);
                                }
                                if (n16) atomicAdd(&ws.clhist[16], n16);
                                for (uint32_t i = 1; i < nlit; ++i) {
                                    ws.clsym[at] = (uint8_t)v;
                                    ws.clext[at++] = 0;
                                }
                                atomicAdd(&ws.clhist[v], nlit);
                            }
                        }
                        ncl += (int)__shfl_sync(FULL, incl, 31);
                    }
                }
                __syncwarp();
                huffman_lengths(ws.clhist, 19, 7, ws.cllen, ws.k.sortbuf, ws.k.weight, ws.k.parent, ws.bl_count, lane);
                canonical_codes(ws.cllen, 19, ws.clcode, ws.bl_count, lane);
                hclen = 19;
                while (hclen > 4 && ws.cllen[order[hclen - 1]] == 0) --hclen;

                // ---- size of the dynamic block in bits (header + tokens + end of block): stored instead if that is smaller
                uint32_t hdr_bits = 0;
                for (int q = lane; q < ncl; q += 32) {
                    const uint32_t s = ws.clsym[q];
                    hdr_bits += ws.cllen[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0);
                }
    #pragma unroll
                for (int d = 16; d; d >>= 1) hdr_bits += __shfl_xor_sync(FULL, hdr_bits, d);
                dyn_bits = tok_bits + hdr_bits + 3 + 5 + 5 + 4 + 3 * hclen;
            }
            const uint32_t stored_bits = 8u * blen + 40u;

            auto emit_stored = [&]() {
                // ---- stored block: pad to a byte boundary, LEN, NLEN, raw bytes
                if ((bo.bitpos >> 3) + 16 > DEF_OUT) bo.flush(lane, false);
                if (lane == 0) bo.put(bo.bitpos, last ? 1u : 0u, 3);
                bo.bitpos = (bo.bitpos + 3 + 7) & ~7u;
                if (lane == 0) {
                    bo.put(bo.bitpos, blen, 16);
                    bo.put(bo.bitpos + 16, blen ^ 0xffffu, 16);
                }
                bo.bitpos += 32;
                for (uint32_t t0 = 0; t0 < blen; t0 += 32) {
                    if ((bo.bitpos >> 3) + 40 > DEF_OUT) bo.flush(lane, false);
                    const uint32_t i = t0 + lane;
                    if (i < blen) bo.put(bo.bitpos + 8 * lane, src[b0 + i], 8);
                    bo.bitpos += 8 * min(32u, blen - t0);
                }
            };
            uint32_t canned_start = 0;
            if (!canned && dyn_bits >= stored_bits + 7u) {
                emit_stored();
            } else {
                if (canned) {
                    // ---- BFINAL / BTYPE and the canned header, a word per lane.  Room first for the whole block (header, at
                    // most 12 bits per byte, end of block): it is measured after it has been coded, and taking it back only
                    // works while nothing of it has left the buffer
                    if ((bo.bitpos >> 3) + (blen * 3u) / 2u + 160u > DEF_OUT) bo.flush(lane, false);
                    canned_start = bo.bitpos;
                    if (lane == 0) bo.put(bo.bitpos, (last ? 1u : 0u) | (2u << 1), 3);
                    bo.bitpos += 3;
                    if (lane < CANNED_HDR_WORDS)
                        bo.put(bo.bitpos + 32u * lane, g_canned_hdr[lane], min(32u, CANNED_HDR_BITS - 32u * lane));
                    bo.bitpos += CANNED_HDR_BITS;
                } else if (prebuilt) {
                    // ---- BFINAL / BTYPE and the header bits the header kernel left in the workspace, a word per lane
                    if ((bo.bitpos >> 3) + HDR_WORDS * 4 + 16 > DEF_OUT) bo.flush(lane, false);
                    if (lane == 0) bo.put(bo.bitpos, (last ? 1u : 0u) | (2u << 1), 3);
                    bo.bitpos += 3;
                    const uint32_t *hdr = wk.hdr + (blk0 + k) * HDR_WORDS;
                    for (uint32_t w = lane; w * 32 < pre_bits; w += 32) bo.put(bo.bitpos + 32u * w, hdr[w], min(32u, pre_bits - 32u * w));
                    bo.bitpos += pre_bits;
                } else {
                    // ---- dynamic block header (lane 0; a few hundred bits)
                    if ((bo.bitpos >> 3) + 24 > DEF_OUT) bo.flush(lane, false);
                    if (lane == 0) {
                        uint32_t p = bo.bitpos;
                        bo.put(p, (last ? 1u : 0u) | (2u << 1), 3);
                        p += 3;
                        bo.put(p, (uint32_t)(hlit - 257), 5);
                        p += 5;
                        bo.put(p, (uint32_t)(hdist - 1), 5);
                        p += 5;
                        bo.put(p, (uint32_t)(hclen - 4), 4);
                        p += 4;
                        for (int q = 0; q < hclen; ++q) {
                            bo.put(p, ws.cllen[order[q]], 3);
                            p += 3;
                        }
                        bo.bitpos = p;
                    }
                    bo.bitpos = __shfl_sync(FULL, bo.bitpos, 0);
                    for (int k0 = 0; k0 < ncl; k0 += 32) {
                        if ((bo.bitpos >> 3) + 64 > DEF_OUT) bo.flush(lane, false);
                        const int q = k0 + lane;
                        uint32_t bits = 0, nbits = 0;
                        if (q < ncl) {
                            const uint32_t s = ws.clsym[q];
                            nbits = ws.cllen[s];
                            bits = ws.clcode[s];
                            const uint32_t xb = s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0;
                            bits |= (uint32_t)ws.clext[q] << nbits;
                            nbits += xb;
                        }
                        uint32_t incl = nbits;
    #pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t t = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += t;
                        }
                        bo.put(bo.bitpos + incl - nbits, bits, nbits);
                        bo.bitpos += __shfl_sync(FULL, incl, 31);
                    }
                }
                // ---- tokens.  Same tokenisation as the counting kernel; a warp prefix scan over the token bit counts
                // places every lane's bits.
                if (mode == MODE_LIT) {
                    const uint32_t nstrip = (blen + 127) >> 7;
                    for (uint32_t t = 0; t < nstrip; ++t) {
                        if ((bo.bitpos >> 3) + 256 > DEF_OUT) bo.flush(lane, false);
                        const uint32_t v = rd.next(t, lane);
                        const uint32_t pos = (t << 7) + 4u * lane;
                        const uint32_t nv = pos >= blen ? 0u : min(4u, blen - pos);
                        uint32_t e0 = ws.tab[v & 255u], e1 = ws.tab[(v >> 8) & 255u], e2 = ws.tab[(v >> 16) & 255u],
                                 e3 = ws.tab[v >> 24];
                        if (nv < 4) {
                            if (nv < 1) e0 = 0;
                            if (nv < 2) e1 = 0;
                            if (nv < 3) e2 = 0;
                            e3 = 0;
                        }
                        const uint32_t l0 = e0 >> 16, l1 = e1 >> 16, l2 = e2 >> 16, l3 = e3 >> 16;
                        // two halves of at most 30 bits each, then one 64-bit word
                        const uint32_t lo = (e0 & 0xffffu) | ((e1 & 0xffffu) << l0);
                        const uint32_t hi = (e2 & 0xffffu) | ((e3 & 0xffffu) << l2);
                        const uint32_t nlo = l0 + l1, nbits = nlo + l2 + l3;
                        const uint64_t acc = (uint64_t)lo | ((uint64_t)hi << nlo);
                        uint32_t incl = nbits;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t q = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += q;
                        }
                        put64(bo, bo.bitpos + incl - nbits, acc, nbits);
                        bo.bitpos += __shfl_sync(FULL, incl, 31);
                    }
                } else {
                    uint32_t prev0 = b0 ? (uint32_t)src[b0 - 1] : 0x200u;
                    for (uint32_t t0 = 0; t0 < blen; t0 += 32) {
                        if ((bo.bitpos >> 3) + 80 > DEF_OUT) bo.flush(lane, false);
                        const uint32_t b = run_next;  // (the strip after this one is requested now: one iteration to land)
                        const uint32_t i2 = t0 + 32 + lane;
                        run_next = i2 < blen ? (uint32_t)src[b0 + i2] : 0x100u;
                        const Token tk = strip_token(b, prev0, lane);
                        prev0 = __shfl_sync(FULL, b, 31);
                        uint32_t bits = 0, nbits = 0;
                        if (tk.kind == 1) {
                            const uint32_t e = ws.tab[b];
                            bits = e & 0xffffu;
                            nbits = e >> 16;
                        } else if (tk.kind == 2) {
                            uint32_t sym, xb, xv;
                            len_code(tk.mlen, sym, xb, xv);
                            const uint32_t e = ws.tab[257 + sym];
                            nbits = e >> 16;
                            bits = (e & 0xffffu) | (xv << nbits);
                            nbits += xb;
                            nbits += dist_len;  // distance code 0 (distance 1): its one-bit canonical code is 0
                        }
                        uint32_t incl = nbits;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t q = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += q;
                        }
                        bo.put(bo.bitpos + incl - nbits, bits, nbits);
                        bo.bitpos += __shfl_sync(FULL, incl, 31);
                    }
                }
                if ((bo.bitpos >> 3) + 8 > DEF_OUT) bo.flush(lane, false);
                if (lane == 0) bo.put(bo.bitpos, ws.tab[256] & 0xffffu, ws.tab[256] >> 16);
                bo.bitpos += ws.tab[256] >> 16;
                if (canned && bo.bitpos - canned_start >= stored_bits + 7u) {
                    // the canned code did not shrink this block (not the kind of bytes it was made for): take the bits back
                    // and store the block
                    __syncwarp();
                    const uint32_t w0 = canned_start >> 5, w1 = bo.bitpos >> 5;
                    if (lane == 0) ws.out[w0] &= (1u << (canned_start & 31u)) - 1u;
                    for (uint32_t w = w0 + 1 + lane; w <= w1; w += 32) ws.out[w] = 0;
                    __syncwarp();
                    bo.bitpos = canned_start;
                    emit_stored();
                }
            }
            __syncwarp();
        }
        // ---- Adler-32 trailer, big endian, byte aligned
        bo.bitpos = (bo.bitpos + 7) & ~7u;
        if ((bo.bitpos >> 3) + 8 > DEF_OUT) bo.flush(lane, false);
        if (lane == 0) bo.put(bo.bitpos, __byte_perm(wk.adler[r], 0, 0x0123), 32);
        bo.bitpos += 32;
        bo.flush(lane, true);
        if (lane == 0) a.out_len[r] = (uint32_t)bo.written;
        __syncwarp();
    }
}

struct DefOcc {
    int cnt = 0, emit = 0;
    bool ready = false;
};
static DefOcc &def_occ() {
    static DefOcc o;
    if (!o.ready) {
        if (cudaFuncSetAttribute(deflate_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(EmitWarpSmem) * EMIT_WARPS)) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o.emit, deflate_emit_kernel, EMIT_WARPS * 32,
                                                          sizeof(EmitWarpSmem) * EMIT_WARPS) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o.cnt, deflate_count_kernel, CNT_WARPS * 32, 0) == cudaSuccess)
            o.ready = o.cnt > 0 && o.emit > 0;
    }
    return o;
}

}  // namespace defl
using namespace defl;

int deflate_blocks_per_sm() { return def_occ().ready ? def_occ().emit : 0; }

uint64_t deflate_bound(uint64_t len) { return len + 6 * (len / DEF_BLOCK + 2) + 8; }

// a record of len bytes has at most len / DEF_BLOCK + 2 blocks (one rounding up on each side of the split point)
static uint64_t def_max_blocks(uint64_t in_capacity, uint64_t n_reads) { return in_capacity / DEF_BLOCK + 2 * n_reads + 1; }

size_t deflate_work_bytes(uint64_t in_capacity, uint64_t n_reads) {
    const uint64_t mb = def_max_blocks(in_capacity, n_reads);
    return (size_t)(n_reads * 4 + (n_reads + 1) * 8 + n_reads * 4 + mb * 16 + mb * DEF_ROW * 9 + mb * HDR_WORDS * 4 +
                    compact_scratch_bytes(n_reads) + 512);
}

cudaError_t launch_deflate(const DeflateArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    (void)blocks_per_sm;
    if (a.n_reads == 0) return cudaSuccess;
    if (!a.work || a.work_bytes < deflate_work_bytes(a.in_capacity, a.n_reads)) return cudaErrorInvalidValue;
    DefOcc &occ = def_occ();
    if (!occ.ready) return cudaErrorInvalidDeviceFunction;
    // carve the workspace (16-byte aligned pieces)
    DefWork w;
    const uint64_t n = a.n_reads, mb = def_max_blocks(a.in_capacity, n);
    uint8_t *p = static_cast<uint8_t *>(a.work);
    auto take = [&](size_t bytes) {
        uint8_t *q = p;
        p += (bytes + 15) & ~size_t(15);
        return q;
    };
    w.binfo = reinterpret_cast<uint4 *>(take(mb * 16));
    w.blk_off = reinterpret_cast<uint64_t *>(take((n + 1) * 8));
    w.keys = reinterpret_cast<uint32_t *>(take(mb * DEF_ROW * 4));
    w.nblk = reinterpret_cast<uint32_t *>(take(n * 4));
    w.adler = reinterpret_cast<uint32_t *>(take(n * 4));
    w.lens = take(mb * DEF_ROW);
    w.tab = reinterpret_cast<uint32_t *>(take(mb * DEF_ROW * 4));
    w.hdr = reinterpret_cast<uint32_t *>(take(mb * HDR_WORDS * 4));
    w.scan_scratch = take(compact_scratch_bytes(n));
    w.max_blocks = mb;
    {
        static const bool off = [] {
            const char *e = getenv("S5B_DEFLATE_CANNED");
            return e && e[0] == '0';
        }();
        w.canned = off ? 0u : 1u;
        static const bool by_warp = [] {
            const char *e = getenv("S5B_DEFLATE_HDR");
            return e && e[0] == 'w';
        }();
        w.hdr_by_warp = by_warp ? 1u : 0u;
    }
    deflate_plan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, w);
    cudaError_t e = launch_scan(w.nblk, n, 1, w.blk_off, w.scan_scratch, st);
    if (e != cudaSuccess) return e;
    auto grid_for = [&](int warps, int bps) {
        uint64_t want = (n + warps - 1) / warps;
        uint64_t cap = (uint64_t)num_sms * (bps > 0 ? bps : 1);
        unsigned g = (unsigned)(want < cap ? want : cap);
        return g ? g : 1u;
    };
    e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    deflate_count_kernel<<<grid_for(CNT_WARPS, occ.cnt), CNT_WARPS * 32, 0, st>>>(a, w);
    // one thread per block; the block count lives on the device (blk_off[n]): the grid covers the bound, idle warps leave
    // (a warp takes a window of 64 blocks)
    deflate_tree_kernel<<<(unsigned)((mb + 2 * TREE_THREADS - 1) / (2 * TREE_THREADS)), TREE_THREADS, 0, st>>>(w, w.blk_off + n);
    if (!w.hdr_by_warp) deflate_header_kernel<<<(unsigned)((mb + 255) / 256), 128, 0, st>>>(w, w.blk_off + n);
    e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    deflate_emit_kernel<<<grid_for(EMIT_WARPS, occ.emit), EMIT_WARPS * 32, sizeof(EmitWarpSmem) * EMIT_WARPS, st>>>(a, w);
    return cudaGetLastError();
}

}  // namespace s5b
