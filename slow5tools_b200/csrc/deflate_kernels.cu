// deflate_kernels.cu -- sm_100a zlib/DEFLATE encoder, one warp per record.
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
//   * the record is cut into blocks (a caller-supplied split point -- the boundary between the
//     header+key bytes and the data bytes of the svb stream -- then every DEF_BLOCK bytes); each block
//     gets its own dynamic Huffman code;
//   * tokens are literals plus distance-1 run matches found with warp ballots inside 32-byte strips
//     (one byte per lane): no hash chains, no serial match search;
//   * per block: byte histogram by shared-memory atomics -> warp bitonic sort -> two-queue Huffman
//     construction -> 15-bit length limiting -> canonical codes -> run-length coded header (RFC 1951
//     3.2.7) with its own 7-bit-limited code -> token bits placed by a warp prefix scan over the code
//     lengths and OR-ed into a shared-memory bit buffer that leaves as 128-bit coalesced stores;
//   * a block that would not shrink is emitted as stored blocks, like zlib does.
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "huff_common.cuh"
#include "../../include/slow5b200.h"

namespace s5b {

namespace {

constexpr int DEF_WARPS = 4;
constexpr int DEF_BLOCK = 6144;  // max input bytes per deflate block (multiple of 32)
constexpr int DEF_OUT = HC_OUT;
constexpr int DEF_OUT_SLACK = HC_OUT_SLACK;
constexpr uint32_t ADLER_MOD = 65521u;

struct DefTreeScratch {  // live only while a block's codes are being constructed
    uint32_t sortbuf[512];
    uint32_t weight[576];
    uint16_t parent[576];
};
struct __align__(128) DefWarpSmem {
    // The staged block and the code-construction scratch share their bytes: the block is staged, counted,
    // overwritten by the scratch, and staged again (from L2 this time) for the emit pass.  Halves the shared
    // memory per warp, i.e. doubles the resident warps of this latency-bound kernel.
    union {
        uint8_t in[16 + DEF_BLOCK + 16];  // block staged from the 16-byte granule below its first byte
        DefTreeScratch k;
    };
    uint32_t out[(DEF_OUT + DEF_OUT_SLACK) / 4];
    uint32_t hist[288];
    uint16_t code[288];  // bit-reversed canonical codes (LSB-first ready)
    uint8_t len[288];
    uint8_t dlen[32];
    uint8_t clsym[320];  // run-length coded code lengths: symbol 0..18
    uint8_t clext[320];  // and its extra-bits value
    uint32_t clhist[32];
    uint16_t clcode[20];
    uint8_t cllen[20];
    alignas(4) uint16_t bl_count[16];
    unsigned long long bar;
};

__constant__ uint8_t c_def_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

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

// Adler-32 over n staged bytes (whole warp)
__device__ __forceinline__ void adler_update(uint32_t &a, uint32_t &b, const uint8_t *p, uint32_t n, int lane) {
    const uint32_t per = (n + 31) / 32;
    const uint32_t start = min(n, per * lane), end = min(n, start + per);
    uint32_t s1 = 0, s2 = 0;
    for (uint32_t i = start; i < end; ++i) {
        const uint32_t d = p[i];
        s1 += d;
        s2 += (end - i) * d;
    }
    uint32_t contrib = (uint32_t)(((uint64_t)(n - end) * s1 + s2) % ADLER_MOD);
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        s1 += __shfl_xor_sync(FULL, s1, d);
        contrib += __shfl_xor_sync(FULL, contrib, d);
    }
    b = (uint32_t)((b + (uint64_t)n * a + contrib) % ADLER_MOD);
    a = (a + s1) % ADLER_MOD;
}

// Token of lane `lane` in the 32-byte strip starting at block offset `t0`: a literal, the start of a
// distance-1 run match of length m (3..32, entirely inside the strip), or nothing (covered by a match).
struct Token {
    uint32_t kind;  // 0 none, 1 literal, 2 match
    uint32_t byte;
    uint32_t mlen;
};
__device__ __forceinline__ Token strip_token(const uint8_t *blk, uint32_t t0, uint32_t blen, bool have_prev, int lane) {
    Token tk;
    const uint32_t i = t0 + lane;
    const bool valid = i < blen;
    const uint32_t b = valid ? blk[i] : 0x100u;
    uint32_t prev = __shfl_up_sync(FULL, b, 1);
    if (lane == 0) prev = (t0 > 0 || have_prev) ? blk[(int)t0 - 1] : 0x200u;  // blk[-1] is the byte before the block
    const bool eq = valid && b == prev;
    const uint32_t mask = __ballot_sync(FULL, eq);
    tk.byte = b;
    tk.kind = valid ? 1u : 0u;
    tk.mlen = 0;
    if (eq) {
        const uint32_t below = ~mask & ((1u << lane) - 1u);
        const uint32_t s = below ? 32u - __clz(below) : 0u;        // first lane of this run of equal bytes
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

}  // namespace

__global__ void __launch_bounds__(DEF_WARPS * 32) deflate_kernel(const DeflateArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    DefWarpSmem &ws = reinterpret_cast<DefWarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t bar = smem_u32(&ws.bar);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t phase = 0;

    for (;;) {
        unsigned long long r = 0;
        if (lane == 0) r = atomicAdd(a.work_counter, 1ULL);
        r = __shfl_sync(FULL, r, 0);
        if (r >= a.n_reads) break;
        const uint64_t ioff = a.in_off[r];
        const uint32_t ilen = a.in_len[r];
        const uint64_t ooff = a.out_off[r];
        const uint64_t ocap = a.out_off[r + 1] - ooff;
        // worst case: stored blocks (5 bytes each) + zlib header + Adler-32 + one byte of bit padding per block
        const uint64_t nblocks_max = (uint64_t)ilen / DEF_BLOCK + 2;
        if (ioff + ilen > a.in_capacity || ocap < (uint64_t)ilen + 6 * nblocks_max + 8) {
            if (lane == 0) {
                a.status[r] = ioff + ilen > a.in_capacity ? S5B_ERR_ARG : S5B_ERR_NOSPACE;
                a.out_len[r] = 0;
            }
            continue;
        }
        const uint8_t *src = a.in + ioff;
        uint8_t *dst = a.out + ooff;
        uint32_t split = a.split ? a.split[r] : 0;
        if (split >= ilen) split = 0;

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
        uint32_t ad_a = 1, ad_b = 0;

        uint32_t b0 = 0;
        do {  // at least one block, so an empty record still gets a (final) block
            uint32_t b1 = ilen;
            if (split > b0) b1 = split;
            if (b1 - b0 > DEF_BLOCK) b1 = b0 + DEF_BLOCK;
            const uint32_t blen = b1 - b0;
            const bool last = b1 == ilen;
            // ---- stage [b0 - 1, b1) : the byte before the block is wanted for distance-1 matches
            const uint8_t *g0 = src + b0 - (b0 ? 1 : 0);
            const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(g0) & 15u);
            const uint8_t *g16 = g0 - skew;
            uint64_t bytes = ((uint64_t)skew + blen + (b0 ? 1 : 0) + 15) & ~15ull;
            const uint64_t room = a.in_capacity - (uint64_t)(g16 - a.in);
            if (bytes > room) bytes = room & ~15ull;
            auto stage_block = [&]() {
                if (!bytes) return;
                fence_proxy_async_smem();  // every lane's earlier generic accesses to these bytes come first
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(bar, (uint32_t)bytes);
                    bulk_g2s(smem_u32(ws.in), g16, (uint32_t)bytes, bar);
                }
                mbar_wait(bar, phase);
                phase ^= 1u;
            };
            stage_block();
            const uint8_t *blk = ws.in + skew + (b0 ? 1 : 0);  // blk[0] = first byte of the block, blk[-1] valid if b0 > 0
            const bool have_prev = b0 > 0;
            adler_update(ad_a, ad_b, blk, blen, lane);

            // ---- pass 1: histogram of literal/length symbols; count matches
            for (int s = lane; s < 288; s += 32) ws.hist[s] = 0;
            __syncwarp();
            uint32_t nmatch = 0;
            for (uint32_t t0 = 0; t0 < blen; t0 += 32) {
                const Token tk = strip_token(blk, t0, blen, have_prev, lane);
                if (tk.kind == 1) {
                    atomicAdd(&ws.hist[tk.byte], 1u);
                } else if (tk.kind == 2) {
                    uint32_t sym, xb, xv;
                    len_code(tk.mlen, sym, xb, xv);
                    atomicAdd(&ws.hist[257 + sym], 1u);
                    ++nmatch;
                }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) nmatch += __shfl_xor_sync(FULL, nmatch, d);
            if (lane == 0) ws.hist[256] = 1;  // end of block
            __syncwarp();

            // ---- code construction
            huffman_lengths(ws.hist, 286, 15, ws.len, ws.k.sortbuf, ws.k.weight, ws.k.parent, ws.bl_count, lane);
            canonical_codes(ws.len, 286, ws.code, ws.bl_count, lane);
            // distance alphabet: only distance 1 (code 0) is ever used.  One code of one bit when there are
            // matches, one code of zero bits when the block is all literals (RFC 1951 3.2.7).
            const uint32_t dist_len = nmatch ? 1u : 0u;
            int hlit = 286;
            while (hlit > 257 && ws.len[hlit - 1] == 0) --hlit;
            const int hdist = 1;
            // ---- header: run-length code the hlit + hdist code lengths (lane 0), then its 19-symbol code
            int ncl = 0;
            if (lane < 19) ws.clhist[lane] = 0;
            __syncwarp();
            {
                // run-length coding of the hlit + hdist code lengths (RFC 1951 3.2.7), whole warp: find the runs of
                // equal lengths, turn every run into its symbols (16: repeat previous 3-6, 17: zeros 3-10,
                // 18: zeros 11-138, greedy like zlib's send_tree), place them with a prefix scan over the runs
                const int total = hlit + hdist;
                auto length_at = [&](int k) -> uint32_t { return k < hlit ? ws.len[k] : dist_len; };
                uint16_t *run_start = ws.k.parent;  // scratch: free once the code lengths exist
                int nruns = 0;
                for (int k0 = 0; k0 < total; k0 += 32) {
                    const int k = k0 + lane;
                    const bool st = k < total && (k == 0 || length_at(k) != length_at(k - 1));
                    const uint32_t m = __ballot_sync(FULL, st);
                    if (st) run_start[nruns + __popc(m & ((1u << lane) - 1u))] = (uint16_t)k;
                    nruns += __popc(m);
                }
                __syncwarp();
                for (int r0 = 0; r0 < nruns; r0 += 32) {
                    const int r = r0 + lane;
                    uint32_t v = 0, n18 = 0, n17 = 0, n16 = 0, nlit = 0, tail = 0;
                    // n18 / n16: full-size repeat symbols; tail: size of one more, smaller repeat symbol (0 = none)
                    if (r < nruns) {
                        const int s0 = run_start[r];
                        const int s1 = r + 1 < nruns ? run_start[r + 1] : total;
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
                    if (r < nruns) {
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
            const uint8_t *order = c_def_cl_order;
            int hclen = 19;
            while (hclen > 4 && ws.cllen[order[hclen - 1]] == 0) --hclen;

            // ---- size of the dynamic block in bits (header + tokens + end of block)
            uint32_t hdr_bits = 0;
            for (int k = lane; k < ncl; k += 32) {
                const uint32_t s = ws.clsym[k];
                hdr_bits += ws.cllen[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0);
            }
            uint32_t tok_bits = 0;
            for (int s = lane; s < 286; s += 32) {
                uint32_t xb = 0;
                if (s >= 265 && s < 285) xb = (uint32_t)(s - 261) >> 2;
                tok_bits += ws.hist[s] * (ws.len[s] + xb + (s > 256 ? dist_len : 0));
            }
            // (hist[] was bumped for dummy symbols by huffman_lengths: they have no tokens, but the estimate only
            //  has to be an upper bound for the stored-block decision; the exact positions come from the scan below)
            uint32_t dyn_bits = tok_bits + hdr_bits;
#pragma unroll
            for (int d = 16; d; d >>= 1) dyn_bits += __shfl_xor_sync(FULL, dyn_bits, d);
            dyn_bits += 3 + 5 + 5 + 4 + 3 * hclen;
            const uint32_t stored_bits = 8u * blen + 40u;
            stage_block();  // the construction scratch overwrote the block

            if (dyn_bits >= stored_bits + 7u) {
                // ---- stored block: pad to a byte boundary, LEN, NLEN, raw bytes
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
                    if (i < blen) bo.put(bo.bitpos + 8 * lane, blk[i], 8);
                    bo.bitpos += 8 * min(32u, blen - t0);
                }
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
                    for (int k = 0; k < hclen; ++k) {
                        bo.put(p, ws.cllen[order[k]], 3);
                        p += 3;
                    }
                    bo.bitpos = p;
                }
                bo.bitpos = __shfl_sync(FULL, bo.bitpos, 0);
                for (int k0 = 0; k0 < ncl; k0 += 32) {
                    if ((bo.bitpos >> 3) + 64 > DEF_OUT) bo.flush(lane, false);
                    const int k = k0 + lane;
                    uint32_t bits = 0, nb = 0;
                    if (k < ncl) {
                        const uint32_t s = ws.clsym[k];
                        nb = ws.cllen[s];
                        bits = ws.clcode[s];
                        const uint32_t xb = s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0;
                        bits |= (uint32_t)ws.clext[k] << nb;
                        nb += xb;
                    }
                    uint32_t incl = nb;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += t;
                    }
                    bo.put(bo.bitpos + incl - nb, bits, nb);
                    bo.bitpos += __shfl_sync(FULL, incl, 31);
                }
                // ---- pass 2: tokens.  Same tokenisation as pass 1; a warp prefix scan over the token bit counts
                // places every lane's bits.
                for (uint32_t t0 = 0; t0 < blen; t0 += 32) {
                    if ((bo.bitpos >> 3) + 80 > DEF_OUT) bo.flush(lane, false);
                    const Token tk = strip_token(blk, t0, blen, have_prev, lane);
                    uint32_t bits = 0, nb = 0;
                    if (tk.kind == 1) {
                        bits = ws.code[tk.byte];
                        nb = ws.len[tk.byte];
                    } else if (tk.kind == 2) {
                        uint32_t sym, xb, xv;
                        len_code(tk.mlen, sym, xb, xv);
                        nb = ws.len[257 + sym];
                        bits = ws.code[257 + sym] | (xv << nb);
                        nb += xb;
                        // distance code 0 (distance 1): its one-bit canonical code is 0
                        nb += dist_len;
                    }
                    uint32_t incl = nb;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += t;
                    }
                    bo.put(bo.bitpos + incl - nb, bits, nb);
                    bo.bitpos += __shfl_sync(FULL, incl, 31);
                }
                if (lane == 0) bo.put(bo.bitpos, ws.code[256], ws.len[256]);
                bo.bitpos += ws.len[256];
            }
            b0 = b1;
            __syncwarp();
        } while (b0 < ilen);
        // ---- Adler-32 trailer, big endian, byte aligned
        bo.bitpos = (bo.bitpos + 7) & ~7u;
        if ((bo.bitpos >> 3) + 8 > DEF_OUT) bo.flush(lane, false);
        if (lane == 0) {
            const uint32_t ad = (ad_b << 16) | ad_a;
            bo.put(bo.bitpos, __byte_perm(ad, 0, 0x0123), 32);
        }
        bo.bitpos += 32;
        bo.flush(lane, true);
        if (lane == 0) {
            a.out_len[r] = (uint32_t)bo.written;
            a.status[r] = S5B_OK;
        }
        __syncwarp();
    }
}

int deflate_blocks_per_sm() {
    int n = 0;
    if (cudaFuncSetAttribute(deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(DefWarpSmem) * DEF_WARPS)) != cudaSuccess)
        return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, deflate_kernel, DEF_WARPS * 32,
                                                      sizeof(DefWarpSmem) * DEF_WARPS) != cudaSuccess)
        return 0;
    return n;
}

uint64_t deflate_bound(uint64_t len) { return len + 6 * (len / DEF_BLOCK + 2) + 8; }

cudaError_t launch_deflate(const DeflateArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t want = (a.n_reads + DEF_WARPS - 1) / DEF_WARPS;
    uint64_t cap = (uint64_t)num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (!grid) grid = 1;
    deflate_kernel<<<grid, DEF_WARPS * 32, sizeof(DefWarpSmem) * DEF_WARPS, st>>>(a);
    return cudaGetLastError();
}

}  // namespace s5b
