// zstd_encode_kernels.cu -- sm_100a Zstandard frame encoder, one warp per record.
//
// Replaces ptr_compress_zstd (slow5lib/src/slow5_press.c:1183-1202: ZSTD_compress(.., level 1) from system
// libzstd) for whole batches: every record becomes one single-segment frame that carries its content size (the
// reference's decoder refuses frames that do not, :1206-1211).  The compressed bytes of libzstd are not pinned by
// the reference (test/test_view.sh:204-214); the contract is that libzstd / the reference binary regenerate the
// exact input, and that the size stays within the tolerance stated in tests/test_zstd_encode_gpu.py.
//
// Coding choice: BLOW5 records are svb-zd streams -- small, noisy integers whose redundancy is almost entirely in
// the byte distribution, not in repeats (libzstd level 1 finds next to no matches in them).  A block is therefore
// written as one Huffman-coded literals section (4 interleaved streams, max 11 bits, FSE-compressed weights) and
// an empty sequences section; RLE and raw blocks cover the degenerate cases.  Blocks are cut at the caller's
// split hint (header+keys | data bytes, whose statistics differ) and every ZE_BLOCK bytes; a block whose
// statistics fit the previous table reuses it (treeless literals).
//
//   * record staged HBM -> smem by 1-D bulk async copies (UBLKCP) block by block;
//   * histogram with match-aggregated shared-memory atomics, warp bitonic sort + two-queue tree + Kraft-exact
//     length limiting (huff_common.cuh), zstd code numbering, tree description by lane 0 (zstd_enc_core.h);
//   * the four streams are written back to front: lane l takes symbol end-1-l of a 32-symbol strip, a warp
//     prefix scan over the code lengths places its bits in the shared bit buffer, which drains in 128-bit stores.
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "huff_common.cuh"
#include "zstd_enc_core.h"
#include "../../include/slow5b200.h"

namespace s5b {

namespace {

using namespace s5bz;

constexpr int ZE_WARPS = 4;
constexpr int ZE_BLOCK = 6144;  // max input bytes per block (multiple of 32)

struct ZeTreeScratch {  // live only while a block's code is being constructed
    uint32_t sortbuf[512];
    uint32_t weight[512];
    uint16_t parent[512];
};
struct __align__(128) ZeWarpSmem {
    // the staged block and the construction scratch share their bytes (the block is staged again, from L2, once
    // the code exists): fewer bytes per warp, more resident warps
    union {
        uint8_t in[16 + ZE_BLOCK + 16];
        ZeTreeScratch k;
    };
    uint32_t out[(HC_OUT + HC_OUT_SLACK) / 4];
    uint32_t hist[256];
    uint16_t code[256];
    uint16_t prev_code[256];
    uint8_t len[256];
    uint8_t prev_len[256];
    uint8_t weights[256];
    uint8_t tree[HUF_TREE_MAX_BYTES];
    alignas(4) uint16_t bl_count[16];
    WeightEnc we;
    unsigned long long bar;
};

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// appends n whole bytes (n <= 32 per call round) from smem to the bit buffer at a byte-aligned position
__device__ __forceinline__ void put_bytes(BitOut &bo, const uint8_t *src, uint32_t n, int lane) {
    for (uint32_t t0 = 0; t0 < n; t0 += 32) {
        if ((bo.bitpos >> 3) + 40 > HC_OUT) bo.flush(lane, false);
        const uint32_t i = t0 + lane;
        if (i < n) bo.put(bo.bitpos + 8 * lane, src[i], 8);
        bo.bitpos += 8 * min(32u, n - t0);
    }
}
// appends the low `nbytes` (<= 8) bytes of v (warp-uniform)
__device__ __forceinline__ void put_le(BitOut &bo, uint64_t v, int nbytes, int lane) {
    if ((bo.bitpos >> 3) + 16 > HC_OUT) bo.flush(lane, false);
    if (lane < nbytes) bo.put(bo.bitpos + 8 * lane, (uint32_t)(v >> (8 * lane)) & 0xffu, 8);
    bo.bitpos += 8 * nbytes;
}

}  // namespace

__global__ void __launch_bounds__(ZE_WARPS * 32) zstd_encode_kernel(const DeflateArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ZeWarpSmem &ws = reinterpret_cast<ZeWarpSmem *>(smem_raw)[threadIdx.x >> 5];
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
        // worst case: frame header + every block raw (3 bytes of header each; the split adds one block)
        const uint64_t nblocks_max = (uint64_t)ilen / ZE_BLOCK + 2;
        if (ioff + ilen > a.in_capacity || ocap < (uint64_t)ilen + 3 * nblocks_max + 12) {
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
        for (uint32_t i = lane; i < (HC_OUT + HC_OUT_SLACK) / 4; i += 32) ws.out[i] = 0;
        __syncwarp();
        {
            int fh = 0;
            if (lane == 0) fh = write_frame_header(ws.tree, ilen);
            fh = __shfl_sync(FULL, fh, 0);
            __syncwarp();
            put_bytes(bo, ws.tree, (uint32_t)fh, lane);
            __syncwarp();
        }
        bool have_table = false;  // ws.prev_len / ws.prev_code hold the table of the last Huffman-coded block

        uint32_t b0 = 0;
        do {  // at least one block, so an empty record still gets its (last, raw, empty) block
            uint32_t b1 = ilen;
            if (split > b0) b1 = split;
            if (b1 - b0 > ZE_BLOCK) b1 = b0 + ZE_BLOCK;
            const uint32_t n = b1 - b0;
            const bool last = b1 == ilen;
            // ---- stage the block
            const uint8_t *g0 = src + b0;
            const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(g0) & 15u);
            const uint8_t *g16 = g0 - skew;
            uint64_t bytes = ((uint64_t)skew + n + 15) & ~15ull;
            const uint64_t room = a.in_capacity - (uint64_t)(g16 - a.in);
            if (bytes > room) bytes = room & ~15ull;
            auto stage_block = [&]() {
                if (!bytes || !n) return;
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
            const uint8_t *blk = ws.in + skew;

            // ---- pass 1: byte histogram (one atomic per distinct value in a strip)
            for (int s = lane; s < 256; s += 32) ws.hist[s] = 0;
            __syncwarp();
            for (uint32_t t0 = 0; t0 < n; t0 += 32) {
                const uint32_t i = t0 + lane;
                const uint32_t b = i < n ? blk[i] : 0x100u;
                const uint32_t peers = __match_any_sync(FULL, b);
                if (i < n && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&ws.hist[b], (uint32_t)__popc(peers));
            }
            __syncwarp();
            uint32_t used = 0, maxc = 0;
            for (int s = lane; s < 256; s += 32) {
                used += ws.hist[s] != 0;
                maxc = max(maxc, ws.hist[s]);
            }
            used = warp_sum(used);

            // 0 = raw, 1 = RLE, 2 = Huffman with tree, 3 = Huffman treeless
            uint32_t mode = 0;
            uint32_t tree_bytes = 0, comp = 0, sz0 = 0, sz1 = 0, sz2 = 0, sz3 = 0;
            const bool four = n >= 256;
            const uint32_t q = four ? (n + 3) / 4 : n;
            if (n && used == 1) mode = 1;
            if (used >= 2 && n >= 16) {
                // ---- code construction: lengths (<= 11 bits), zstd numbering, tree description
                // cost with the previous table first (hist[] is modified by huffman_lengths only when used < 2)
                uint32_t cost_prev = 0xffffffffu;
                if (have_table) {
                    uint32_t c = 0, miss = 0;
                    for (int s = lane; s < 256; s += 32) {
                        c += ws.hist[s] * ws.prev_len[s];
                        miss += ws.hist[s] != 0 && ws.prev_len[s] == 0;
                    }
                    c = warp_sum(c);
                    miss = warp_sum(miss);
                    if (!miss) cost_prev = c;
                }
                huffman_lengths(ws.hist, 256, HUF_MAX_BITS, ws.len, ws.k.sortbuf, ws.k.weight, ws.k.parent, ws.bl_count, lane);
                int tb = 0;
                if (lane == 0) {
                    int nsym = 0;
                    const int mb = huf_codes_from_lengths(ws.len, ws.weights, ws.code, &nsym);
                    if (mb) tb = huf_write_tree(ws.we, ws.weights, nsym, ws.tree);
                }
                tb = __shfl_sync(FULL, tb, 0);
                stage_block();  // the construction scratch overwrote the block
                uint32_t cost_new = 0;
                for (int s = lane; s < 256; s += 32) cost_new += ws.hist[s] * ws.len[s];
                cost_new = warp_sum(cost_new);
                if (tb) {
                    mode = 2;
                    tree_bytes = (uint32_t)tb;
                }
                if (cost_prev != 0xffffffffu && (!tb || cost_prev <= cost_new + 8u * (uint32_t)tb)) {
                    mode = 3;
                    tree_bytes = 0;
                    for (int s = lane; s < 256; s += 32) {
                        ws.len[s] = ws.prev_len[s];
                        ws.code[s] = ws.prev_code[s];
                    }
                    __syncwarp();
                }
            }
            if (mode >= 2) {
                // ---- pass 2: exact stream sizes
                uint32_t bits[4] = {0, 0, 0, 0};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t s0 = k * q, s1 = min(n, s0 + q);
                    if (k == 0 || four)
                        for (uint32_t i = s0 + lane; i < s1; i += 32) bits[k] += ws.len[blk[i]];
                    bits[k] = warp_sum(bits[k]);
                }
                sz0 = (bits[0] + 8) >> 3;  // + end mark, rounded up to bytes
                if (four) {
                    sz1 = (bits[1] + 8) >> 3;
                    sz2 = (bits[2] + 8) >> 3;
                    sz3 = (bits[3] + 8) >> 3;
                }
                comp = tree_bytes + (four ? 6u : 0u) + sz0 + sz1 + sz2 + sz3;
                const uint32_t bsize = (uint32_t)literals_header_len(four, n, comp) + comp + 1u;
                // the jump table holds 16-bit sizes; a block that does not shrink is stored raw
                if (bsize >= n || sz0 > 0xffffu || sz1 > 0xffffu || sz2 > 0xffffu) mode = 0;
            }

            if (mode == 0) {
                put_le(bo, block_header(last, 0, n), 3, lane);
                put_bytes(bo, blk, n, lane);
            } else if (mode == 1) {
                put_le(bo, block_header(last, 1, n), 3, lane);
                put_le(bo, blk[0], 1, lane);
            } else {
                uint64_t lh;
                const int lhn = literals_header(mode, four, n, comp, &lh);
                put_le(bo, block_header(last, 2, (uint32_t)lhn + comp + 1u), 3, lane);
                put_le(bo, lh, lhn, lane);
                if (mode == 2) put_bytes(bo, ws.tree, tree_bytes, lane);
                if (four) put_le(bo, (uint64_t)sz0 | ((uint64_t)sz1 << 16) | ((uint64_t)sz2 << 32), 6, lane);
                // ---- pass 3: the streams, each written from its last symbol to its first
                for (int k = 0; k < (four ? 4 : 1); ++k) {
                    const int s0 = (int)(k * q), s1 = (int)min(n, (uint32_t)s0 + q);
                    for (int e = s1; e > s0; e -= 32) {
                        if ((bo.bitpos >> 3) + 64 > HC_OUT) bo.flush(lane, false);
                        const int i = e - 1 - lane;
                        uint32_t nb = 0, cw = 0;
                        if (i >= s0) {
                            const uint32_t b = blk[i];
                            nb = ws.len[b];
                            cw = ws.code[b];
                        }
                        uint32_t incl = nb;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t t = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += t;
                        }
                        bo.put(bo.bitpos + incl - nb, cw, nb);
                        bo.bitpos += __shfl_sync(FULL, incl, 31);
                    }
                    if (lane == 0) bo.put(bo.bitpos, 1u, 1);  // end mark
                    bo.bitpos = (bo.bitpos + 1 + 7) & ~7u;
                }
                put_le(bo, 0, 1, lane);  // Sequences_Section_Header: no sequences
                if (mode == 2) {
                    for (int s = lane; s < 256; s += 32) {
                        ws.prev_len[s] = ws.len[s];
                        ws.prev_code[s] = ws.code[s];
                    }
                    have_table = true;
                }
            }
            b0 = b1;
            __syncwarp();
        } while (b0 < ilen);
        bo.flush(lane, true);
        if (lane == 0) {
            a.out_len[r] = (uint32_t)bo.written;
            a.status[r] = S5B_OK;
        }
        __syncwarp();
    }
}

int zstd_encode_blocks_per_sm() {
    int n = 0;
    if (cudaFuncSetAttribute(zstd_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(ZeWarpSmem) * ZE_WARPS)) != cudaSuccess)
        return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, zstd_encode_kernel, ZE_WARPS * 32,
                                                      sizeof(ZeWarpSmem) * ZE_WARPS) != cudaSuccess)
        return 0;
    return n;
}

uint64_t zstd_encode_bound(uint64_t len) { return len + 3 * (len / ZE_BLOCK + 2) + 12; }

cudaError_t launch_zstd_encode(const DeflateArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t want = (a.n_reads + ZE_WARPS - 1) / ZE_WARPS;
    uint64_t cap = (uint64_t)num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (!grid) grid = 1;
    zstd_encode_kernel<<<grid, ZE_WARPS * 32, sizeof(ZeWarpSmem) * ZE_WARPS, st>>>(a);
    return cudaGetLastError();
}

}  // namespace s5b
