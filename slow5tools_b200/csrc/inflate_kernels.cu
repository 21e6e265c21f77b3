// inflate_kernels.cu -- sm_100a zlib/DEFLATE decoder, one warp per record stream.
//
// Replaces, for whole batches, ptr_depress_zlib_solo (slow5lib/src/slow5_press.c:973-1010), i.e.
// inflateInit2(15) + inflate(Z_NO_FLUSH) loop + inflateEnd from system zlib, with the same observable
// behaviour: RFC 1950 header check, stored / fixed / dynamic blocks, 32 KiB window, Adler-32 check;
// malformed data -> S5B_ERR_PRESS (zlib's Z_DATA_ERROR / Z_NEED_DICT arms, :993-999); input that ends
// early is NOT an error there (Z_BUF_ERROR / Z_OK fall through, :1001-1003) and yields the bytes decoded
// so far -- mirrored here; bytes after the Adler-32 trailer are ignored.
//
// Every record is an independent complete zlib stream (slow5.c:4046, slow5_press.c:868-871): one warp per
// stream.  Inside a Huffman block the symbols are decoded by ALL 32 lanes at once ("window decode"): the
// compressed bits ahead are cut into 32 chunks, every lane decodes its chunk speculatively from the chunk's
// first bit, and because Huffman codes self-synchronise a wrongly started lane falls onto true symbol
// boundaries within a few symbols; lanes then restart at their left neighbour's end position until all
// boundaries agree (usually one extra round), a warp prefix scan over the per-lane output byte counts gives
// the output offsets, and a last pass writes the literals while LZ77 matches are queued and resolved in stream
// order by the whole warp.  Anything unusual (stored blocks, invalid codes, the last < 48 bits of a stream,
// a too small output slot, windows that do not fit the staging buffer) is handed to the serial decoder --
// lane 0 running out of shared memory (input ring filled by 1-D bulk async copies, 10-bit / 8-bit first-level
// tables) -- which also owns every error / truncation verdict, so behaviour matches zlib's exactly.  All
// lanes cooperate on table construction, LZ77 copies, the Adler-32 and the 128-bit stores of the staging
// buffer.
#include <cstdlib>
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "../../include/slow5b200.h"

namespace s5b {

namespace {

constexpr int INF_WARPS = 4;
constexpr int INF_BLK = 512;   // input ring block (the ring only feeds the serial decoder: headers and tails)
constexpr int INF_NB = 2;
constexpr int INF_RING = INF_BLK * INF_NB;
constexpr int INF_STAGE = 4096;  // output staging bytes (a multiple of 16)
constexpr int PAR_CHUNK_MAX_BITS = 384;   // window decode: at most this many compressed bits per lane and window
                                          // (1.5 KiB of input, whose output should fit the free part of the stage)
constexpr int PAR_CHUNK_MIN_BITS = 64;
constexpr int PAR_CHUNK_START_BITS = 128; // first window of a block (its end is unknown: bits past the end-of-block
                                          // symbol are decoded for nothing), x3 for every further window
constexpr int PAR_LIST = 128;             // queued matches per window
constexpr int PAR_MAX_ROUNDS = 8;         // boundary rounds before only the agreed prefix of lanes is committed
constexpr uint32_t PAR_SYM_BITS = 48;     // longest length/distance pair: 15 + 5 + 15 + 13
constexpr int LIT_FAST_BITS = 10;
constexpr int DIST_FAST_BITS = 8;
constexpr uint32_t ADLER_MOD = 65521u;

struct __align__(128) InfWarpSmem {
    uint8_t ring[INF_RING];
    uint8_t stage[INF_STAGE + 16];
    uint16_t lit_fast[1 << LIT_FAST_BITS];    // (symbol << 4) | code length, 0 = not a short code
    uint16_t dist_fast[1 << DIST_FAST_BITS];
    uint16_t lit_sorted[288];                 // symbols ordered by (length, symbol)
    uint16_t dist_sorted[32];
    uint16_t lit_count[16];
    uint16_t dist_count[16];
    uint16_t cl_fast[128];
    uint16_t cl_sorted[20];
    uint16_t cl_count[16];
    uint16_t tmp_base[16];
    uint16_t lit_fcode[16], lit_fidx[16];    // canonical first code / first sorted index per length
    uint16_t lit_lim[16];                     // end of the codes of each length, left-aligned to 15 bits (saturated)
    uint16_t dist_fcode[16], dist_fidx[16];
    uint16_t cl_fcode[16], cl_fidx[16];
    uint8_t lens[384];                        // code lengths: lit/len at 0, dist at 288; scratch from 32
    uint2 list[PAR_LIST];                     // window decode: (stage position | length << 16, distance)
    unsigned long long bar[INF_NB];
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                        35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                         1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                                         9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// events lane 0 raises for the whole warp
enum : uint32_t { EV_FLUSH = 1, EV_MATCH = 2, EV_TABLES = 3, EV_END = 4 };
// how a stream ended
enum : int32_t { END_OK = 0, END_TRUNC = 1, END_ERR = 2 };

// ---- input side (lane 0 only) ----------------------------------------------------------------
struct RingState {
    uint32_t issued, waited, phase_bits;
};
// issue bulk copies for every ring block that is free: blocks [issued, min(nblk, done + INF_NB))
__device__ __noinline__ uint32_t ring_issue(uint32_t bar0, uint32_t ring0, const uint8_t *src16, uint64_t lim, uint32_t nblk,
                                            uint32_t done, uint32_t issued) {
    while (issued < nblk && issued < done + INF_NB) {
        const uint32_t slot = issued % INF_NB;
        const uint64_t b0 = (uint64_t)issued * INF_BLK;
        uint64_t bytes = lim - b0;
        if (bytes > INF_BLK) bytes = INF_BLK;
        mbar_arrive_expect_tx(bar0 + 8 * slot, (uint32_t)bytes);
        bulk_g2s(ring0 + slot * INF_BLK, src16 + b0, (uint32_t)bytes, bar0 + 8 * slot);
        ++issued;
    }
    return issued;
}
// wait until `need` blocks have landed (issuing into free blocks first, as the waits advance)
__device__ __noinline__ RingState ring_wait(uint32_t bar0, uint32_t ring0, const uint8_t *src16, uint64_t lim, uint32_t nblk,
                                            uint32_t done, uint32_t issued, uint32_t waited, uint32_t phase_bits,
                                            uint32_t need) {
    while (waited < need) {
        while (issued < nblk && issued < done + INF_NB) {
            const uint32_t slot = issued % INF_NB;
            const uint64_t b0 = (uint64_t)issued * INF_BLK;
            uint64_t bytes = lim - b0;
            if (bytes > INF_BLK) bytes = INF_BLK;
            mbar_arrive_expect_tx(bar0 + 8 * slot, (uint32_t)bytes);
            bulk_g2s(ring0 + slot * INF_BLK, src16 + b0, (uint32_t)bytes, bar0 + 8 * slot);
            ++issued;
        }
        const uint32_t slot = waited % INF_NB;
        mbar_wait(bar0 + 8 * slot, (phase_bits >> slot) & 1u);
        phase_bits ^= 1u << slot;
        ++waited;
    }
    RingState r;
    r.issued = issued;
    r.waited = waited;
    r.phase_bits = phase_bits;
    return r;
}

struct BitReader {
    const uint8_t *ring;     // smem
    const uint8_t *p0;       // first byte of the stream (global)
    const uint8_t *slab_end; // end of the input slab (bulk copies never read past it)
    uint32_t full_len;       // stream bytes
    uint32_t base;           // stream byte the ring was (re)started at: ipos / in_len count from here
    const uint8_t *src16;    // 16-byte aligned global address of ring position 0
    uint64_t lim;            // bulk-copyable bytes from src16
    uint32_t bar0, ring0;    // smem addresses
    uint32_t nblk, issued, waited, phase_bits;
    uint32_t skew;           // stream byte 0 is at ring position skew
    uint32_t in_len;         // stream bytes
    uint32_t ipos;           // stream bytes moved into the bit buffer
    uint64_t bb;
    uint32_t nb;

    // The ring maintenance (issue bulk copies into free blocks, wait for a block) lives in two out-of-line functions:
    // refill() is inlined at every symbol-decoding site, and carrying the TMA issue sequence along made the kernel
    // ~200 KB of code (instruction-fetch stalls were the largest stall reason).  State goes in and out by value so
    // the reader itself stays in registers.
    // a ring block may only be overwritten once every byte of it has been moved into bb
    __device__ __forceinline__ void recycle() {
        const uint32_t done = (skew + ipos) / INF_BLK;
        if (issued < nblk && issued < done + INF_NB) issued = ring_issue(bar0, ring0, src16, lim, nblk, done, issued);
    }
    // make ring position rp readable
    __device__ __forceinline__ void ensure(uint32_t rp) {
        const uint32_t need = rp / INF_BLK + 1;
        if (waited < need) {
            const RingState r = ring_wait(bar0, ring0, src16, lim, nblk, (skew + ipos) / INF_BLK, issued, waited, phase_bits, need);
            issued = r.issued;
            waited = r.waited;
            phase_bits = r.phase_bits;
        }
    }
    __device__ __forceinline__ void refill() {
        if (nb > 32) return;
        uint32_t rp = skew + ipos;
        if ((rp & 3u) == 0 && ipos + 4 <= in_len) {
            ensure(rp + 3);
            const uint32_t w = *reinterpret_cast<const uint32_t *>(ring + (rp & (INF_RING - 1)));
            bb |= (uint64_t)w << nb;
            nb += 32;
            ipos += 4;
            if (((skew + ipos) & (INF_BLK - 1)) == 0) recycle();
        } else {
            while (nb <= 56 && ipos < in_len) {
                ensure(rp);
                bb |= (uint64_t)ring[rp & (INF_RING - 1)] << nb;
                nb += 8;
                ++ipos;
                ++rp;
            }
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(uint32_t n) {
        bb >>= n;
        nb -= n;
    }
    // drain copies that were issued but never consumed (stream ended early / trailing bytes)
    __device__ __forceinline__ void drain() {
        while (waited < issued) {
            const uint32_t slot = waited % INF_NB;
            mbar_wait(bar0 + 8 * slot, (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            ++waited;
        }
    }
    // (re)start the ring at stream byte `byte`
    __device__ __forceinline__ void start_at(uint32_t byte) {
        const uint8_t *p = p0 + byte;
        base = byte;
        skew = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u);
        src16 = p - skew;
        in_len = full_len - byte;
        lim = ((uint64_t)skew + in_len + 15) & ~15ull;
        const uint64_t room = (uint64_t)(slab_end - src16);
        if (lim > room) lim = room & ~15ull;
        nblk = in_len ? (uint32_t)((lim + INF_BLK - 1) / INF_BLK) : 0;
        issued = waited = 0;
        ipos = 0;
        bb = 0;
        nb = 0;
        recycle();
    }
    // stream bit position of the next unread bit
    __device__ __forceinline__ uint32_t bitpos() const { return (base + ipos) * 8u - nb; }
    // continue at an arbitrary stream bit (after a window decode ran ahead of the ring)
    __device__ __forceinline__ void seek_bit(uint32_t bit) {
        drain();
        start_at(bit >> 3);
        refill();
        const uint32_t frac = bit & 7u;
        if (nb >= frac) drop(frac);
    }
};

// ---- window decode (all lanes) ------------------------------------------------------------------
// how a lane's run over its chunk ended
enum : uint32_t { PS_NONE = 0, PS_EOB = 1, PS_INVALID = 2, PS_TAIL = 3 };
// what a window did
enum : uint32_t { PAR_CONT = 0, PAR_EOB = 1, PAR_SERIAL = 2, PAR_FLUSH = 3, PAR_ERR = 4 };

struct ParIn {               // warp-uniform view of the compressed stream
    const uint32_t *w4;      // 4-byte aligned global address at or below the first stream byte
    uint32_t bit0;           // position of stream bit 0 inside w4[0] (0, 8, 16 or 24)
    uint32_t nwords;         // words that hold stream bytes
    uint32_t end_bits;       // stream length in bits
};

struct LaneBits {
    uint64_t bb;
    uint32_t nb, wi;  // wi: index of the word held in `ahead`
    uint32_t ahead;   // loaded one refill early, so its latency hides behind the symbols in between
    __device__ __forceinline__ void seek(const ParIn &in, uint32_t p) {
        const uint32_t q = p + in.bit0;
        wi = q >> 5;
        const uint32_t w = wi < in.nwords ? __ldg(in.w4 + wi) : 0u;
        ++wi;
        ahead = wi < in.nwords ? __ldg(in.w4 + wi) : 0u;
        bb = w >> (q & 31u);
        nb = 32u - (q & 31u);
    }
    __device__ __forceinline__ void refill(const ParIn &in) {
        if (nb <= 32) {
            bb |= (uint64_t)ahead << nb;
            nb += 32;
            ++wi;
            ahead = wi < in.nwords ? __ldg(in.w4 + wi) : 0u;
        }
    }
    __device__ __forceinline__ void drop(uint32_t n) {
        bb >>= n;
        nb -= n;
    }
};

// canonical walk for codes longer than the first-level table, on a lane's own bit buffer (>= 15 valid bits)
__device__ __forceinline__ int slow_walk(uint64_t bits, const uint16_t *count, const uint16_t *sorted, const uint16_t *fcode,
                                         const uint16_t *fidx, const int fast_bits, uint32_t *used) {
    uint32_t code = __brev((uint32_t)bits) >> (32 - fast_bits);
    uint64_t b = bits >> fast_bits;
    for (uint32_t len = fast_bits + 1; len <= 15; ++len) {
        code = (code << 1) | (uint32_t)(b & 1u);
        b >>= 1;
        const uint32_t d = code - fcode[len];
        if (d < count[len]) {
            *used = len;
            return sorted[fidx[len] + d];
        }
    }
    return -1;
}

struct LenDistTabs {  // CTA-shared copies of the RFC 1951 base / extra-bit tables (divergent indices)
    uint16_t len_base[32];
    uint16_t dist_base[32];
    uint8_t len_extra[32];
    uint8_t dist_extra[32];
};

// One lane decodes the symbols that START in [start, limit).  WRITE = false: count output bytes and matches;
// WRITE = true: literals go to out[], matches are queued as (stage position, length, distance).
template <bool WRITE>
__device__ __forceinline__ void par_run(const InfWarpSmem &ws, const LenDistTabs &ld, const ParIn &in, const uint32_t start,
                                        const uint32_t limit, uint32_t &end, uint32_t &nbytes, uint32_t &nmatch,
                                        uint32_t &stop, uint8_t *stage, uint32_t o, uint2 *list, uint32_t list_at) {
    LaneBits b;
    uint32_t p = start, nby = 0, nm = 0, st = PS_NONE;
    // the last PAR_SYM_BITS of the stream belong to the serial decoder (it owns the truncation verdicts)
    static_assert(LIT_FAST_BITS == 10, "the long-code path below compares against the limits of lengths 11..15");
    const uint32_t lim11 = ws.lit_lim[11], lim12 = ws.lit_lim[12], lim13 = ws.lit_lim[13], lim14 = ws.lit_lim[14],
                   lim15 = ws.lit_lim[15];
    const uint32_t safe = in.end_bits - PAR_SYM_BITS;
    const uint32_t lim2 = min(limit, safe + 1u);
    if (p < lim2) b.seek(in, p);
    if (p < limit && p > safe) st = PS_TAIL;
    while (p < lim2) {
        b.refill(in);
        uint32_t e = ws.lit_fast[(uint32_t)b.bb & ((1u << LIT_FAST_BITS) - 1u)];
        uint32_t l = e & 15u;
        int sym = (int)(e >> 4);
        if (l == 0) {
            // code longer than the first-level table: canonical codes, left-aligned to 15 bits, ascend with their
            // length, so four compares against the per-length limits give the length
            const uint32_t c15 = __brev((uint32_t)b.bb) >> 17;
            if (c15 >= lim15) {
                st = PS_INVALID;
                break;
            }
            l = 11u + (c15 >= lim11) + (c15 >= lim12) + (c15 >= lim13) + (c15 >= lim14);
            sym = ws.lit_sorted[ws.lit_fidx[l] + ((c15 >> (15u - l)) - ws.lit_fcode[l])];
        }
        if (sym < 256) {
            if (WRITE) stage[o] = (uint8_t)sym;
            ++o;
            ++nby;
            b.drop(l);
            p += l;
            continue;
        }
        if (sym == 256) {
            p += l;
            st = PS_EOB;
            break;
        }
        if (sym > 285) {
            st = PS_INVALID;
            break;
        }
        b.drop(l);
        const uint32_t li = (uint32_t)sym - 257u;
        const uint32_t lx = ld.len_extra[li];
        const uint32_t mlen = ld.len_base[li] + ((uint32_t)b.bb & ((1u << lx) - 1u));
        b.drop(lx);
        b.refill(in);
        const uint32_t de = ws.dist_fast[(uint32_t)b.bb & ((1u << DIST_FAST_BITS) - 1u)];
        uint32_t dl = de & 15u;
        int dsym = (int)(de >> 4);
        if (dl == 0) {
            dsym = slow_walk(b.bb, ws.dist_count, ws.dist_sorted, ws.dist_fcode, ws.dist_fidx, DIST_FAST_BITS, &dl);
            if (dsym < 0) {
                st = PS_INVALID;
                break;
            }
        }
        if (dsym > 29) {
            st = PS_INVALID;
            break;
        }
        b.drop(dl);
        const uint32_t dx = ld.dist_extra[dsym];
        const uint32_t dist = ld.dist_base[dsym] + ((uint32_t)b.bb & ((1u << dx) - 1u));
        b.drop(dx);
        p += l + lx + dl + dx;
        if (WRITE) list[list_at + nm] = make_uint2(o | (mlen << 16), dist);
        o += mlen;
        nby += mlen;
        ++nm;
    }
    if (st == PS_NONE && p < limit) st = PS_TAIL;
    end = p;
    nbytes = nby;
    nmatch = nm;
    stop = st;
}

// canonical decode for codes longer than the first-level table: the first `fast_bits` bits are already known
// not to form a code, so the walk starts at length fast_bits + 1 with the per-length first code / first index
// that build_table left in shared memory.  Returns the symbol, -1 (no such code) or -2 (ran out of bits).
__device__ __forceinline__ int slow_decode(const BitReader &br, const uint16_t *count, const uint16_t *sorted,
                                           const uint16_t *fcode, const uint16_t *fidx, const int fast_bits,
                                           uint32_t *used) {
    uint32_t code = __brev((uint32_t)br.bb) >> (32 - fast_bits);  // stream bits are the code MSB first
    uint64_t b = br.bb >> fast_bits;
    for (uint32_t len = fast_bits + 1; len <= 15; ++len) {
        if (len > br.nb) return -2;
        code = (code << 1) | (uint32_t)(b & 1u);
        b >>= 1;
        const uint32_t d = code - fcode[len];
        if (d < count[len]) {
            *used = len;
            return sorted[fidx[len] + d];
        }
    }
    return br.nb < 15 ? -2 : -1;
}

// ---- Huffman table construction (whole warp) ----------------------------------------------------
// lens[0..n): code lengths (0 = unused).  Builds count[], sorted[] (canonical order) and the first-level
// table fast[] of 2^fast_bits entries.  Returns 0 ok, 1 over-subscribed, 2 incomplete (caller decides,
// zlib's inflate_table rules: inftrees.c).
__device__ int build_table(const uint8_t *lens, int n, uint16_t *count, uint16_t *sorted, uint16_t *fast,
                           int fast_bits, uint16_t *base /*[16] scratch*/, uint16_t *fcode, uint16_t *fidx, int lane,
                           int *max_len_out, uint16_t *lim = nullptr) {
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const uint32_t l = lens[s];
        if (l) atomicAdd(reinterpret_cast<unsigned int *>(count) + (l >> 1), (l & 1) ? 0x10000u : 1u);
    }
    __syncwarp();
    // Kraft check + first index per length (serial over 15 lengths; every lane computes the same values)
    int left = 1, status = 0, maxl = 0;
    uint32_t idx = 0;
    uint32_t first_idx[16];
    uint32_t first_code[16];
    uint32_t code = 0;
    first_idx[0] = 0;
    first_code[0] = 0;
#pragma unroll
    for (int l = 1; l <= 15; ++l) {
        const int c = count[l];
        left <<= 1;
        left -= c;
        if (left < 0) status = 1;
        if (c) maxl = l;
        first_idx[l] = idx;
        first_code[l] = code;
        idx += c;
        code = (code + c) << 1;
    }
    if (status == 0 && left > 0) status = 2;
    *max_len_out = maxl;
    if (status == 1) return 1;
    if (lane < 16) {
        uint32_t v = 0, fc = 0;
#pragma unroll
        for (int l = 1; l <= 15; ++l)
            if (l == lane) {
                v = first_idx[l];
                fc = first_code[l];
            }
        base[lane] = (uint16_t)v;
        fidx[lane] = (uint16_t)v;
        fcode[lane] = (uint16_t)fc;
        if (lim) lim[lane] = lane ? (uint16_t)min(0xffffu, (fc + count[lane]) << (15 - lane)) : (uint16_t)0;
    }
    for (int i = lane; i < (1 << fast_bits); i += 32) fast[i] = 0;
    __syncwarp();
    // rank of each symbol inside its length class, in symbol order, 32 symbols at a time
    for (int s0 = 0; s0 < n; s0 += 32) {
        const int s = s0 + lane;
        const uint32_t l = s < n ? lens[s] : 0;
        const unsigned same = __match_any_sync(FULL, l);
        uint32_t rev = 0, e = 0;
        bool fill = false;
        if (l) {
            const uint32_t rank = base[l] + __popc(same & ((1u << lane) - 1u));
            sorted[rank] = (uint16_t)s;
            if ((int)l <= fast_bits) {
                const uint32_t cw = fcode[l] + (rank - fidx[l]);  // canonical code, MSB first
                rev = __brev(cw) >> (32 - l);                      // as it appears in the LSB-first bit stream
                e = ((uint32_t)s << 4) | l;
                fill = true;
            }
        }
        // codes that own >= 32 table entries are filled by the whole warp, the others by their lane
        const int wide = fast_bits - 5;
        unsigned widemask = __ballot_sync(FULL, fill && (int)l <= wide);
        if (fill && (int)l > wide)
            for (uint32_t i = rev; i < (1u << fast_bits); i += (1u << l)) fast[i] = (uint16_t)e;
        while (widemask) {
            const int src = __ffs(widemask) - 1;
            widemask &= widemask - 1;
            const uint32_t r2 = __shfl_sync(FULL, rev, src), l2 = __shfl_sync(FULL, l, src), e2 = __shfl_sync(FULL, e, src);
            for (uint32_t i = r2 + ((uint32_t)lane << l2); i < (1u << fast_bits); i += (32u << l2)) fast[i] = (uint16_t)e2;
        }
        __syncwarp();
        if (l && __ffs(same) - 1 == lane) base[l] += (uint16_t)__popc(same);  // one leader per length class
        __syncwarp();
    }
    return status;
}

// ---- Adler-32 over n staged bytes (whole warp) -----------------------------------------------
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

}  // namespace

__global__ void __launch_bounds__(INF_WARPS * 32) inflate_kernel(const InflateArgs a, const uint32_t min_len) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    InfWarpSmem &ws = reinterpret_cast<InfWarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(&ws.bar[0]);
    __shared__ LenDistTabs ld;
    if (threadIdx.x < 32) {
        ld.len_base[threadIdx.x] = threadIdx.x < 29 ? c_len_base[threadIdx.x] : 0;
        ld.len_extra[threadIdx.x] = threadIdx.x < 29 ? c_len_extra[threadIdx.x] : 0;
        ld.dist_base[threadIdx.x] = threadIdx.x < 30 ? c_dist_base[threadIdx.x] : 0;
        ld.dist_extra[threadIdx.x] = threadIdx.x < 30 ? c_dist_extra[threadIdx.x] : 0;
    }
    __syncthreads();
    if (lane == 0) {
        for (int s = 0; s < INF_NB; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t phase_bits = 0;

    for (;;) {
        unsigned long long r = 0;
        if (lane == 0) r = atomicAdd(a.work_counter, 1ULL);
        r = __shfl_sync(FULL, r, 0);
        if (r >= a.n_reads) break;
        const uint64_t ioff = a.in_off[r];
        const uint32_t ilen = a.in_len[r];
        if (ilen < min_len) continue;  // short streams belong to the thread-per-stream kernel (inflate_thread_kernels.cu)
        const uint64_t ooff = a.out_off[r];
        const uint64_t ocap = a.out_off[r + 1] - ooff;
        if (ioff + ilen > a.in_capacity) {
            if (lane == 0) {
                a.status[r] = S5B_ERR_ARG;
                a.out_len[r] = 0;
            }
            continue;
        }
        uint8_t *dst = a.out + ooff;

        // ---- input side: only lane 0 ever advances it
        BitReader br;
        ParIn pin;
        {
            const uint8_t *p = a.in + ioff;
            br.ring = ws.ring;
            br.p0 = p;
            br.slab_end = a.in + a.in_capacity;
            br.full_len = ilen;
            br.bar0 = bar0;
            br.ring0 = smem_u32(&ws.ring[0]);
            br.phase_bits = phase_bits;
            br.issued = br.waited = 0;
            if (lane == 0) br.start_at(0);
            const uint32_t sk4 = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
            pin.w4 = reinterpret_cast<const uint32_t *>(p - sk4);
            pin.bit0 = sk4 * 8u;
            pin.nwords = (sk4 + ilen + 3u) >> 2;
            pin.end_bits = ilen * 8u;
        }
        // window decode is attempted once per Huffman block, right after its tables are built (warp-uniform)
        bool try_par = false;
        uint32_t par_pos = 0;  // stream bit the next window starts at while try_par is set
        uint32_t par_chunk = PAR_CHUNK_START_BITS;
        const bool par_allowed = ilen < (1u << 28);
        // ---- output staging (warp-uniform): stage[i] <-> global gbase[i], gbase 16-byte aligned.
        //   [vstart, spos) valid bytes not yet stored, [astart, spos) not yet counted in total / Adler-32.
        //   absolute output position of stage[i] = total - astart + i.
        uint32_t vstart = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
        uint8_t *gbase = dst - vstart;
        uint32_t astart = vstart, spos = vstart;
        uint64_t total = 0;
        bool store = true;  // false once the slot overflowed: count only
        uint32_t ad_a = 1, ad_b = 0;
        int32_t end_kind = END_OK;
        bool stream_done = false;
        // lane-0 block state
        bool last_block = false, in_block = false;
        uint32_t stored_left = 0, btype = 0;

        // ---- zlib header (RFC 1950), lane 0
        if (lane == 0) {
            br.refill();
            if (br.nb < 16) {
                end_kind = END_TRUNC;
            } else {
                const uint32_t cmf = br.peek(8), flg = br.peek(16) >> 8;
                br.drop(16);
                if (((cmf << 8) | flg) % 31u != 0 || (cmf & 15u) != 8 || (cmf >> 4) > 7 || (flg & 0x20u)) end_kind = END_ERR;
            }
        }
        end_kind = __shfl_sync(FULL, end_kind, 0);
        stream_done = end_kind != END_OK;

        while (!stream_done) {
            uint32_t ev = 0, ev_a = 0, ev_b = 0;
            if (try_par) {
                // ---- window decode: all lanes, see the header comment
                uint32_t result = PAR_SERIAL;
                const uint32_t P0 = par_pos;
                const uint32_t rem = pin.end_bits - P0;
                if (rem >= 2u * PAR_SYM_BITS) {
                    const uint32_t W = min(rem, 32u * par_chunk);
                    par_chunk = min(par_chunk * 3u, (uint32_t)PAR_CHUNK_MAX_BITS);
                    const uint32_t C = max((uint32_t)PAR_CHUNK_MIN_BITS, (W + 31u) >> 5);
                    const uint32_t We = P0 + W;
                    uint32_t start = min(P0 + (uint32_t)lane * C, We);
                    const uint32_t limit = min(P0 + ((uint32_t)lane + 1u) * C, We);
                    uint32_t end, nby, nm, stop;
                    par_run<false>(ws, ld, pin, start, limit, end, nby, nm, stop, nullptr, 0, nullptr, 0);
                    // boundary rounds: a lane whose left neighbour ended somewhere else than it started decodes again
                    uint32_t agreed = 32;
                    for (int round = 1;; ++round) {
                        const uint32_t pe = __shfl_up_sync(FULL, end, 1);
                        const uint32_t ps = __shfl_up_sync(FULL, stop, 1);
                        const bool need = lane > 0 && ps == PS_NONE && pe != start;
                        const uint32_t needmask = __ballot_sync(FULL, need);
                        if (!needmask) break;
                        if (round > PAR_MAX_ROUNDS) {
                            agreed = (uint32_t)__ffs(needmask) - 1u;  // lanes below the first disagreement are exact
                            break;
                        }
                        if (need) {
                            start = pe;
                            par_run<false>(ws, ld, pin, start, limit, end, nby, nm, stop, nullptr, 0, nullptr, 0);
                        }
                    }
                    // lanes [0, nvalid) hold the true decode; the last of them may have stopped (EOB / invalid / tail)
                    const uint32_t stopmask = __ballot_sync(FULL, stop != PS_NONE) & (agreed < 32 ? (1u << agreed) - 1u : FULL);
                    const uint32_t nvalid = stopmask ? (uint32_t)__ffs(stopmask) : agreed;
                    const uint32_t cb = (uint32_t)lane < nvalid ? nby : 0u, cm = (uint32_t)lane < nvalid ? nm : 0u;
                    uint32_t ib = cb, im = cm;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t tb = __shfl_up_sync(FULL, ib, d), tm = __shfl_up_sync(FULL, im, d);
                        if (lane >= d) {
                            ib += tb;
                            im += tm;
                        }
                    }
                    const uint32_t win_bytes = __shfl_sync(FULL, ib, 31);
                    if (total + (spos - astart) + win_bytes <= ocap) {  // else: the serial decoder owns slot overflow
                        const uint32_t room = (uint32_t)INF_STAGE - spos;
                        const bool fits = (uint32_t)lane < nvalid && ib <= room && im <= (uint32_t)PAR_LIST;
                        const uint32_t fitmask = __ballot_sync(FULL, fits);
                        const uint32_t m = fitmask == FULL ? 32u : (uint32_t)__ffs(~fitmask) - 1u;  // leading lanes that fit
                        if (m == 0) {
                            result = spos > 64 ? PAR_FLUSH : PAR_SERIAL;
                        } else {
                            if ((uint32_t)lane < m) {
                                uint32_t e2, b2, m2, s2;
                                par_run<true>(ws, ld, pin, start, limit, e2, b2, m2, s2, ws.stage, spos + ib - cb, ws.list, im - cm);
                            }
                            __syncwarp();
                            const uint32_t tot_b = __shfl_sync(FULL, ib, m - 1), tot_m = __shfl_sync(FULL, im, m - 1);
                            // queued matches, in stream order; sources that left the stage are read back through L2
                            const int64_t stage0 = (int64_t)total - (int64_t)astart;  // absolute position of stage[0]
                            bool bad = false;
                            for (uint32_t k = 0; k < tot_m; ++k) {
                                const uint2 mt = ws.list[k];
                                const uint32_t o = mt.x & 0xffffu, mlen = mt.x >> 16, dist = mt.y;
                                const int64_t pos = stage0 + o;
                                if ((int64_t)dist > pos) {  // "invalid distance too far back"
                                    bad = true;
                                    break;
                                }
                                if (dist <= o - vstart) {  // the usual case: source inside the stage
                                    const uint8_t *srcb = ws.stage + (o - dist);
                                    if (dist >= mlen) {
                                        for (uint32_t kk = lane; kk < mlen; kk += 32) ws.stage[o + kk] = srcb[kk];
                                    } else {
                                        for (uint32_t kk = lane; kk < mlen; kk += 32) ws.stage[o + kk] = srcb[kk % dist];
                                    }
                                } else {
                                    for (uint32_t kk = lane; kk < mlen; kk += 32) {
                                        const int64_t sp = pos - dist + (kk % dist);
                                        const int64_t si = sp - stage0;
                                        ws.stage[o + kk] = si >= (int64_t)vstart ? ws.stage[si] : __ldcg(dst + sp);
                                    }
                                }
                                __syncwarp();
                            }
                            if (bad) {
                                result = PAR_ERR;
                            } else {
                                spos += tot_b;
                                par_pos = __shfl_sync(FULL, end, m - 1);
                                const uint32_t last_stop = __shfl_sync(FULL, stop, m - 1);
                                result = last_stop == PS_NONE ? PAR_CONT : last_stop == PS_EOB ? PAR_EOB : PAR_SERIAL;
                            }
                        }
                    }
                }
                if (result == PAR_ERR) {
                    ev = EV_END;
                    ev_a = END_ERR;
                } else if (result == PAR_FLUSH) {
                    ev = EV_FLUSH;
                } else {
                    if (result != PAR_CONT) {  // leaving the block (EOB) or handing the rest of it to lane 0
                        try_par = false;
                        if (lane == 0) {
                            if (br.bitpos() != par_pos) br.seek_bit(par_pos);
                            if (result == PAR_EOB) in_block = false;
                        }
                    }
                    if (spos > (uint32_t)INF_STAGE / 4) ev = EV_FLUSH;  // keep room for the next window's output
                }
                __syncwarp();
            } else {
            if (lane == 0) {
                // ---- decode until something needs the whole warp
                for (;;) {
                    if (!in_block) {
                        if (last_block) {
                            // Adler-32 trailer: big endian, byte aligned, after the final block
                            ev = EV_END;
                            ev_a = END_OK;
                            br.drop(br.nb & 7u);
                            br.refill();
                            if (br.nb < 32) {
                                ev_a = END_TRUNC;
                            } else {
                                ev_b = __byte_perm((uint32_t)br.bb, 0, 0x0123);
                                br.drop(32);
                            }
                            break;
                        }
                        br.refill();
                        if (br.nb < 3) {
                            ev = EV_END;
                            ev_a = END_TRUNC;
                            break;
                        }
                        last_block = br.peek(1);
                        btype = br.peek(3) >> 1;
                        br.drop(3);
                        if (btype == 3) {  // "invalid block type"
                            ev = EV_END;
                            ev_a = END_ERR;
                            break;
                        }
                        if (btype == 0) {
                            br.drop(br.nb & 7u);
                            br.refill();
                            if (br.nb < 32) {
                                ev = EV_END;
                                ev_a = END_TRUNC;
                                break;
                            }
                            const uint32_t len = br.peek(16), nlen = (uint32_t)(br.bb >> 16) & 0xffffu;
                            br.drop(32);
                            if ((len ^ 0xffffu) != nlen) {  // "invalid stored block lengths"
                                ev = EV_END;
                                ev_a = END_ERR;
                                break;
                            }
                            stored_left = len;
                            in_block = true;
                            continue;
                        }
                        ev = EV_TABLES;  // fixed or dynamic: the warp builds the tables
                        break;
                    }
                    if (btype == 0) {
                        // stored bytes travel through the bit buffer like literals (rare path)
                        if (stored_left == 0) {
                            in_block = false;
                            continue;
                        }
                        br.refill();
                        if (br.nb < 8) {
                            ev = EV_END;
                            ev_a = END_TRUNC;
                            break;
                        }
                        if (store) ws.stage[spos] = (uint8_t)br.peek(8);
                        br.drop(8);
                        ++spos;
                        --stored_left;
                        if (spos >= INF_STAGE) {
                            ev = EV_FLUSH;
                            break;
                        }
                        continue;
                    }
                    // ---- Huffman-coded symbols.  Fast loop first: literals with short codes, everything in registers.
                    // A table entry is (symbol << 4) | length: 1..4095 means "literal, length known"; 0 (long code)
                    // and >= 4096 (end of block / length symbol) leave the loop for the general code below.
                    if (store) {
                        uint64_t bb = br.bb;
                        uint32_t nb = br.nb, sp = spos;
                        const uint16_t *fast = ws.lit_fast;
                        uint8_t *stage = ws.stage;
                        for (;;) {
                            if (nb <= 32) {
                                br.bb = bb;
                                br.nb = nb;
                                br.refill();
                                bb = br.bb;
                                nb = br.nb;
                                if (nb < 15) break;  // input nearly exhausted: let the careful path decide
                            }
                            uint32_t e = fast[(uint32_t)bb & ((1u << LIT_FAST_BITS) - 1u)];
                            if (sp >= (uint32_t)INF_STAGE) break;
                            if (e - 1u >= 4095u) {
                                if (e != 0) break;  // end of block or a length symbol
                                // code longer than the table: short canonical walk (nb >= 15 here)
                                br.bb = bb;
                                br.nb = nb;
                                uint32_t l2 = 0;
                                const int s2 = slow_decode(br, ws.lit_count, ws.lit_sorted, ws.lit_fcode, ws.lit_fidx, LIT_FAST_BITS, &l2);
                                if (s2 < 0 || s2 >= 256) break;  // invalid or non-literal: the careful path redoes it
                                e = ((uint32_t)s2 << 4) | l2;
                            }
                            const uint32_t l = e & 15u;
                            stage[sp++] = (uint8_t)(e >> 4);
                            bb >>= l;
                            nb -= l;
                        }
                        br.bb = bb;
                        br.nb = nb;
                        spos = sp;
                        if (spos >= INF_STAGE) {
                            ev = EV_FLUSH;
                            break;
                        }
                    }
                    br.refill();
                    uint32_t e = ws.lit_fast[br.peek(LIT_FAST_BITS)];
                    uint32_t l = e & 15u;
                    int sym = (int)(e >> 4);
                    if (l == 0) {
                        sym = slow_decode(br, ws.lit_count, ws.lit_sorted, ws.lit_fcode, ws.lit_fidx, LIT_FAST_BITS, &l);
                        if (sym < 0) {
                            ev = EV_END;
                            ev_a = sym == -2 ? END_TRUNC : END_ERR;
                            break;
                        }
                    } else if (l > br.nb) {
                        ev = EV_END;
                        ev_a = END_TRUNC;
                        break;
                    }
                    if (sym < 256) {
                        br.drop(l);
                        if (store) ws.stage[spos] = (uint8_t)sym;
                        ++spos;
                        if (spos >= INF_STAGE) {
                            ev = EV_FLUSH;
                            break;
                        }
                        continue;
                    }
                    if (sym == 256) {
                        br.drop(l);
                        in_block = false;
                        continue;
                    }
                    if (sym > 285) {  // "invalid literal/length code"
                        ev = EV_END;
                        ev_a = END_ERR;
                        break;
                    }
                    // length / distance pair: nothing is consumed for good until the whole pair is available
                    const uint64_t save_bb = br.bb;
                    const uint32_t save_nb = br.nb, save_ipos = br.ipos;
                    br.drop(l);
                    const uint32_t li = (uint32_t)sym - 257u;
                    const uint32_t lx = c_len_extra[li];
                    bool trunc = br.nb < lx;
                    uint32_t mlen = 0;
                    int dsym = 0;
                    uint32_t dl = 0, dist = 0;
                    bool err = false;
                    if (!trunc) {
                        mlen = c_len_base[li] + br.peek(lx);
                        br.drop(lx);
                        br.refill();
                        const uint32_t de = ws.dist_fast[br.peek(DIST_FAST_BITS)];
                        dl = de & 15u;
                        dsym = (int)(de >> 4);
                        if (dl == 0) {
                            dsym = slow_decode(br, ws.dist_count, ws.dist_sorted, ws.dist_fcode, ws.dist_fidx, DIST_FAST_BITS, &dl);
                            if (dsym == -1) err = true;
                            if (dsym == -2) trunc = true;
                        } else if (dl > br.nb) {
                            trunc = true;
                        }
                    }
                    if (!trunc && !err) {
                        if (dsym > 29) {  // "invalid distance code"
                            err = true;
                        } else {
                            br.drop(dl);
                            const uint32_t dx = c_dist_extra[dsym];
                            if (br.nb < dx) {
                                trunc = true;
                            } else {
                                dist = c_dist_base[dsym] + br.peek(dx);
                                br.drop(dx);
                            }
                        }
                    }
                    if (err) {
                        ev = EV_END;
                        ev_a = END_ERR;
                        break;
                    }
                    if (trunc) {
                        // (the refill above may have advanced ipos; the saved buffer is a prefix of the new one,
                        //  so only bb/nb/ipos need restoring and the stream simply ends here)
                        br.bb = save_bb;
                        br.nb = save_nb;
                        br.ipos = save_ipos;
                        ev = EV_END;
                        ev_a = END_TRUNC;
                        break;
                    }
                    if ((uint64_t)dist > total + (spos - astart)) {  // "invalid distance too far back"
                        ev = EV_END;
                        ev_a = END_ERR;
                        break;
                    }
                    if (!store) {
                        spos += mlen;  // counting only; folded into `total` at the next flush event
                        if (spos >= INF_STAGE) {
                            ev = EV_FLUSH;
                            break;
                        }
                        continue;
                    }
                    if (mlen <= 12 && dist <= spos - vstart && spos + mlen <= INF_STAGE) {
                        // short, near match: lane 0 copies it byte by byte (overlap-safe)
                        for (uint32_t i = 0; i < mlen; ++i) ws.stage[spos + i] = ws.stage[spos - dist + i];
                        spos += mlen;
                        if (spos >= INF_STAGE) {
                            ev = EV_FLUSH;
                            break;
                        }
                        continue;
                    }
                    ev = EV_MATCH;
                    ev_a = mlen;
                    ev_b = dist;
                    break;
                }
            }
            __syncwarp();
            ev = __shfl_sync(FULL, ev, 0);
            ev_a = __shfl_sync(FULL, ev_a, 0);
            ev_b = __shfl_sync(FULL, ev_b, 0);
            spos = __shfl_sync(FULL, spos, 0);
            }

            if (ev == EV_TABLES) {
                // ---- fixed (btype 1) or dynamic (btype 2) code tables
                const uint32_t bt = __shfl_sync(FULL, btype, 0);
                int hlit = 288, hdist = 30, bad = 0;  // bad: 2 = data error, 3 = input ended
                if (bt == 1) {
                    for (int s = lane; s < 288; s += 32) ws.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
                    ws.lens[288 + lane] = 5;  // all 32 five-bit codes (30 and 31 decode to "invalid distance code")
                    hdist = 32;
                    __syncwarp();
                } else {
                    if (lane == 0) {
                        uint32_t hclen = 0;
                        br.refill();
                        if (br.nb < 14) {
                            bad = 3;
                        } else {
                            hlit = br.peek(5) + 257;
                            hdist = (br.peek(10) >> 5) + 1;
                            hclen = (br.peek(14) >> 10) + 4;
                            br.drop(14);
                            if (hlit > 286 || hdist > 30) bad = 2;  // "too many length or distance symbols"
                        }
                        for (int i = 0; i < 19; ++i) ws.lens[i] = 0;
                        for (uint32_t i = 0; i < hclen && !bad; ++i) {
                            br.refill();
                            if (br.nb < 3) {
                                bad = 3;
                                break;
                            }
                            ws.lens[c_cl_order[i]] = (uint8_t)br.peek(3);
                            br.drop(3);
                        }
                    }
                    bad = __shfl_sync(FULL, bad, 0);
                    hlit = __shfl_sync(FULL, hlit, 0);
                    hdist = __shfl_sync(FULL, hdist, 0);
                    __syncwarp();
                    int maxl = 0;
                    if (!bad) {
                        // "invalid code lengths set": the code-length code must be complete (inftrees.c, type CODES)
                        if (build_table(ws.lens, 19, ws.cl_count, ws.cl_sorted, ws.cl_fast, 7, ws.tmp_base, ws.cl_fcode, ws.cl_fidx, lane, &maxl) != 0) bad = 2;
                    }
                    __syncwarp();
                    // the hlit + hdist code lengths, run-length coded, lane 0 -> lens[32 ..)
                    if (lane == 0 && !bad) {
                        int i = 0;
                        const int nsym = hlit + hdist;
                        while (i < nsym) {
                            br.refill();
                            const uint32_t e = ws.cl_fast[br.peek(7)];
                            const uint32_t l = e & 15u;
                            const uint32_t sym = e >> 4;
                            if (l == 0 || l > br.nb) {
                                bad = 3;  // complete 7-bit code: a miss can only mean the bits ran out
                                break;
                            }
                            if (sym < 16) {
                                br.drop(l);
                                ws.lens[32 + i++] = (uint8_t)sym;
                                continue;
                            }
                            const uint32_t need = sym == 16 ? 2 : sym == 17 ? 3 : 7;
                            if (br.nb < l + need) {
                                bad = 3;
                                break;
                            }
                            br.drop(l);
                            uint32_t rep, val = 0;
                            if (sym == 16) {
                                if (i == 0) {
                                    bad = 2;  // "invalid bit length repeat"
                                    break;
                                }
                                val = ws.lens[32 + i - 1];
                                rep = 3 + br.peek(2);
                            } else if (sym == 17) {
                                rep = 3 + br.peek(3);
                            } else {
                                rep = 11 + br.peek(7);
                            }
                            br.drop(need);
                            if (i + (int)rep > nsym) {
                                bad = 2;  // "invalid bit length repeat"
                                break;
                            }
                            while (rep--) ws.lens[32 + i++] = (uint8_t)val;
                        }
                        if (!bad && ws.lens[32 + 256] == 0) bad = 2;  // "invalid code -- missing end-of-block"
                    }
                    bad = __shfl_sync(FULL, bad, 0);
                    __syncwarp();
                    if (!bad) {
                        // canonical places: lit/len at lens[0..hlit), dist at lens[288..288+hdist)
                        uint8_t tmp[10];
#pragma unroll
                        for (int k = 0; k < 10; ++k) {
                            const int i = lane + 32 * k;
                            tmp[k] = i < hlit + hdist ? ws.lens[32 + i] : 0;
                        }
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 10; ++k) {
                            const int i = lane + 32 * k;
                            if (i < hlit) ws.lens[i] = tmp[k];
                            else if (i < hlit + hdist) ws.lens[288 + (i - hlit)] = tmp[k];
                        }
                        __syncwarp();
                    }
                }
                if (!bad) {
                    int maxl = 0;
                    int st = build_table(ws.lens, hlit, ws.lit_count, ws.lit_sorted, ws.lit_fast, LIT_FAST_BITS, ws.tmp_base, ws.lit_fcode, ws.lit_fidx, lane, &maxl, ws.lit_lim);
                    // inftrees.c: over-subscribed never; incomplete only for a single 1-bit code
                    if (st == 1 || (st == 2 && maxl != 1)) bad = 2;  // "invalid literal/lengths set"
                    __syncwarp();
                    if (!bad) {
                        st = build_table(ws.lens + 288, hdist, ws.dist_count, ws.dist_sorted, ws.dist_fast, DIST_FAST_BITS, ws.tmp_base, ws.dist_fcode, ws.dist_fidx, lane, &maxl);
                        if (st == 1 || (st == 2 && maxl > 1)) bad = 2;  // "invalid distances set"
                    }
                    __syncwarp();
                }
                if (bad) {
                    ev = EV_END;
                    ev_a = bad == 3 ? END_TRUNC : END_ERR;
                } else {
                    in_block = true;
                    if (par_allowed && store) {
                        try_par = true;
                        par_pos = __shfl_sync(FULL, br.bitpos(), 0);
                        par_chunk = PAR_CHUNK_START_BITS;
                    }
                }
            }

            // ---- flush the staging buffer: complete 16-byte segments, everything at the end of the stream
            const bool final = ev == EV_END;
            if (ev == EV_FLUSH || final || (ev == EV_MATCH && spos + ev_a > INF_STAGE)) {
                const uint32_t nnew = spos - astart;
                if (store && total + nnew > ocap) store = false;  // slot overflow: only count from here on
                if (!store) {
                    total += nnew;
                    spos = astart = vstart = 0;
                } else {
                    adler_update(ad_a, ad_b, ws.stage + astart, nnew, lane);
                    total += nnew;
                    const uint32_t wseg = final ? (spos + 15) >> 4 : spos >> 4;
                    const uint4 *s4 = reinterpret_cast<const uint4 *>(ws.stage);
                    uint4 *g4 = reinterpret_cast<uint4 *>(gbase);
                    for (uint32_t seg = lane; seg < wseg; seg += 32) {
                        const uint32_t lo = seg * 16, hi = lo + 16;
                        if (lo >= vstart && hi <= spos) {
                            g4[seg] = s4[seg];
                        } else {  // ragged first / last segment of the stream
                            for (uint32_t i = max(lo, vstart); i < min(hi, spos); ++i) gbase[i] = ws.stage[i];
                        }
                    }
                    if (wseg) {
                        const uint32_t full = final ? spos : wseg * 16;
                        const uint32_t keep = spos - full;
                        uint8_t t = 0;
                        if (lane < (int)keep) t = ws.stage[full + lane];
                        __syncwarp();
                        if (lane < (int)keep) ws.stage[lane] = t;
                        gbase += wseg * 16;
                        vstart = 0;
                        astart = spos = keep;
                    } else {
                        astart = spos;  // fewer than 16 bytes so far: counted, still waiting to be stored
                    }
                    __syncwarp();
                }
            }
            if (ev == EV_MATCH) {
                const uint32_t mlen = ev_a, dist = ev_b;
                if (store) {
                    // out[pos+k] = out[pos-dist+(k mod dist)]; sources that already left the stage are read back
                    // from global memory through L2 (this warp stored them earlier in this launch)
                    const int64_t stage0 = (int64_t)total - (int64_t)astart;  // absolute position of stage[0]
                    const int64_t pos = stage0 + spos;
                    __threadfence_block();
                    for (uint32_t k = lane; k < mlen; k += 32) {
                        const int64_t sp = pos - dist + (k % dist);
                        const int64_t si = sp - stage0;
                        ws.stage[spos + k] = si >= (int64_t)vstart ? ws.stage[si] : __ldcg(dst + sp);
                    }
                    __syncwarp();
                }
                spos += mlen;
            }
            if (final) {
                stream_done = true;
                end_kind = (int32_t)ev_a;
                if (end_kind == END_OK && store && ((ad_b << 16) | ad_a) != ev_b) end_kind = END_ERR;  // "incorrect data check"
            }
        }
        // ---- epilogue
        if (lane == 0) {
            br.drain();
            phase_bits = br.phase_bits;
        }
        phase_bits = __shfl_sync(FULL, phase_bits, 0);
        __syncwarp();
        if (lane == 0) {
            int32_t st = S5B_OK;
            if (end_kind == END_ERR) st = S5B_ERR_PRESS;
            else if (!store) st = S5B_ERR_NOSPACE;
            a.status[r] = st;
            // on overflow out_len reports the size the stream needs (the caller retries with a larger slot)
            a.out_len[r] = end_kind == END_ERR ? 0u : (uint32_t)total;
        }
    }
}

int inflate_blocks_per_sm() {
    int n = 0;
    if (cudaFuncSetAttribute(inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(InfWarpSmem) * INF_WARPS)) != cudaSuccess)
        return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, inflate_kernel, INF_WARPS * 32,
                                                      sizeof(InfWarpSmem) * INF_WARPS) != cudaSuccess)
        return 0;
    return n;
}

// Streams of at most INF_THREAD_MAX_LEN bytes are decoded one per THREAD (inflate_thread_kernels.cu), longer ones one per
// warp with the speculative window decoder above; both kernels sweep the whole batch and skip the other's streams.
// S5B_INFLATE_THREAD_MAX (bytes, read once) moves the boundary; 0 sends everything to the warp kernel.
static uint32_t inflate_thread_max_len() {
    static const uint32_t v = [] {
        const char *e = getenv("S5B_INFLATE_THREAD_MAX");
        return e ? (uint32_t)strtoul(e, nullptr, 10) : 16384u;
    }();
    return v;
}

cudaError_t launch_inflate(const InflateArgs &a, int num_sms, int blocks_per_sm, cudaStream_t st) {
    const uint32_t tmax = inflate_thread_max_len();
    if (tmax) {
        cudaError_t e0 = launch_inflate_threads(a, tmax, num_sms, st);
        if (e0 != cudaSuccess) return e0;
    }
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t want = (a.n_reads + INF_WARPS - 1) / INF_WARPS;
    uint64_t cap = (uint64_t)num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (!grid) grid = 1;
    inflate_kernel<<<grid, INF_WARPS * 32, sizeof(InfWarpSmem) * INF_WARPS, st>>>(a, tmax ? tmax + 1 : 0u);
    return cudaGetLastError();
}

}  // namespace s5b
