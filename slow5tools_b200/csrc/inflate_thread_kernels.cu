// inflate_thread_kernels.cu -- sm_100a zlib/DEFLATE decoder for SHORT streams: one THREAD per record stream.
//
// BLOW5 records are independent zlib streams of a few KiB (slow5.c:4046; 4096 samples -> ~3.5 KiB).  A bit stream is
// serial by nature; the warp-per-stream decoder (inflate_kernels.cu) buys parallelism inside a stream with speculative
// window decoding and pays for it with three decodes of every symbol plus divergence (60 k warp instructions per record).
// A batch has 10^5..10^6 streams, so the parallelism is there without speculation: here every lane runs a serial decoder
// on its own stream, 32 streams per warp.  Two things make that work on a GPU:
//   * LOCK STEP.  A warp takes 32 consecutive records and all lanes move through the same loops -- block header and code
//     construction in phases, then up to three literals (or one other symbol, or up to four match bytes, or one stored
//     byte) per lane and iteration.
//     Every loop is steered by a warp-wide vote or has a fixed trip count, which is also what keeps the 32 decoders
//     converged: left to themselves, 32 data-dependent decoders diverge for good and independent thread scheduling
//     runs them one after the other (measured: 12x slower than the warp-per-stream kernel).
//   * NO LOOKUP TABLES.  A first-level table per lane (512 B + overflow lists) fills shared memory at 4 warps per SM and
//     the dependent load-shift-load chain of a serial decoder then leaves the SM idle 85 % of the time.  Instead a symbol
//     is decoded canonically: the next 15 bits, MSB first, are compared against the per-length code limits held in
//     REGISTERS (a four-level select tree, ~20 instructions, no memory access), and the symbol comes from the list of
//     symbols sorted by (length, symbol) that every lane keeps in a global-memory scratch row (L1 resident: the frequent
//     symbols are its first bytes).  Per lane that leaves 160 bytes of shared memory (output ring, code bases), so the
//     kernel runs at register-limited occupancy and hides its latencies with other warps.
// Streams longer than the caller's threshold stay with the warp-per-stream kernel (a serial decode of a 200 KiB stream
// would be a long tail for the 31 lanes next to it).
//
// Observable behaviour is that of ptr_depress_zlib_solo (slow5lib/src/slow5_press.c:973-1010) as mirrored by
// inflate_kernels.cu: RFC 1950 header check, stored / fixed / dynamic blocks, 32 KiB window, Adler-32 check; malformed
// data -> S5B_ERR_PRESS; input that ends early is not an error and yields the bytes decoded so far; a slot that is too
// small -> S5B_ERR_NOSPACE with the size needed in out_len.
#include "s5b_kernels.h"
#include "s5b_ptx.cuh"
#include "../../include/slow5b200.h"

namespace s5b {

namespace inft {

// Measured on B200: 22 warps per SM.  Fewer hide too little of the serial decoders' latency, more shrink the L1 (shared memory
// grows) and cost registers (the launch bound forces spills from 11 CTAs on).  The first sweep (profiles/
// r2_inflate_thread_variants.md) was taken on 262 144-record chunks, where 18..24 warps per SM all need three 32-record rounds per
// chunk and 20 looked best; on one 1 M-record chunk (tools/dev/ti_ctas_sweep.sh, profiles/r2_ti_ctas_sweep.txt) the records per
// round and the time of a round give 9: 29.7, 10: 31.4, 11: 33.0, 12: 32.6 M records/s.
#ifndef S5B_TI_CTAS
#define S5B_TI_CTAS 11
#endif
#ifndef S5B_TI_RING
#define S5B_TI_RING 64
#endif
#ifndef S5B_TI_HOT
#define S5B_TI_HOT 160
#endif
constexpr int TI_WARPS = 2;
constexpr int TI_CTAS_PER_SM = S5B_TI_CTAS;  // x 2 warps per SM
constexpr int RING = S5B_TI_RING;            // per-lane output ring (bytes): staging for 16-byte stores, source of near matches
constexpr int FLUSH_EVERY = RING / 8;        // iterations; a lane produces <= 4 bytes per iteration: RING / 2 + 15 unflushed < RING
constexpr uint32_t ADLER_MOD = 65521u;

// The deflate encoder of this library codes the front part of every record under one code fixed ahead of time
// (deflate_kernels.cu): such a block's header is always the same 480 bits.  A lane that finds exactly those bits behind
// BFINAL / BTYPE loads the finished decode tables below instead of reading the header and building them (one of the two
// blocks of every record this library wrote).  Any other header -- every stream zlib wrote -- goes the general way.
#include "inflate_canned.inc"

constexpr int HOT = S5B_TI_HOT;  // sorted symbols kept in shared memory (canonical order puts the frequent ones first)
// per-lane shared memory: what the symbol loop touches on every iteration; the code lengths of a block header are parked
// on the same bytes (as nibbles) while the block's codes are being built
struct __align__(4) TiSmem {
    uint8_t ring[RING];
    union {
        uint8_t nib[160];      // code lengths, two per byte: literal/length at 0..287, distance at 288..319
        uint8_t sorted8[HOT];  // after construction: low 8 bits of the first HOT literal/length symbols sorted by (length, symbol)
    };
    // After construction, as 32-bit words indexed by code length l: low half = sorted index of the first code of length l
    // minus that code (mod 2^16), high half = sorted index of the first symbol >= 256 of that length (within one length
    // the symbols ascend, so "literal" is one compare): sym index = (lo + code) & 0xffff, literal iff index < hi.
    // During construction: entries 0..15 are the builders' scratch, bytes 32..63 the header's staging area.
    uint16_t lbn[32];
    // The 32 lanes of a warp touch the same member at about the same index at the same time (lock step).  With a stride of 72
    // words lanes l and l + 4 met in one bank (8-way conflicts on every ring store and table load); 73 words is odd: the 32
    // lanes land in 32 different banks.
    uint32_t bank_skew;
    __device__ __forceinline__ uint8_t *tmp() { return reinterpret_cast<uint8_t *>(lbn + 16); }
    __device__ __forceinline__ uint32_t len_at(int i) const { return (nib[i >> 1] >> ((i & 1) * 4)) & 15u; }
    __device__ __forceinline__ void set_len(int i, uint32_t v) {
        const uint32_t b = nib[i >> 1];
        nib[i >> 1] = (uint8_t)((i & 1) ? ((b & 0x0fu) | (v << 4)) : ((b & 0xf0u) | v));
    }
};
static_assert(HOT <= 160, "sorted8 lives on the parked code lengths");
static_assert((sizeof(TiSmem) / 4) % 2 == 1 && sizeof(TiSmem) % 4 == 0, "per-lane stride must be an odd number of words");
// per-lane global scratch row: code construction output and the rarely used codes
struct TiScratch {
    uint16_t lit_sorted[288];
    uint16_t aux_sorted[32];
    uint16_t lit_lim[16];    // end of the codes of each length, left-aligned to 15 bits
    int16_t aux_base[16];    // base / lim of the code-length code (while a header is read) / the distance code (after)
    uint16_t aux_lim[16];
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                        35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                         1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                                         9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum : int { END_OK = 0, END_TRUNC = 1, END_ERR = 2 };
enum : int { ST_DONE = 0, ST_BLOCK = 1, ST_SYM = 2, ST_MATCH = 3, ST_STORED = 4 };

// ---- input side of one lane ---------------------------------------------------------------------------------------
// Two aligned 32-bit words of the stream and a bit offset into the first: the next 32 bits are one funnel shift, taking
// n bits is an add plus a word shuffle.  `avail` counts the stream bits not yet consumed, so truncation is one compare.
// Words that hold no stream byte are never loaded.
struct Bits {
    const uint2 *base; // aligned pair of words holding the first stream byte
    uint32_t pi, np;   // next pair to load, number of pairs that hold stream bytes
    uint32_t w0, w1;   // the current window: the next 32 bits are one funnel shift
    uint32_t q0;       // the word behind the window
    // The words behind q0 arrive two at a time (one 64-bit load per 8 bytes of stream: the 32 lanes read 32 different lines,
    // so every load instruction is up to 32 trips through the L1) into two 64-bit registers that take turns: q0 is fed from
    // pa (low word, high word), then from pb, and a pair is requested again the moment its high word has been taken -- into
    // the same registers, first touched two window shifts later.  (Every formulation in which the loaded value had to be
    // moved -- unpacked into 32-bit registers, or handed from a "pending" to a "next" variable -- made the register
    // allocator load into a temporary and copy it right behind the LDG, which waits for the load on the spot: 15 % of the
    // kernel's stall samples.)
    unsigned long long pa, pb;
    uint32_t cnt;      // which word q0 takes next: 0 pa.lo, 1 pa.hi, 2 pb.lo, 3 pb.hi
    uint32_t off;      // consumed bits of w0 (0..31)
    uint32_t avail;    // stream bits left (from the current position)
    // the next pair of the stream into v.  Behind the end of the stream the last pair is read again: those bits are never
    // taken for stream bits (`avail`), and an unconditional load keeps the compiler from wrapping it in a branch.
    __device__ __forceinline__ void ld(unsigned long long &v) {
        const uint2 *p = base + min(pi, np - 1u);
        asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
        ++pi;
    }
    // position the reader `bit` bits into the stream of `len` bytes at p (bit <= 8 len)
    __device__ __forceinline__ void start(const uint8_t *p, uint32_t len, uint32_t bit = 0) {
        const uint32_t sk = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 7u);
        base = reinterpret_cast<const uint2 *>(p - sk);
        np = (sk + len + 7u) >> 3;
        const uint32_t b = sk * 8u + bit;
        pi = b >> 6;
        unsigned long long a;
        ld(a);
        ld(pa);
        ld(pb);
        w0 = (uint32_t)a, w1 = (uint32_t)(a >> 32), q0 = (uint32_t)pa;
        cnt = 1;
        off = b & 63u;
        avail = len * 8u - bit;
        if (off >= 32) {  // the position is in the second word of its pair
            off -= 32;
            shift();
        }
    }
    // one word leaves the window
    __device__ __forceinline__ void shift() {
        w0 = w1;
        w1 = q0;
        const unsigned long long x = (cnt & 2u) ? pb : pa;
        q0 = (cnt & 1u) ? (uint32_t)(x >> 32) : (uint32_t)x;
        if (cnt == 1u) ld(pa);
        if (cnt == 3u) ld(pb);
        cnt = (cnt + 1u) & 3u;
    }
    __device__ __forceinline__ uint32_t peek32() const { return __funnelshift_r(w0, w1, off); }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return peek32() & ((1u << n) - 1u); }
    // the 32 bits that start `ahead` (<= 32) bits behind the current position: they may begin in the second word
    __device__ __forceinline__ uint32_t peek32_at(uint32_t ahead) const {
        const uint32_t o = off + ahead;  // < 64
        return __funnelshift_r(o < 32 ? w0 : w1, o < 32 ? w1 : q0, o);  // (the shift amount is taken mod 32)
    }
    // n <= 64 and n <= avail
    __device__ __forceinline__ void drop_long(uint32_t n) {
        off += n;
        avail -= n;
        if (off >= 32) {
            off -= 32;
            shift();
            if (off >= 32) {
                off -= 32;
                shift();
            }
        }
    }
    // n <= 32 and n <= avail
    __device__ __forceinline__ void drop(uint32_t n) {
        off += n;
        avail -= n;
        if (off >= 32) {
            off -= 32;
            shift();
        }
    }
    // to the next byte boundary of the stream (the padding bits are there whenever a whole byte follows)
    __device__ __forceinline__ void align_byte() {
        const uint32_t pad = (8u - (off & 7u)) & 7u;
        if (pad <= avail) drop(pad);
        else avail = 0;
    }
};

// ---- output side of one lane ----------------------------------------------------------------------------------------
// Bytes go into the ring; complete 16-byte segments of the destination (absolute alignment) leave as one 128-bit store.
// Flushing is driven by the warp (every FLUSH_EVERY iterations of the lock-step loop all lanes store what they have), so
// that the store + Adler-32 code runs once for 32 lanes instead of on whichever lane crosses a 16-byte boundary.
struct Out {
    uint8_t *ring;
    uint8_t *dst;        // slot base
    uint32_t cap;        // slot bytes
    uint32_t total;      // bytes produced
    uint32_t flushed;    // bytes [0, flushed) have been stored
    uint32_t ad_a, ad_b;
    uint32_t bias;       // output byte i lives at ring[(i + bias) & (RING - 1)]: 16-byte segments of the destination are
                         // 16-byte segments of the ring
    bool store;
    // the ring takes every byte; whether it may be stored is decided when it is flushed (a slot that is too small only
    // needs the final count)
    __device__ __forceinline__ void put(uint32_t byte) {
        ring[(total + bias) & (RING - 1)] = (uint8_t)byte;
        ++total;
    }
    __device__ __forceinline__ void flush_bytes(uint32_t end) {
        for (uint32_t i = flushed; i < end; ++i) {
            const uint32_t d = ring[(i + bias) & (RING - 1)];
            dst[i] = (uint8_t)d;
            ad_a += d;
            ad_b += ad_a;
        }
        ad_a %= ADLER_MOD;
        ad_b %= ADLER_MOD;
        flushed = end;
    }
    // store every complete 16-byte segment (all == false) or everything (all == true)
    __device__ __forceinline__ void flush(bool all) {
        if (total > cap) store = false;
        if (!store) return;
        for (;;) {
            const uint32_t seg_end = ((flushed + bias) | 15u) + 1u - bias;  // end of the destination segment holding `flushed`
            if (seg_end > total) break;
            if (seg_end - flushed == 16) {
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(ring + ((flushed + bias) & (RING - 1)));
                const uint4 v = make_uint4(rw[0], rw[1], rw[2], rw[3]);
                *reinterpret_cast<uint4 *>(dst + flushed) = v;
                // Adler-32 over 16 bytes: b += 16 a + sum (16 - i) d_i, a += sum d_i
                const uint32_t sum = __dp4a(v.x, 0x01010101u, 0u) + __dp4a(v.y, 0x01010101u, 0u) + __dp4a(v.z, 0x01010101u, 0u) +
                                     __dp4a(v.w, 0x01010101u, 0u);
                const uint32_t wsum = __dp4a(v.x, 0x0d0e0f10u, 0u) + __dp4a(v.y, 0x090a0b0cu, 0u) + __dp4a(v.z, 0x05060708u, 0u) +
                                      __dp4a(v.w, 0x01020304u, 0u);
                ad_b = (ad_b + 16u * ad_a + wsum) % ADLER_MOD;
                ad_a = (ad_a + sum) % ADLER_MOD;
                flushed = seg_end;
            } else {
                flush_bytes(seg_end);  // ragged first segment of the slot
            }
        }
        if (all && total > flushed) flush_bytes(total);
    }
    // one byte of a match: out[total] = out[total - dist].  Bytes not yet stored are in the ring (it holds the last RING
    // bytes and the unflushed tail is shorter); everything older has been stored by this thread.
    __device__ __forceinline__ void copy1(uint32_t dist) {
        const uint32_t from = total - dist;
        put(from >= flushed || !store ? (uint32_t)ring[(from + bias) & (RING - 1)] : (uint32_t)dst[from]);
    }
};

// ---- canonical codes ------------------------------------------------------------------------------------------------
// From lens[0..n): lim[l] = end of the codes of length l left-aligned to 15 bits (non-decreasing in l), base[l] = sorted
// index of the first code of length l minus that code, sorted[] = symbols by (length, symbol).  All loops have fixed trip
// counts (NMAX) and `active` lanes are merely predicated, so the warp stays converged through the call.
// Returns 0 ok, 1 over-subscribed, 2 incomplete (zlib's inflate_table rules decide what that means); *max_len too.
template <int NMAX, int STRIDE, typename LenAt>
__device__ __forceinline__ int build_code(bool active, LenAt lens, int n, uint16_t *lim, int16_t *base, uint16_t *next,
                                          uint16_t *sorted, int *max_len) {
    // base[l * STRIDE], next[l * STRIDE]: the literal/length code interleaves the two (TiSmem::lbn)
    if (active) {
#pragma unroll
        for (int l = 0; l < 16; ++l) next[l * STRIDE] = 0;  // used as the per-length count first
    }
    for (int s = 0; s < NMAX; ++s) {
        if (active && s < n) {
            const int l = (int)lens(s);
            if (l) ++next[l * STRIDE];
        }
    }
    int left = 1, status = 0, maxl = 0;
    if (active) {
        uint32_t code = 0, at = 0;
#pragma unroll
        for (int l = 1; l <= 15; ++l) {
            const uint32_t c = next[l * STRIDE];
            left <<= 1;
            left -= (int)c;
            if (left < 0) status = 1;
            if (c) maxl = l;
            base[l * STRIDE] = (int16_t)((int)at - (int)code);
            lim[l] = (uint16_t)min((code + c) << (15 - l), 0xffffu);
            next[l * STRIDE] = (uint16_t)at;
            at += c;
            code = (code + c) << 1;
        }
        if (status == 0 && left > 0) status = 2;
    }
    for (int s = 0; s < NMAX; ++s) {
        if (active && status != 1 && s < n) {
            const int l = (int)lens(s);
            if (l) sorted[next[l * STRIDE]++] = (uint16_t)s;
        }
    }
    *max_len = maxl;
    return status;
}

// generic canonical decode from the arrays (code-length and distance codes): returns the symbol, -1 (no such code) or -2
// (the stream ends inside the code); *used = code length
__device__ __forceinline__ int canon_decode(const Bits &b, const uint16_t *lim, const int16_t *base, const uint16_t *sorted,
                                            int max_len, uint32_t *used) {
    const uint32_t c15 = __brev(b.peek32()) >> 17;
    int l = 1;
    while (l <= max_len && c15 >= lim[l]) ++l;
    if (l > max_len) return b.avail < 15u ? -2 : -1;
    if ((uint32_t)l > b.avail) return -2;
    *used = (uint32_t)l;
    return sorted[base[l] + (int)(c15 >> (15 - l))];
}

// do the CANNED_HDR_BITS bits that start `bit` bits into the stream at p (len bytes, all of them inside it) equal the canned header?
__device__ __forceinline__ bool canned_header_at(const uint8_t *p, uint32_t len, uint32_t bit) {
    const uint32_t sk = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - sk);
    const uint32_t last = (sk + len - 1u) >> 2;  // last word that holds a stream byte
    const uint32_t b = sk * 8u + bit;
    const uint32_t sh = b & 31u;
    uint32_t i = b >> 5;
    uint32_t lo = __ldg(w + i);
    bool same = true;
#pragma unroll 5
    for (int q = 0; q < CANNED_HDR_WORDS; ++q) {
        ++i;
        const uint32_t hi = __ldg(w + min(i, last));
        same = same && __funnelshift_r(lo, hi, sh) == g_canned_hdr[q];
        lo = hi;
    }
    return same;
}

__global__ void __launch_bounds__(TI_WARPS * 32, TI_CTAS_PER_SM) inflate_thread_kernel(const InflateArgs a, uint32_t max_len,
                                                                                        TiScratch *scratch_rows) {
    __shared__ TiSmem smem[TI_WARPS * 32];
    TiSmem &sm = smem[threadIdx.x];
    TiScratch &sc = scratch_rows[(size_t)blockIdx.x * (TI_WARPS * 32) + threadIdx.x];
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long rbase = 0;
        if (lane == 0) rbase = atomicAdd(a.work_counter, 32ULL);
        rbase = __shfl_sync(FULL, rbase, 0);
        if (rbase >= a.n_reads) break;
        const unsigned long long r = rbase + lane;
        uint32_t ilen = 0;
        bool mine = false;
        if (r < a.n_reads) {
            ilen = a.in_len[r];
            mine = ilen <= max_len;  // longer streams: the warp-per-stream kernel takes them
        }
        Bits in;
        Out out;
        int end_kind = -1;  // still decoding
        int state = ST_DONE;
        bool last_block = false;
        uint32_t aux = 0, m_dist = 0;  // stored bytes left / match bytes left, match distance
        int dist_max = 0;              // longest distance code of the current block
        const uint8_t *src_p = a.in;   // first byte of the lane's stream
        const uint16_t *rare = sc.lit_sorted;  // sorted symbols beyond the ones in shared memory: the lane's own row, or the canned table
        // the literal/length code limits of the current block, one register per length
        uint32_t L1 = 0, L2 = 0, L3 = 0, L4 = 0, L5 = 0, L6 = 0, L7 = 0, L8 = 0, L9 = 0, L10 = 0, L11 = 0, L12 = 0, L13 = 0,
                 L14 = 0, L15 = 0;
        in.base = reinterpret_cast<const uint2 *>(a.in);
        in.pi = in.np = in.w0 = in.w1 = in.q0 = in.off = in.avail = 0;
        in.pa = in.pb = 0;
        in.cnt = 0;
        in.np = 1;
        out.ring = sm.ring;
        out.dst = a.out;
        out.cap = out.total = out.flushed = out.bias = 0;
        out.ad_a = 1;
        out.ad_b = 0;
        out.store = true;
        if (mine) {
            const uint64_t ioff = a.in_off[r];
            if (ioff + ilen > a.in_capacity) {
                a.status[r] = S5B_ERR_ARG;
                a.out_len[r] = 0;
                mine = false;
            } else {
                src_p = a.in + ioff;
                in.start(src_p, ilen);
                out.dst = a.out + a.out_off[r];
                const uint64_t cap64 = a.out_off[r + 1] - a.out_off[r];
                out.cap = cap64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)cap64;
                out.bias = (uint32_t)(reinterpret_cast<uintptr_t>(out.dst) & 15u);
                // ---- zlib header (RFC 1950)
                if (in.avail < 16) {
                    end_kind = END_TRUNC;
                } else {
                    const uint32_t cmf = in.peek(8), flg = in.peek(16) >> 8;
                    in.drop(16);
                    if (((cmf << 8) | flg) % 31u != 0 || (cmf & 15u) != 8 || (cmf >> 4) > 7 || (flg & 0x20u)) end_kind = END_ERR;
                }
                state = end_kind < 0 ? ST_BLOCK : ST_DONE;
            }
        }
        while (__any_sync(FULL, state != ST_DONE)) {
            // ================= block boundary, in phases that all 32 lanes walk through together =================
            const bool hb = state == ST_BLOCK;
            bool dynamic = false, tables = false, canned = false;
            int hlit = 288, hdist = 32;
            uint32_t hclen = 0;
            // ---- phase 1: trailer after the final block, else the 3 header bits and what follows them directly
            if (hb) {
                state = ST_DONE;  // every early exit below ends the stream
                if (last_block) {
                    // Adler-32 trailer: big endian, byte aligned, after the final block
                    in.align_byte();
                    if (in.avail < 32) {
                        end_kind = END_TRUNC;
                    } else {
                        const uint32_t want = __byte_perm(in.peek32(), 0, 0x0123);
                        out.flush(true);
                        end_kind = (!out.store || ((out.ad_b << 16) | out.ad_a) == want) ? END_OK : END_ERR;  // "incorrect data check"
                    }
                } else if (in.avail < 3) {
                    end_kind = END_TRUNC;
                } else {
                    last_block = in.peek(1);
                    const uint32_t btype = in.peek(3) >> 1;
                    in.drop(3);
                    if (btype == 3) {  // "invalid block type"
                        end_kind = END_ERR;
                    } else if (btype == 0) {
                        in.align_byte();
                        if (in.avail < 32) {
                            end_kind = END_TRUNC;
                        } else {
                            const uint32_t v = in.peek32();
                            in.drop(32);
                            aux = v & 0xffffu;
                            if ((aux ^ 0xffffu) != (v >> 16)) end_kind = END_ERR;  // "invalid stored block lengths"
                            else state = aux ? ST_STORED : ST_BLOCK;
                        }
                    } else if (btype == 1) {
                        tables = true;
                    } else if (in.avail < 14) {
                        end_kind = END_TRUNC;
                    } else if (in.avail >= CANNED_HDR_BITS && canned_header_at(src_p, ilen, ilen * 8u - in.avail)) {
                        canned = true;
                    } else {
                        hlit = (int)in.peek(5) + 257;
                        hdist = (int)(in.peek(10) >> 5) + 1;
                        hclen = (in.peek(14) >> 10) + 4;
                        in.drop(14);
                        if (hlit > 286 || hdist > 30) end_kind = END_ERR;  // "too many length or distance symbols"
                        else dynamic = true;
                    }
                }
            }
            if (__any_sync(FULL, canned)) {
                // ---- the encoder's canned block: finished tables instead of phases 2 to 5
                if (canned) {
                    in.start(src_p, ilen, ilen * 8u - in.avail + CANNED_HDR_BITS);
                    uint32_t *lbnw = reinterpret_cast<uint32_t *>(sm.lbn);
#pragma unroll
                    for (int q = 0; q < 16; ++q) lbnw[q] = g_canned_lbn[q];
                    uint32_t *s8w = reinterpret_cast<uint32_t *>(sm.sorted8);
                    for (int q = 0; q < HOT / 4; ++q) s8w[q] = g_canned_sorted8[q];
                    L1 = g_canned_L[1], L2 = g_canned_L[2], L3 = g_canned_L[3], L4 = g_canned_L[4], L5 = g_canned_L[5];
                    L6 = g_canned_L[6], L7 = g_canned_L[7], L8 = g_canned_L[8], L9 = g_canned_L[9], L10 = g_canned_L[10];
                    L11 = g_canned_L[11], L12 = g_canned_L[12], L13 = g_canned_L[13], L14 = g_canned_L[14], L15 = g_canned_L[15];
                    // one distance code (distance 1), one bit long
                    sc.aux_lim[1] = 0x4000;
                    sc.aux_base[1] = 0;
                    sc.aux_sorted[0] = 0;
                    dist_max = 1;
                    rare = g_canned_sorted;
                    state = ST_SYM;
                }
            }
            if (__any_sync(FULL, tables)) {
                // fixed code (btype 1): lengths 8/9/7/8, all 32 five-bit distance codes (30, 31: "invalid distance code")
                for (int s = 0; s < 320; ++s)
                    if (tables) sm.set_len(s, s < 144 ? 8u : s < 256 ? 9u : s < 280 ? 7u : s < 288 ? 8u : 5u);
            }
            if (__any_sync(FULL, dynamic)) {
                // ---- phase 2: the code-length code (3 bits per length, fixed order), its canonical arrays
                for (int i = 0; i < 19; ++i) {
                    if (dynamic) {
                        uint32_t v = 0;
                        if ((uint32_t)i < hclen) {
                            if (in.avail < 3) {
                                end_kind = END_TRUNC;
                                dynamic = false;
                            } else {
                                v = in.peek(3);
                                in.drop(3);
                            }
                        }
                        sm.tmp()[c_cl_order[i]] = (uint8_t)v;
                    }
                }
                int cl_max = 0;
                // "invalid code lengths set": the code-length code must be complete (inftrees.c, type CODES)
                if (build_code<19, 1>(dynamic, [&](int q) { return (uint32_t)sm.tmp()[q]; }, 19, sc.aux_lim, sc.aux_base, sm.lbn,
                                   sc.aux_sorted, &cl_max) != 0 && dynamic) {
                    end_kind = END_ERR;
                    dynamic = false;
                }
                // ---- phase 3: the hlit + hdist code lengths, run-length coded; one length per lane and iteration.  The limits
                // of the (at most 7-bit) code-length code sit in registers for the duration.
                uint32_t C1 = 0, C2 = 0, C3 = 0, C4 = 0, C5 = 0, C6 = 0, C7 = 0;
                int cb1 = 0, cb2 = 0, cb3 = 0, cb4 = 0, cb5 = 0, cb6 = 0, cb7 = 0;
                if (dynamic) {
                    C1 = sc.aux_lim[1], C2 = sc.aux_lim[2], C3 = sc.aux_lim[3], C4 = sc.aux_lim[4], C5 = sc.aux_lim[5];
                    C6 = sc.aux_lim[6], C7 = sc.aux_lim[7];
                    cb1 = sc.aux_base[1], cb2 = sc.aux_base[2], cb3 = sc.aux_base[3], cb4 = sc.aux_base[4], cb5 = sc.aux_base[5];
                    cb6 = sc.aux_base[6], cb7 = sc.aux_base[7];
                }
                const int nsym = hlit + hdist;
                int i = 0;
                uint32_t rep = 0, val = 0;
                while (__any_sync(FULL, dynamic && i < nsym)) {
                    if (dynamic && i < nsym) {
                        if (rep) {
                            sm.set_len(i++, val);
                            --rep;
                        } else {
                            const uint32_t c15 = __brev(in.peek32()) >> 17;
                            const uint32_t l = 1u + (c15 >= C1) + (c15 >= C2) + (c15 >= C3) + (c15 >= C4) + (c15 >= C5) + (c15 >= C6) +
                                               (c15 >= C7);
                            int sym = -1;
                            if (l <= 7u && l <= in.avail) {
                                const int cb = l == 1 ? cb1 : l == 2 ? cb2 : l == 3 ? cb3 : l == 4 ? cb4 : l == 5 ? cb5 : l == 6 ? cb6 : cb7;
                                sym = sc.aux_sorted[cb + (int)(c15 >> (15u - l))];
                            }
                            if (sym < 0) {  // complete code: a miss can only mean the bits ran out
                                end_kind = END_TRUNC;
                                dynamic = false;
                            } else if (sym < 16) {
                                in.drop(l);
                                val = (uint32_t)sym;
                                sm.set_len(i++, (uint32_t)sym);
                            } else {
                                const uint32_t need = sym == 16 ? 2 : sym == 17 ? 3 : 7;
                                if (in.avail < l + need) {
                                    end_kind = END_TRUNC;
                                    dynamic = false;
                                } else {
                                    in.drop(l);
                                    if (sym == 16) {
                                        if (i == 0) {  // "invalid bit length repeat"
                                            end_kind = END_ERR;
                                            dynamic = false;
                                        }
                                        rep = 3 + in.peek(2);  // val = the previous length
                                    } else {
                                        val = 0;
                                        rep = (sym == 17 ? 3 : 11) + in.peek(need);
                                    }
                                    in.drop(need);
                                    if (dynamic && i + (int)rep > nsym) {  // "invalid bit length repeat"
                                        end_kind = END_ERR;
                                        dynamic = false;
                                    }
                                }
                            }
                        }
                    }
                }
                // ---- phase 4: distance lengths to their place, end-of-block check
                if (dynamic && sm.len_at(256) == 0) {  // "invalid code -- missing end-of-block"
                    end_kind = END_ERR;
                    dynamic = false;
                }
                for (int s = 0; s < 32; ++s) {
                    if (dynamic) sm.tmp()[s] = s < hdist ? (uint8_t)sm.len_at(hlit + s) : (uint8_t)0;
                }
                for (int s = 0; s < 32; ++s) {
                    if (dynamic) sm.set_len(288 + s, sm.tmp()[s]);
                }
                if (dynamic) tables = true;
            }
            if (__any_sync(FULL, tables)) {
                // ---- phase 5: the two codes of the block
                int maxl = 0;
                int st = build_code<32, 1>(tables, [&](int q) { return sm.len_at(288 + q); }, hdist, sc.aux_lim, sc.aux_base, sm.lbn,
                                        sc.aux_sorted, &maxl);
                if (tables && (st == 1 || (st == 2 && maxl > 1))) {  // "invalid distances set"
                    end_kind = END_ERR;
                    tables = false;
                }
                if (tables) dist_max = maxl;  // (lanes that loaded canned tables in this round keep theirs)
                st = build_code<288, 2>(tables, [&](int q) { return sm.len_at(q); }, hlit, sc.lit_lim,
                                        reinterpret_cast<int16_t *>(sm.lbn), sm.lbn + 1, sc.lit_sorted, &maxl);
                // inftrees.c: over-subscribed never; incomplete only for a single 1-bit code
                if (tables && (st == 1 || (st == 2 && maxl != 1))) {  // "invalid literal/lengths set"
                    end_kind = END_ERR;
                    tables = false;
                }
                // lbn[2l + 1] is now the end of length l's symbols; stepping back over its symbols >= 256 (end of block, the
                // match lengths) leaves the index of the first of them
                for (int q = 256; q < 288; ++q) {
                    if (tables && q < hlit) {
                        const uint32_t l = sm.len_at(q);
                        if (l) --sm.lbn[2 * l + 1];
                    }
                }
                // the frequent symbols move next to the decoder (the code lengths parked on these bytes are done with)
                for (int i = 0; i < HOT; i += 2) {
                    if (tables) {
                        const uint32_t v = *reinterpret_cast<const uint32_t *>(&sc.lit_sorted[i]);
                        *reinterpret_cast<uint16_t *>(&sm.sorted8[i]) = (uint16_t)__byte_perm(v, 0u, 0x4420);
                    }
                }
                if (tables) {
                    // limit of length l, left-aligned to 15 bits, with l in the low 4 bits: one register names both
                    L1 = sc.lit_lim[1] << 4 | 1u, L2 = sc.lit_lim[2] << 4 | 2u, L3 = sc.lit_lim[3] << 4 | 3u;
                    L4 = sc.lit_lim[4] << 4 | 4u, L5 = sc.lit_lim[5] << 4 | 5u, L6 = sc.lit_lim[6] << 4 | 6u;
                    L7 = sc.lit_lim[7] << 4 | 7u, L8 = sc.lit_lim[8] << 4 | 8u, L9 = sc.lit_lim[9] << 4 | 9u;
                    L10 = sc.lit_lim[10] << 4 | 10u, L11 = sc.lit_lim[11] << 4 | 11u, L12 = sc.lit_lim[12] << 4 | 12u;
                    L13 = sc.lit_lim[13] << 4 | 13u, L14 = sc.lit_lim[14] << 4 | 14u, L15 = sc.lit_lim[15] << 4 | 15u;
                    rare = sc.lit_sorted;
                    state = ST_SYM;
                }
            }
            __syncwarp();
            // ================= symbols: one step per lane and iteration until every lane has left its block =================
            uint32_t iter = 0;
            const uint32_t *lbn32 = reinterpret_cast<const uint32_t *>(sm.lbn);
            // cq = the next 15 stream bits, MSB first, << 4 | 15 (compares like the bare 15 bits against limit << 4 | length)
            auto lit_len = [&](uint32_t cq) -> uint32_t {
                const bool s8 = cq >= L8;
                const bool s4 = cq >= (s8 ? L12 : L4);
                const bool s2 = cq >= (s8 ? (s4 ? L14 : L10) : (s4 ? L6 : L2));
                const uint32_t m1 = s8 ? (s4 ? (s2 ? L15 : L13) : (s2 ? L11 : L9)) : (s4 ? (s2 ? L7 : L5) : (s2 ? L3 : L1));
                return (m1 & 15u) + (cq >= m1 ? 1u : 0u);
            };
            while (__any_sync(FULL, state >= ST_SYM)) {
                if (state == ST_SYM) {
                    // canonical decode: find the first limit above the next 15 bits (binary search over registers; the
                    // register found carries its length)
                    const uint32_t w32 = in.peek32();
                    uint32_t l = lit_len((__brev(w32) >> 13) | 15u);
                    if (l > 15u || l > in.avail) {
                        // no such code ("invalid literal/length code" when 15 bits were there), or the stream ends inside it
                        end_kind = (l > 15u && in.avail >= 15u) ? END_ERR : END_TRUNC;
                        state = ST_DONE;
                    } else {
                        const uint32_t lw = lbn32[l];
                        const uint32_t si = (lw + (__brev(w32) >> (32u - l))) & 0xffffu;
                        const uint32_t lo8 = si < (uint32_t)HOT ? (uint32_t)sm.sorted8[si] : (uint32_t)(rare[si] & 0xffu);
                        if (si < (lw >> 16)) {
                            // a literal -- and, most of the time, another one behind it: the 32-bit window holds two codes,
                            // so the second is decoded on the spot (anything else is left for the next iteration)
                            out.put(lo8);
                            const uint32_t r2 = __brev(w32 >> l);
                            const uint32_t l2 = lit_len((r2 >> 13) | 15u);
                            if (l2 <= 15u && l + l2 <= in.avail) {
                                const uint32_t lw2 = lbn32[l2];
                                const uint32_t si2 = (lw2 + (r2 >> (32u - l2))) & 0xffffu;
                                if (si2 < (lw2 >> 16)) {
                                    out.put(si2 < (uint32_t)HOT ? (uint32_t)sm.sorted8[si2] : (uint32_t)(rare[si2] & 0xffu));
                                    l += l2;
                                    // ... and a third one from the word behind the window
                                    const uint32_t r3 = __brev(in.peek32_at(l));
                                    const uint32_t l3 = lit_len((r3 >> 13) | 15u);
                                    if (l3 <= 15u && l + l3 <= in.avail) {
                                        const uint32_t lw3 = lbn32[l3];
                                        const uint32_t si3 = (lw3 + (r3 >> (32u - l3))) & 0xffffu;
                                        if (si3 < (lw3 >> 16)) {
                                            out.put(si3 < (uint32_t)HOT ? (uint32_t)sm.sorted8[si3] : (uint32_t)(rare[si3] & 0xffu));
                                            l += l3;
                                        }
                                    }
                                }
                            }
                            in.drop_long(l);
                        } else if (in.drop(l), lo8 == 0u) {  // symbol 256
                            state = ST_BLOCK;
                        } else if (lo8 > 29u) {  // symbols 286, 287: "invalid literal/length code"
                            end_kind = END_ERR;
                            state = ST_DONE;
                        } else {
                            // length / distance pair.  A pair cut by the end of the input ends the stream before it (the bytes
                            // decoded so far are the result), so nothing needs undoing.
                            const uint32_t li = lo8 - 1u;
                            const uint32_t lx = c_len_extra[li];
                            int bad = -1;
                            uint32_t mlen = 0, dist = 0;
                            if (in.avail < lx) {
                                bad = END_TRUNC;
                            } else {
                                mlen = c_len_base[li] + in.peek(lx);
                                in.drop(lx);
                                uint32_t dl = 0;
                                const int dsym = canon_decode(in, sc.aux_lim, sc.aux_base, sc.aux_sorted, dist_max, &dl);
                                if (dsym < 0) bad = dsym == -2 ? END_TRUNC : END_ERR;
                                else if (dsym > 29) bad = END_ERR;  // "invalid distance code"
                                if (bad < 0) {
                                    in.drop(dl);
                                    const uint32_t dx = c_dist_extra[dsym];
                                    if (in.avail < dx) {
                                        bad = END_TRUNC;
                                    } else {
                                        dist = c_dist_base[dsym] + in.peek(dx);
                                        in.drop(dx);
                                        if (dist > out.total) bad = END_ERR;  // "invalid distance too far back"
                                    }
                                }
                            }
                            if (bad >= 0) {
                                end_kind = bad;
                                state = ST_DONE;
                            } else {
                                aux = mlen;
                                m_dist = dist;
                                state = ST_MATCH;
                            }
                        }
                    }
                } else if (state == ST_MATCH) {
                    const uint32_t n = min(aux, 4u);
                    for (uint32_t k = 0; k < n; ++k) out.copy1(m_dist);
                    aux -= n;
                    if (aux == 0) state = ST_SYM;
                } else if (state == ST_STORED) {
                    if (in.avail < 8) {
                        end_kind = END_TRUNC;
                        state = ST_DONE;
                    } else {
                        out.put(in.peek(8));
                        in.drop(8);
                        if (--aux == 0) state = ST_BLOCK;
                    }
                }
                if ((++iter & (FLUSH_EVERY - 1)) == 0) out.flush(false);
            }
            out.flush(false);
        }
        if (mine) {
            if (end_kind != END_OK) out.flush(true);  // a truncated stream returns what it decoded
            int32_t st = S5B_OK;
            if (end_kind == END_ERR) st = S5B_ERR_PRESS;
            else if (!out.store) st = S5B_ERR_NOSPACE;
            a.status[r] = st;
            // on overflow out_len reports the size the stream needs (the caller retries with a larger slot)
            a.out_len[r] = end_kind == END_ERR ? 0u : out.total;
        }
    }
}

}  // namespace inft

size_t inflate_work_bytes(int num_sms) {
    return (size_t)num_sms * inft::TI_CTAS_PER_SM * inft::TI_WARPS * 32 * sizeof(inft::TiScratch) + 256;
}

cudaError_t launch_inflate_threads(const InflateArgs &a, uint32_t max_len, int num_sms, cudaStream_t st) {
    if (!a.work || a.work_bytes < inflate_work_bytes(num_sms)) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t want = (a.n_reads + inft::TI_WARPS * 32 - 1) / (inft::TI_WARPS * 32);
    uint64_t cap = (uint64_t)num_sms * inft::TI_CTAS_PER_SM;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (!grid) grid = 1;
    inft::inflate_thread_kernel<<<grid, inft::TI_WARPS * 32, 0, st>>>(a, max_len, static_cast<inft::TiScratch *>(a.work));
    return cudaGetLastError();
}

}  // namespace s5b
