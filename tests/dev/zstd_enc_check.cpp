// Development harness (not part of the product, not run by pytest): a serial model of the frame layout written by
// slow5tools_b200/csrc/zstd_encode_kernels.cu, built from the same zstd_enc_core.h pieces, decoded by system
// libzstd (dlopen'ed) and by the repo's own serial decoder.
//   g++ -O1 -g -std=c++17 tests/dev/zstd_enc_check.cpp -o /tmp/zenc -ldl && /tmp/zenc
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../slow5tools_b200/csrc/zstd_enc_core.h"

typedef size_t (*decompress_fn)(void *, size_t, const void *, size_t);
typedef unsigned (*iserr_fn)(size_t);
typedef const char *(*errname_fn)(size_t);
typedef unsigned long long (*fcs_fn)(const void *, size_t);

// host-only: length-limited Huffman lengths (same repair rule as huff_common.cuh)
static void lengths(const uint32_t *hist, int limit, uint8_t *len) {
    std::vector<std::pair<uint32_t, int>> leaves;
    for (int s = 0; s < 256; ++s) {
        len[s] = 0;
        if (hist[s]) leaves.push_back({hist[s], s});
    }
    std::sort(leaves.begin(), leaves.end());
    const int used = (int)leaves.size();
    std::vector<uint32_t> weight(2 * used);
    std::vector<int> parent(2 * used, 0);
    for (int i = 0; i < used; ++i) weight[i] = leaves[i].first;
    int li = 0, ii = used, next = used;
    for (int j = 0; j < used - 1; ++j) {
        int pick[2];
        for (int t = 0; t < 2; ++t) {
            if (li < used && (ii >= next || weight[li] <= weight[ii])) pick[t] = li++;
            else pick[t] = ii++;
        }
        weight[next] = weight[pick[0]] + weight[pick[1]];
        parent[pick[0]] = parent[pick[1]] = next;
        ++next;
    }
    const int root = next - 1;
    std::vector<int> depth(2 * used);
    depth[root] = 0;
    for (int v = root - 1; v >= 0; --v) depth[v] = depth[parent[v]] + 1;
    int bl[16] = {0};
    bool clamped = false;
    uint32_t kraft = 0;
    for (int i = 0; i < used; ++i) {
        int d = depth[i];
        if (d > limit) { d = limit; clamped = true; }
        bl[d]++;
        kraft += 1u << (limit - d);
    }
    if (clamped) {
        int excess = (int)kraft - (1 << limit);
        while (excess > 0) {
            int bits = limit - 1;
            while (bl[bits] == 0) --bits;
            bl[bits]--; bl[bits + 1] += 2; bl[limit]--; --excess;
        }
        int i = 0;
        for (int bits = limit; bits >= 1; --bits)
            for (int c = bl[bits]; c > 0; --c) depth[i++] = bits;
    }
    for (int i = 0; i < used; ++i) len[leaves[i].second] = (uint8_t)depth[i];
}

static void put_bits(std::vector<uint8_t> &b, size_t base_bit, uint32_t v, int n) {
    for (int i = 0; i < n; ++i)
        if ((v >> i) & 1u) {
            const size_t p = base_bit + i;
            if ((p >> 3) >= b.size()) b.resize((p >> 3) + 1, 0);
            b[p >> 3] |= (uint8_t)(1u << (p & 7));
        }
}

static void encode_block(const uint8_t *src, uint32_t n, bool last, std::vector<uint8_t> &out) {
    auto hdr3 = [&](uint32_t v) { out.push_back(v); out.push_back(v >> 8); out.push_back(v >> 16); };
    uint32_t hist[256] = {0};
    for (uint32_t i = 0; i < n; ++i) hist[src[i]]++;
    int used = 0;
    for (int s = 0; s < 256; ++s) used += hist[s] != 0;
    auto raw = [&]() { hdr3(s5bz::block_header(last, 0, n)); out.insert(out.end(), src, src + n); };
    if (n == 0) { raw(); return; }
    if (used == 1) { hdr3(s5bz::block_header(last, 1, n)); out.push_back(src[0]); return; }
    uint8_t len[256], weights[256];
    uint16_t code[256];
    lengths(hist, s5bz::HUF_MAX_BITS, len);
    int nsym;
    const int mb = s5bz::huf_codes_from_lengths(len, weights, code, &nsym);
    if (!mb) { raw(); return; }
    static s5bz::WeightEnc we;
    uint8_t tree[s5bz::HUF_TREE_MAX_BYTES];
    const int tb = s5bz::huf_write_tree(we, weights, nsym, tree);
    if (!tb) { raw(); return; }
    const bool four = n >= 256;
    const uint32_t q = four ? (n + 3) / 4 : n;
    uint32_t sz[4] = {0, 0, 0, 0};
    const int ns = four ? 4 : 1;
    for (int k = 0; k < ns; ++k) {
        const uint32_t s0 = k * q, s1 = std::min(n, s0 + q);
        uint32_t bits = 0;
        for (uint32_t i = s0; i < s1; ++i) bits += len[src[i]];
        sz[k] = (bits + 1 + 7) / 8;
    }
    const uint32_t comp = tb + (four ? 6 : 0) + sz[0] + sz[1] + sz[2] + sz[3];
    uint64_t lh;
    const int lhn = s5bz::literals_header(2, four, n, comp, &lh);
    const uint32_t bsize = lhn + comp + 1;
    if (bsize >= n) { raw(); return; }
    hdr3(s5bz::block_header(last, 2, bsize));
    for (int i = 0; i < lhn; ++i) out.push_back((uint8_t)(lh >> (8 * i)));
    out.insert(out.end(), tree, tree + tb);
    if (four)
        for (int k = 0; k < 3; ++k) { out.push_back(sz[k]); out.push_back(sz[k] >> 8); }
    for (int k = 0; k < ns; ++k) {
        const uint32_t s0 = k * q, s1 = std::min(n, s0 + q);
        std::vector<uint8_t> st(sz[k], 0);
        size_t bit = 0;
        for (uint32_t i = s1; i-- > s0;) {
            put_bits(st, bit, code[src[i]], len[src[i]]);
            bit += len[src[i]];
        }
        put_bits(st, bit, 1, 1);
        if (st.size() != sz[k]) { printf("size mismatch\n"); exit(1); }
        out.insert(out.end(), st.begin(), st.end());
    }
    out.push_back(0);  // Sequences_Section_Header: no sequences
}

int main() {
    void *h = dlopen("libzstd.so.1", RTLD_NOW);
    if (!h) { printf("no libzstd\n"); return 2; }
    decompress_fn zd = (decompress_fn)dlsym(h, "ZSTD_decompress");
    iserr_fn ze = (iserr_fn)dlsym(h, "ZSTD_isError");
    errname_fn zn = (errname_fn)dlsym(h, "ZSTD_getErrorName");
    fcs_fn zf = (fcs_fn)dlsym(h, "ZSTD_getFrameContentSize");
    std::mt19937 rng(11);
    int fails = 0, cases = 0;
    size_t tot_in = 0, tot_out = 0;
    static s5bz::Tables t;
    std::vector<uint8_t> lit(128 << 10);
    for (int iter = 0; iter < 4000; ++iter) {
        const int kind = iter % 9;
        size_t n = (iter < 300) ? iter : (rng() % (iter % 50 == 0 ? 300000 : 20000));
        std::vector<uint8_t> raw(n);
        for (size_t i = 0; i < n; ++i) {
            switch (kind) {
                case 0: raw[i] = rng(); break;
                case 1: raw[i] = (uint8_t)(std::normal_distribution<double>(9, 6)(rng)); break;
                case 2: raw[i] = "the quick brown fox "[i % 20] ^ ((rng() % 50 == 0) ? 1 : 0); break;
                case 3: raw[i] = 0; break;
                case 4: raw[i] = (i / 7) & 0xff; break;
                case 5: raw[i] = (rng() % 4); break;
                case 6: raw[i] = (i < n / 4) ? ((rng() % 40 == 0) ? 1 : 0) : (uint8_t)(std::normal_distribution<double>(9, 6)(rng)); break;
                case 7: raw[i] = (uint8_t)(std::exponential_distribution<double>(0.02)(rng)); break;  // skewed, all 256 values
                default: raw[i] = (uint8_t)(rng() % (1 + (i % 200))); break;
            }
        }
        std::vector<uint8_t> f(12);
        f.resize(s5bz::write_frame_header(f.data(), n));
        const uint32_t BLK = 6144;
        size_t b0 = 0;
        const size_t split = (kind == 6) ? n / 4 : 0;
        do {
            size_t b1 = n;
            if (split > b0) b1 = split;
            if (b1 - b0 > BLK) b1 = b0 + BLK;
            encode_block(raw.data() + b0, (uint32_t)(b1 - b0), b1 == n, f);
            b0 = b1;
        } while (b0 < n);
        ++cases;
        tot_in += n;
        tot_out += f.size();
        std::vector<uint8_t> out(n + 8, 0xAA);
        const size_t rc = zd(out.data(), n, f.data(), f.size());
        const bool ok1 = !ze(rc) && rc == n && memcmp(out.data(), raw.data(), n) == 0 && zf(f.data(), f.size()) == n;
        uint64_t on = 0;
        std::vector<uint8_t> out2(n + 8, 0xAA);
        const int rc2 = s5bz::decode_frame(t, f.data(), f.size(), out2.data(), n, lit.data(), (uint32_t)lit.size(), &on);
        const bool ok2 = rc2 == 0 && on == n && memcmp(out2.data(), raw.data(), n) == 0;
        if (!ok1 || !ok2) {
            if (++fails < 12)
                printf("FAIL iter %d kind %d n %zu: libzstd %s (%zu), own rc %d\n", iter, kind, n, ze(rc) ? zn(rc) : "ok", rc, rc2);
        }
    }
    printf("%d cases, %d failures, ratio %.4f\n", cases, fails, (double)tot_out / tot_in);
    return fails != 0;
}
