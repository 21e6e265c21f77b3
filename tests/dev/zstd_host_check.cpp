// Development harness (not part of the product, not run by pytest): exercises the serial zstd decoder of
// slow5tools_b200/csrc/zstd_core.h on the CPU against system libzstd (dlopen'ed, no header needed).
//   g++ -O1 -g -std=c++17 tests/dev/zstd_host_check.cpp -o /tmp/zcheck -ldl && /tmp/zcheck
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../slow5tools_b200/csrc/zstd_core.h"

typedef size_t (*compress_fn)(void *, size_t, const void *, size_t, int);
typedef size_t (*bound_fn)(size_t);
typedef void *(*create_fn)();
typedef size_t (*setp_fn)(void *, int, int);
typedef size_t (*compress2_fn)(void *, void *, size_t, const void *, size_t);

int main() {
    void *h = dlopen("libzstd.so.1", RTLD_NOW);
    if (!h) { printf("no libzstd\n"); return 2; }
    compress_fn zc = (compress_fn)dlsym(h, "ZSTD_compress");
    bound_fn zb = (bound_fn)dlsym(h, "ZSTD_compressBound");
    create_fn mk = (create_fn)dlsym(h, "ZSTD_createCCtx");
    setp_fn setp = (setp_fn)dlsym(h, "ZSTD_CCtx_setParameter");
    compress2_fn zc2 = (compress2_fn)dlsym(h, "ZSTD_compress2");
    std::mt19937 rng(7);
    static s5bz::Tables t;
    std::vector<uint8_t> lit(128 << 10);
    int fails = 0, cases = 0;
    for (int iter = 0; iter < 3000; ++iter) {
        const int kind = iter % 8;
        size_t n = (iter < 40) ? iter : (rng() % (iter % 50 == 0 ? 400000 : 20000));
        std::vector<uint8_t> raw(n);
        for (size_t i = 0; i < n; ++i) {
            switch (kind) {
                case 0: raw[i] = rng(); break;
                case 1: raw[i] = (uint8_t)(std::normal_distribution<double>(9, 6)(rng)); break;   // svb-like
                case 2: raw[i] = "the quick brown fox "[i % 20] ^ ((rng() % 50 == 0) ? 1 : 0); break;
                case 3: raw[i] = 0; break;
                case 4: raw[i] = (i / 7) & 0xff; break;
                case 5: raw[i] = (rng() % 4); break;
                case 6: raw[i] = (i < n / 4) ? 0 : (uint8_t)(std::normal_distribution<double>(9, 6)(rng)); break;
                default: raw[i] = (uint8_t)(rng() % (1 + (i % 200))); break;
            }
        }
        const int level = (iter % 3 == 0) ? 1 : 1 + (int)(rng() % 19);
        std::vector<uint8_t> z(zb(n) + 64);
        size_t zn;
        if (iter % 11 == 0) {
            void *c = mk();
            setp(c, 100 /*ZSTD_c_compressionLevel*/, level);
            setp(c, 201 /*ZSTD_c_checksumFlag*/, 1);
            zn = zc2(c, z.data(), z.size(), raw.data(), n);
        } else {
            zn = zc(z.data(), z.size(), raw.data(), n, level);
        }
        std::vector<uint8_t> out(n + 8, 0xAA);
        uint64_t on = 0;
        const int rc = s5bz::decode_frame(t, z.data(), zn, out.data(), n, lit.data(), (uint32_t)lit.size(), &on);
        ++cases;
        if (rc != 0 || on != n || memcmp(out.data(), raw.data(), n) != 0) {
            if (fails < 10) printf("FAIL iter %d kind %d n %zu level %d zn %zu rc %d on %llu\n", iter, kind, n, level, zn, rc, (unsigned long long)on);
            ++fails;
        }
        // corruption must never crash (verdict parity is checked on the GPU tests against libzstd)
        if (zn > 8 && iter % 5 == 0) {
            std::vector<uint8_t> zz(z.begin(), z.begin() + zn);
            zz[4 + rng() % (zn - 4)] ^= 1 << (rng() % 8);
            s5bz::decode_frame(t, zz.data(), zn, out.data(), n, lit.data(), (uint32_t)lit.size(), &on);
            s5bz::decode_frame(t, z.data(), zn - 1 - rng() % (zn > 20 ? 20 : 1), out.data(), n, lit.data(), (uint32_t)lit.size(), &on);
        }
    }
    printf("%d cases, %d failures\n", cases, fails);
    return fails != 0;
}
