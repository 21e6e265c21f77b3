// Development harness: verdict parity of zstd_core.h against libzstd on single-bit corruptions.
//   g++ -O1 -g -std=c++17 -DS5BZ_TRACE tests/dev/zstd_verdicts.cpp -o /tmp/zverd -ldl && /tmp/zverd
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>
#include "../../slow5tools_b200/csrc/zstd_core.h"
typedef size_t (*compress_fn)(void *, size_t, const void *, size_t, int);
typedef size_t (*decompress_fn)(void *, size_t, const void *, size_t);
typedef unsigned (*iserr_fn)(size_t);
typedef const char *(*errname_fn)(size_t);
typedef void *(*create_fn)();
typedef size_t (*setp_fn)(void *, int, int);
typedef size_t (*compress2_fn)(void *, void *, size_t, const void *, size_t);
int main() {
    void *h = dlopen("libzstd.so.1", RTLD_NOW);
    compress_fn zc = (compress_fn)dlsym(h, "ZSTD_compress");
    decompress_fn zd = (decompress_fn)dlsym(h, "ZSTD_decompress");
    iserr_fn ze = (iserr_fn)dlsym(h, "ZSTD_isError");
    errname_fn zn = (errname_fn)dlsym(h, "ZSTD_getErrorName");
    create_fn mk = (create_fn)dlsym(h, "ZSTD_createCCtx");
    setp_fn setp = (setp_fn)dlsym(h, "ZSTD_CCtx_setParameter");
    compress2_fn zc2 = (compress2_fn)dlsym(h, "ZSTD_compress2");
    const bool checksum = getenv("CHECKSUM") != nullptr;
    std::mt19937 rng(3);
    static s5bz::Tables t;
    std::vector<uint8_t> lit(128 << 10);
    std::map<int, int> strict, lenient;
    int total = 0;
    for (int round = 0; round < 40; ++round) {
        size_t n = 2000 + rng() % 9000;
        std::vector<uint8_t> raw(n);
        for (size_t i = 0; i < n; ++i) raw[i] = (i < n / 5) ? 0 : (round % 2 ? (uint8_t)(std::normal_distribution<double>(9, 6)(rng)) : "hello world, hello zstd "[i % 24]);
        std::vector<uint8_t> z(n + 1000);
        size_t zl;
        if (checksum) {
            void *c = mk();
            setp(c, 100, round % 3 == 0 ? 19 : 1);
            setp(c, 201, 1);
            zl = zc2(c, z.data(), z.size(), raw.data(), n);
        } else {
            zl = zc(z.data(), z.size(), raw.data(), n, round % 3 == 0 ? 19 : 1);
        }
        for (int k = 0; k < 400; ++k) {
            std::vector<uint8_t> b(z.begin(), z.begin() + zl);
            size_t at = rng() % zl;
            b[at] ^= 1 << (rng() % 8);
            std::vector<uint8_t> o1(n + 64), o2(n + 64);
            size_t r1 = zd(o1.data(), n, b.data(), zl);
            uint64_t on = 0;
            s5bz_fail_line = 0;
            int r2 = s5bz::decode_frame(t, b.data(), zl, o2.data(), n, lit.data(), (uint32_t)lit.size(), &on);
            ++total;
            bool ok1 = !ze(r1), ok2 = r2 == 0;
            if (ok1 && !ok2) { strict[s5bz_fail_line]++; if (strict[s5bz_fail_line] <= 2) printf("we reject (line %d), libzstd accepts: round %d at byte %zu of %zu\n", s5bz_fail_line, round, at, zl); }
            if (!ok1 && ok2) { lenient[0]++; if (lenient[0] <= 5) printf("we accept, libzstd rejects (%s): round %d at byte %zu of %zu\n", zn(r1), round, at, zl); }
            if (ok1 && ok2 && (r1 != on || memcmp(o1.data(), o2.data(), on))) printf("both accept, bytes differ! round %d at %zu\n", round, at);
        }
    }
    printf("%d corruptions; we-stricter by line:", total);
    for (auto &kv : strict) printf(" L%d:%d", kv.first, kv.second);
    printf("; we-more-lenient: %d\n", lenient[0]);
}
