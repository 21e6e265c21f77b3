// index_main.cpp -- `slow5tools-b200 index FILE` and the library's index builder (part of libslow5b200.so): writes FILE.idx,
// byte-identical to the reference's index
// (src/index.c -> slow5_idx_create / slow5_idx_build / slow5_idx_write, slow5lib/src/slow5_idx.c:155-414).
//
// Index file (slow5_idx.c:360-412): "SLOW5IDX\1", the data file's version (3 bytes), zero padding up to byte 64, then per
// record `u16 read_id_len, read_id, u64 offset, u64 size` in file order -- offset = where the record starts (its u64 size
// prefix for BLOW5, its line for SLOW5), size = bytes up to the next record -- and the marker "XDI5WOLS".
//
// The reference walks the file with one fread per record and, for zlib records, inflates the first 256 bytes of each to
// reach the read_id (:283-334).  Here the file is read in large chunks, the size chain is walked in place, and the
// decompression of every record's head happens in one GPU batch per chunk (s5b_blow5_read_ids_host).
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../../include/slow5b200.h"
#include "blow5_io.hpp"

using namespace s5b;

#define IDX_ERROR(fmt, ...) fprintf(stderr, "[%s::ERROR]\033[1;31m " fmt "\033[0m\n", __func__, __VA_ARGS__)

namespace {

struct Entry {
    uint64_t offset, size;
    uint64_t id_at;  // into the id slab
    uint32_t id_len;
};

bool add_ids(const uint8_t *ids, const uint64_t *id_off, const std::vector<uint64_t> &offs, const std::vector<uint64_t> &sizes,
             std::string &slab, std::vector<Entry> &out, std::unordered_set<std::string> &seen) {
    for (size_t i = 0; i < offs.size(); ++i) {
        const uint64_t len = id_off[i + 1] - id_off[i];
        std::string id(reinterpret_cast<const char *>(ids + id_off[i]), len);
        if (!seen.insert(id).second) {  // slow5_idx_insert refuses duplicates (slow5_idx.c:434-445)
            IDX_ERROR("Read ID '%s' is duplicated", id.c_str());
            return false;
        }
        out.push_back(Entry{offs[i], sizes[i], (uint64_t)slab.size(), (uint32_t)len});
        slab += id;
    }
    return true;
}

}  // namespace

int index_main(int argc, char **argv) {
    const char *path = nullptr;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-h") || !strcmp(argv[i], "--help")) {
            printf("Usage: slow5tools-b200 index [SLOW5|BLOW5_FILE]\nCreate a slow5 or blow5 index file.\n\nOPTIONS:\n"
                   "    -h, --help\n        Display this message and exit.\n");
            return 0;
        }
        if (argv[i][0] == '-' && argv[i][1]) {
            IDX_ERROR("unknown option '%s'", argv[i]);
            return 1;
        }
        if (path) {
            IDX_ERROR("too many files given%s", "");
            return 1;
        }
        path = argv[i];
    }
    if (!path) {
        IDX_ERROR("missing slow5 or blow5 file%s", "");
        return 1;
    }
    return s5b::index_build_file(path);
}

// slow5_idx_create (slow5.c:4138-4150): builds PATH.idx; also what s5b_idx_load falls back to when the index file is missing,
// like the reference's slow5_idx_init does (slow5_idx.c:60-110).  0 = written.
int s5b::index_build_file(const char *path) {
    Reader rd;
    if (!reader_open(rd, path, FMT_UNKNOWN)) {
        IDX_ERROR("File '%s' could not be opened - %s.", path, rd.err.c_str());
        return 1;
    }
    const Header &hdr = rd.hdr;
    std::vector<Entry> entries;
    std::string slab;
    std::unordered_set<std::string> seen;
    const off_t start = ftello(rd.fp);
    int ret = 0;

    if (rd.fmt == FMT_ASCII) {  // one line per record: id = first column (slow5_idx.c:207-236)
        char *line = nullptr;
        size_t cap = 0;
        ssize_t got;
        uint64_t offset = (uint64_t)start;
        while ((got = getline(&line, &cap, rd.fp)) != -1) {
            const char *tab = static_cast<const char *>(memchr(line, '\t', (size_t)got));
            const uint64_t idl = tab ? (uint64_t)(tab - line) : (uint64_t)got;
            const uint64_t id_off[2] = {0, idl};
            if (!add_ids(reinterpret_cast<const uint8_t *>(line), id_off, {offset}, {(uint64_t)got}, slab, entries, seen)) {
                ret = 1;
                break;
            }
            offset += (uint64_t)got;
        }
        free(line);
    } else {
        if (hdr.record_method != PRESS_NONE && hdr.record_method != PRESS_ZLIB && hdr.record_method != PRESS_ZSTD) {
            IDX_ERROR("%s", "unsupported record compression method");
            return 1;
        }
        s5b_ctx_t *gpu = nullptr;
        if (hdr.record_method != PRESS_NONE) {
            const int rc = s5b_ctx_create(-1, &gpu);
            if (rc != S5B_OK) {
                IDX_ERROR("cannot initialise the GPU codec: %s", s5b_strerror(rc));
                return 1;
            }
        }
        const int fd = fileno(rd.fp);
        // ---- mapped walk: the size chain and the record heads are read straight out of the page cache, so an
        // uncompressed or zlib file costs one page touch per record instead of a copy of the whole file (only the
        // first 256 bytes of a zlib record are ever looked at); zstd frames are decoded in full and read sequentially
        struct stat sb;
        const uint8_t *map = nullptr;
        if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
            void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                map = static_cast<const uint8_t *>(m);
                madvise(m, (size_t)sb.st_size, hdr.record_method == PRESS_ZSTD ? MADV_SEQUENTIAL : MADV_RANDOM);
            }
        }
        if (map) {
            const uint64_t fsize = (uint64_t)sb.st_size;
            uint64_t pos = (uint64_t)start;
            std::vector<uint8_t> ids;
            std::vector<uint64_t> id_off, rec_off, offs, sizes;
            std::vector<uint32_t> rec_len;
            bool eof = false;
            const size_t BATCH = 200000;
            const uint64_t BATCH_BYTES = hdr.record_method == PRESS_ZSTD ? (256ull << 20) : ~0ull;  // full records go up
            auto flush = [&]() -> bool {
                if (rec_off.empty()) return true;
                uint64_t guess = 0;
                for (uint32_t l : rec_len) guess += l < 128u ? l : 128u;
                ids.resize(guess + 4096);
                id_off.resize(rec_off.size() + 1);
                int rc = S5B_ERR_NOSPACE;
                while (rc == S5B_ERR_NOSPACE) {
                    rc = s5b_blow5_read_ids_host(gpu, hdr.record_method, map, fsize, rec_off.data(), rec_len.data(),
                                                 rec_off.size(), ids.data(), ids.size(), id_off.data());
                    if (rc == S5B_ERR_NOSPACE) {
                        if (ids.size() > rec_off.size() * 65536ull + 4096) break;
                        ids.resize(ids.size() * 2 + 65536);
                    }
                }
                if (rc != S5B_OK) {
                    IDX_ERROR("could not read the record ids: %s", s5b_strerror(rc));
                    return false;
                }
                const bool ok = add_ids(ids.data(), id_off.data(), offs, sizes, slab, entries, seen);
                rec_off.clear();
                rec_len.clear();
                offs.clear();
                sizes.clear();
                return ok;
            };
            uint64_t batch_bytes = 0;
            while (!eof && ret == 0) {
                const uint64_t left = fsize - pos;
                if (left == 5 && memcmp(map + pos, "5WOLB", 5) == 0) {  // end-of-file marker (slow5_idx.c:252-262)
                    eof = true;
                    break;
                }
                if (left < 8) {
                    IDX_ERROR("Malformed blow5 record. Failed to read the record size.%s",
                              left == 0 ? " Missing blow5 end of file marker." : "");
                    ret = 1;
                    break;
                }
                uint64_t size;
                memcpy(&size, map + pos, 8);
                if (size > (1ull << 32) - 64 || left < 8 + size) {
                    IDX_ERROR("%s", size > (1ull << 32) - 64 ? "implausible record size (corrupt file?)" : "blow5 record is truncated");
                    ret = 1;
                    break;
                }
                rec_off.push_back(pos + 8);
                rec_len.push_back((uint32_t)size);
                offs.push_back(pos);
                sizes.push_back(8 + size);
                pos += 8 + size;
                batch_bytes += size;
                if (rec_off.size() >= BATCH || batch_bytes >= BATCH_BYTES) {
                    if (!flush()) ret = 1;
                    batch_bytes = 0;
                }
            }
            if (ret == 0 && !flush()) ret = 1;
            munmap(const_cast<uint8_t *>(map), (size_t)fsize);
            if (gpu) s5b_ctx_destroy(gpu);
            reader_close(rd);
            if (ret) return ret;
            goto write_index;
        }
        if (lseek(fd, start, SEEK_SET) < 0) {
            IDX_ERROR("%s", "cannot seek in the input file");
            return 1;
        }
        size_t cap = 64u << 20;
        uint8_t *buf = static_cast<uint8_t *>(gpu ? s5b_host_alloc(cap) : malloc(cap));
        std::vector<uint8_t> ids;
        std::vector<uint64_t> id_off, rec_off, offs, sizes;
        std::vector<uint32_t> rec_len;
        uint64_t filled = 0, file_pos = (uint64_t)start;  // file offset of buf[0]
        bool eof = false, file_end = false;
        while (buf && !eof && ret == 0) {
            while (!file_end && filled < cap) {
                const ssize_t got = read(fd, buf + filled, cap - filled);
                if (got < 0) {
                    IDX_ERROR("read failed: %s", strerror(errno));
                    ret = 1;
                    break;
                }
                if (got == 0) file_end = true;
                else filled += (uint64_t)got;
            }
            if (ret) break;
            rec_off.clear();
            rec_len.clear();
            offs.clear();
            sizes.clear();
            uint64_t pos = 0;
            for (;;) {
                const uint64_t left = filled - pos;
                if (file_end && left == 5 && memcmp(buf + pos, "5WOLB", 5) == 0) {  // end-of-file marker (slow5_idx.c:252-262)
                    eof = true;
                    break;
                }
                if (left < 8) {
                    if (file_end) {
                        IDX_ERROR("Malformed blow5 record. Failed to read the record size.%s",
                                  left == 0 ? " Missing blow5 end of file marker." : "");
                        ret = 1;
                    }
                    break;
                }
                uint64_t size;
                memcpy(&size, buf + pos, 8);
                if (size > (1ull << 32) - 64) {
                    IDX_ERROR("%s", "implausible record size (corrupt file?)");
                    ret = 1;
                    break;
                }
                if (left < 8 + size) {
                    if (file_end) {
                        IDX_ERROR("%s", "blow5 record is truncated");
                        ret = 1;
                    }
                    break;
                }
                rec_off.push_back(pos + 8);
                rec_len.push_back((uint32_t)size);
                offs.push_back(file_pos + pos);
                sizes.push_back(8 + size);
                pos += 8 + size;
            }
            if (ret) break;
            if (pos == 0 && !eof) {  // one record larger than the buffer: grow and read on
                const size_t ncap = cap * 2;
                uint8_t *nb = static_cast<uint8_t *>(gpu ? s5b_host_alloc(ncap) : malloc(ncap));
                if (nb) memcpy(nb, buf, filled);
                if (gpu) s5b_host_free(buf);
                else free(buf);
                buf = nb;
                cap = ncap;
                continue;
            }
            if (!rec_off.empty()) {
                uint64_t in_sum = 0;
                for (uint32_t l : rec_len) in_sum += l < 65538u ? l : 65538u;
                ids.resize(in_sum + 64);  // an id is at most 65535 bytes and never longer than its record
                id_off.resize(rec_off.size() + 1);
                int rc = S5B_ERR_NOSPACE;
                while (rc == S5B_ERR_NOSPACE) {  // (a compressed record can hold an id longer than itself: grow and retry)
                    rc = s5b_blow5_read_ids_host(gpu, hdr.record_method, buf, filled, rec_off.data(), rec_len.data(),
                                                 rec_off.size(), ids.data(), ids.size(), id_off.data());
                    if (rc == S5B_ERR_NOSPACE) {
                        if (ids.size() > rec_off.size() * 65536ull + 64) break;
                        ids.resize(ids.size() * 2 + 65536);
                    }
                }
                if (rc != S5B_OK) {
                    IDX_ERROR("could not read the record ids: %s", s5b_strerror(rc));
                    ret = 1;
                    break;
                }
                if (!add_ids(ids.data(), id_off.data(), offs, sizes, slab, entries, seen)) {
                    ret = 1;
                    break;
                }
            }
            memmove(buf, buf + pos, filled - pos);  // carry the cut record (or the marker) to the front
            filled -= pos;
            file_pos += pos;
        }
        if (!buf) {
            IDX_ERROR("%s", "out of memory");
            ret = 1;
        }
        if (gpu) {
            s5b_host_free(buf);
            s5b_ctx_destroy(gpu);
        } else {
            free(buf);
        }
    }
    reader_close(rd);
    if (ret) return ret;

write_index:
    // ---- slow5_idx_write (slow5_idx.c:360-412)
    const std::string out_path = std::string(path) + ".idx";
    FILE *fo = fopen(out_path.c_str(), "wb");
    if (!fo) {
        IDX_ERROR("File '%s' could not be opened - %s.", out_path.c_str(), strerror(errno));
        return 1;
    }
    std::string out;
    out.reserve(64 + slab.size() + entries.size() * 18 + 8);
    out.append("SLOW5IDX\1", 9);
    out.append(reinterpret_cast<const char *>(hdr.version), 3);
    out.append(64 - out.size(), '\0');
    for (const Entry &e : entries) {
        const uint16_t l = (uint16_t)e.id_len;
        out.append(reinterpret_cast<const char *>(&l), 2);
        out.append(slab, e.id_at, e.id_len);
        out.append(reinterpret_cast<const char *>(&e.offset), 8);
        out.append(reinterpret_cast<const char *>(&e.size), 8);
    }
    out.append("XDI5WOLS", 8);
    const bool ok = fwrite(out.data(), 1, out.size(), fo) == out.size();
    if (fclose(fo) != 0 || !ok) {
        IDX_ERROR("could not write '%s'", out_path.c_str());
        return 1;
    }
    return 0;
}
