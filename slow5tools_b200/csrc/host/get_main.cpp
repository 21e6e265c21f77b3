// get_main.cpp -- `slow5tools-b200 get FILE [READ_ID ...]`: random access by read id (src/get.c; slow5_get, slow5.c:3661-3716;
// index loading slow5_idx.c:102-186,414-520).  The ids come from the command line, from --list FILE or from standard input
// (one per line, src/get.c:283-309); the index FILE.idx is loaded, or built first when it does not exist
// (slow5_idx_init, slow5_idx.c:102-153).  Each batch of -K ids is fetched with pread() and goes through the same conversion
// loop as `view` (convert_records): decompress, parse, re-encode for the requested output, codec calls batched on the GPU --
// the slot src/get.c:359 fills with work_db(&core, &db, work_per_single_read_get).
#include <getopt.h>
#include <unistd.h>
#include <sys/stat.h>

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/slow5b200.h"
#include "blow5_io.hpp"

#include "cli_common.hpp"

using namespace s5b;

int index_main(int argc, char **argv);

#define GET_ERROR(fmt, ...) fprintf(stderr, "[%s::ERROR]\033[1;31m " fmt "\033[0m\n", __func__, __VA_ARGS__)
#define GET_WARNING(fmt, ...) fprintf(stderr, "[%s::WARNING]\033[1;33m " fmt "\033[0m\n", __func__, __VA_ARGS__)

namespace {

struct Where {
    uint64_t offset, size;
};

// slow5_idx_read (slow5_idx.c:414-520)
bool load_index(const std::string &path, const uint8_t version[3], std::unordered_map<std::string, Where> &map, std::string &err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        err = "cannot open the index file";
        return false;
    }
    std::vector<uint8_t> b;
    uint8_t tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) b.insert(b.end(), tmp, tmp + got);
    fclose(f);
    if (b.size() < 64 + 8 || memcmp(b.data(), "SLOW5IDX\1", 9) != 0 || memcmp(b.data() + b.size() - 8, "XDI5WOLS", 8) != 0) {
        err = "malformed index file (bad magic number or missing end marker)";
        return false;
    }
    if (memcmp(b.data() + 9, version, 3) != 0) GET_WARNING("%s", "index version differs from the file version");
    size_t pos = 64;
    const size_t end = b.size() - 8;
    while (pos < end) {
        if (pos + 2 > end) break;
        uint16_t n;
        memcpy(&n, b.data() + pos, 2);
        if (pos + 2 + n + 16 > end) {
            err = "malformed index file (truncated entry)";
            return false;
        }
        Where w;
        memcpy(&w.offset, b.data() + pos + 2 + n, 8);
        memcpy(&w.size, b.data() + pos + 2 + n + 8, 8);
        map.emplace(std::string(reinterpret_cast<const char *>(b.data() + pos + 2), n), w);
        pos += 2 + (size_t)n + 16;
    }
    return true;
}

void usage(FILE *f) {
    fprintf(f,
            "Usage: slow5tools-b200 get [OPTIONS] [SLOW5_FILE] [READ_ID]...\n"
            "Display the read entry for each specified read id from a slow5 file.\n"
            "With no READ_ID, read from standard input newline separated read ids.\n\n"
            "OPTIONS:\n"
            "    --to FORMAT                   specify output file format [blow5]\n"
            "    -o, --output [FILE]           output contents to FILE [default: stdout]\n"
            "    -c, --compress REC_MTD        record compression method [zlib] (only for blow5 format)\n"
            "    -s, --sig-compress SIG_MTD    signal compression method [svb-zd] (only for blow5 format)\n"
            "    -t, --threads INT             number of host threads for parsing/formatting [8]\n"
            "    -K, --batchsize INT           number of records fetched at once [4096]\n"
            "    -l, --list [FILE]             list of read ids provided as a single-column text file with one read id per line.\n"
            "    --skip                        warn and continue if a read_id was not found.\n"
            "    --index [FILE]                path to a custom slow5 index (experimental).\n"
            "    -h, --help                    display this message and exit\n"
            "FORMAT: slow5, blow5      REC_MTD: none, zlib, zstd      SIG_MTD: none, svb-zd, ex-zd\n");
}

}  // namespace

int get_main(int argc, char **argv) {
    static const struct option long_opts[] = {
        {"to", required_argument, nullptr, 'b'},        {"compress", required_argument, nullptr, 'c'},
        {"sig-compress", required_argument, nullptr, 's'}, {"batchsize", required_argument, nullptr, 'K'},
        {"output", required_argument, nullptr, 'o'},    {"list", required_argument, nullptr, 'l'},
        {"skip", no_argument, nullptr, 1000},           {"threads", required_argument, nullptr, 't'},
        {"help", no_argument, nullptr, 'h'},            {"index", required_argument, nullptr, 1001},
        {nullptr, 0, nullptr, 0}};
    const char *arg_sig = nullptr, *arg_rec = nullptr, *arg_to = nullptr, *arg_out = nullptr, *arg_list = nullptr,
               *arg_index = nullptr;
    int threads = 8;
    long batch = 4096;
    bool skip = false;
    int opt;
    optind = 1;
    while ((opt = getopt_long(argc, argv, "o:b:c:s:K:l:t:h", long_opts, nullptr)) != -1) {
        switch (opt) {
            case 'b': arg_to = optarg; break;
            case 'c': arg_rec = optarg; break;
            case 's': arg_sig = optarg; break;
            case 't': threads = atoi(optarg); break;
            case 'o': arg_out = optarg; break;
            case 'K': batch = atol(optarg); break;
            case 'l': arg_list = optarg; break;
            case 1000: skip = true; break;
            case 1001: arg_index = optarg; break;
            case 'h': usage(stdout); return 0;
            default: usage(stderr); return 1;
        }
    }
    if (skip) GET_WARNING("Will skip records that are not found%s", "");
    if (threads < 1 || batch < 1) {
        GET_ERROR("%s", "invalid -t / -K value");
        return 1;
    }
    // output format and compression: the same rules as view (src/misc.c:178-249)
    Fmt fmt_out = FMT_UNKNOWN;
    if (arg_to && (fmt_out = fmt_from_name(arg_to)) == FMT_UNKNOWN) {
        GET_ERROR("invalid output format '%s'", arg_to);
        return 1;
    }
    if (arg_out) {
        const Fmt by_ext = fmt_from_path(arg_out);
        if (fmt_out == FMT_UNKNOWN) fmt_out = by_ext;
        else if (by_ext != FMT_UNKNOWN && by_ext != fmt_out) {
            GET_ERROR("%s", "output file extension does not match the output format");
            return 1;
        }
    }
    if (fmt_out == FMT_UNKNOWN) fmt_out = FMT_BINARY;  // get defaults to blow5 (src/misc.c:209-214); view to slow5
    if (fmt_out == FMT_ASCII && (arg_rec || arg_sig)) {
        GET_ERROR("%s", "compression options (-c / -s) are only valid for blow5 output");
        return 1;
    }
    int rec_out = PRESS_ZLIB, sig_out = PRESS_SVB_ZD;
    if (arg_rec && (rec_out = press_from_name(arg_rec)) == PRESS_BAD) {
        GET_ERROR("invalid record compression method '%s'", arg_rec);
        return 1;
    }
    if (arg_sig && (sig_out = press_from_name(arg_sig)) == PRESS_BAD) {
        GET_ERROR("invalid signal compression method '%s'", arg_sig);
        return 1;
    }
    if (fmt_out == FMT_ASCII) rec_out = sig_out = PRESS_NONE;
    if ((rec_out != PRESS_NONE && rec_out != PRESS_ZLIB && rec_out != PRESS_ZSTD) ||
        (sig_out != PRESS_NONE && sig_out != PRESS_SVB_ZD && sig_out != PRESS_EX_ZD)) {
        GET_ERROR("%s", "this build supports record compression none/zlib/zstd and signal compression none/svb-zd/ex-zd only");
        return 1;
    }
    if (optind >= argc) {
        GET_ERROR("missing slow5 or blow5 file%s", "");
        usage(stderr);
        return 1;
    }
    const char *in_path = argv[optind];
    const bool from_stream = optind == argc - 1;  // no ids on the command line: stdin or --list
    FILE *list_in = stdin;
    if (arg_list && !(list_in = fopen(arg_list, "r"))) {
        GET_ERROR("File %s could not be opened - %s.", arg_list, strerror(errno));
        return 1;
    }
    Reader rd;
    if (!reader_open(rd, in_path, FMT_UNKNOWN)) {
        GET_ERROR("cannot open %s. %s", in_path, rd.err.c_str());
        return 1;
    }
    const Header &hdr = rd.hdr;
    if ((hdr.record_method != PRESS_NONE && hdr.record_method != PRESS_ZLIB && hdr.record_method != PRESS_ZSTD) ||
        (hdr.signal_method != PRESS_NONE && hdr.signal_method != PRESS_SVB_ZD && hdr.signal_method != PRESS_EX_ZD)) {
        GET_ERROR("%s", "input uses a compression method this build does not support");
        return 1;
    }
    FILE *fout = stdout;
    if (arg_out && !(fout = fopen(arg_out, "wb"))) {
        GET_ERROR("File '%s' could not be opened - %s.", arg_out, strerror(errno));
        return 1;
    }
    // ---- index: load, building it first when it is not there (slow5_idx_init, slow5_idx.c:102-153)
    std::string idx_path = arg_index ? std::string(arg_index) : std::string(in_path) + ".idx";
    if (!arg_index && access(idx_path.c_str(), R_OK) != 0) {
        fprintf(stderr, "[%s::INFO] Index file not found. Creating an index at '%s'.\n", __func__, idx_path.c_str());
        char a0[] = "index";
        std::string p(in_path);
        char *av[] = {a0, &p[0], nullptr};
        if (index_main(2, av) != 0) {
            GET_ERROR("Error loading index file for %s", in_path);
            return 1;
        }
    }
    std::unordered_map<std::string, Where> index;
    {
        std::string err;
        if (!load_index(idx_path, hdr.version, index, err)) {
            GET_ERROR("Error loading index file for %s: %s", in_path, err.c_str());
            return 1;
        }
    }
    const bool need_gpu = hdr.record_method != PRESS_NONE || hdr.signal_method != PRESS_NONE || rec_out != PRESS_NONE ||
                          sig_out != PRESS_NONE;
    s5b_ctx_t *gpu = nullptr;
    if (need_gpu) {
        const int rc = s5b_ctx_create(-1, &gpu);
        if (rc != S5B_OK) {
            GET_ERROR("cannot initialise the GPU codec: %s", s5b_strerror(rc));
            return 1;
        }
    }
    {
        const std::string h = header_to_mem(hdr, fmt_out, rec_out, sig_out);
        if (fwrite(h.data(), 1, h.size(), fout) != h.size()) {
            GET_ERROR("%s", "Could not write the output header");
            return 1;
        }
    }
    // ---- the id stream
    int next_arg = optind + 1;
    char *line = nullptr;
    size_t line_cap = 0;
    bool failed = false;
    auto next_id = [&](std::string &id) -> bool {
        if (!from_stream) {
            if (next_arg >= argc) return false;
            id = argv[next_arg++];
            return true;
        }
        ssize_t got;
        while ((got = getline(&line, &line_cap, list_in)) != -1) {  // src/get.c:283-309
            size_t n = (size_t)got;
            if (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) --n;
            if (n > 0 && line[n - 1] == '\r') --n;
            if (n == 0) continue;
            id.assign(line, n);
            return true;
        }
        return false;
    };
    const int fd = fileno(rd.fp);
    const bool text_in = rd.fmt == FMT_ASCII;
    auto next_record = [&](std::vector<uint8_t> &mem) -> int {
        std::string id;
        while (next_id(id)) {
            const auto it = index.find(id);
            if (it == index.end()) {  // slow5_get: SLOW5_ERR_NOTFOUND (slow5.c:3686-3693)
                if (skip) {
                    GET_WARNING("Read ID '%s' was not found. Skipping.", id.c_str());
                    continue;
                }
                GET_ERROR("Read ID '%s' was not found.", id.c_str());
                failed = true;
                return -1;
            }
            // binary: the record without its u64 size prefix; text: the line without its newline
            const uint64_t lead = text_in ? 0 : 8, trail = text_in ? 1 : 0;
            if (it->second.size < lead + trail) {
                GET_ERROR("Index entry of '%s' is malformed.", id.c_str());
                failed = true;
                return -1;
            }
            // the index is untrusted input: an entry may not reach past the end of the file
            struct stat fst;
            if (fstat(fd, &fst) == 0 && S_ISREG(fst.st_mode) &&
                (it->second.offset > (uint64_t)fst.st_size || it->second.size > (uint64_t)fst.st_size - it->second.offset)) {
                GET_ERROR("Index entry of '%s' points outside the file.", id.c_str());
                failed = true;
                return -1;
            }
            mem.resize(it->second.size - lead - trail);
            size_t done = 0;
            while (done < mem.size()) {
                const ssize_t r = pread(fd, mem.data() + done, mem.size() - done, (off_t)(it->second.offset + lead + done));
                if (r <= 0) {
                    GET_ERROR("Could not read the record of '%s' - %s.", id.c_str(), r < 0 ? strerror(errno) : "file is shorter than its index says");
                    failed = true;
                    return -1;
                }
                done += (size_t)r;
            }
            return 1;
        }
        return 0;
    };
    int ret = convert_records(hdr, rd.fmt, next_record, fout, gpu, fmt_out, rec_out, sig_out, batch, threads);
    if (failed) {
        GET_ERROR("Could not fetch records.%s", "");
        ret = 1;
    }
    free(line);
    fflush(fout);
    if (ret == 0 && fmt_out == FMT_BINARY && fwrite("5WOLB", 1, 5, fout) != 5) ret = 1;
    if (fout != stdout) {
        if (fclose(fout) != 0) ret = 1;
    } else {
        fflush(fout);
    }
    if (list_in != stdin) fclose(list_in);
    reader_close(rd);
    if (gpu && getenv("S5B_ORDERLY_EXIT")) s5b_ctx_destroy(gpu);
    return ret;
}
