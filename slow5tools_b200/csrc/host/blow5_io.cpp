// blow5_io.cpp -- see blow5_io.hpp.  Host-side BLOW5/SLOW5 framing; no codec arithmetic in this file.
#include "blow5_io.hpp"
#include <sys/stat.h>
#include <new>

#include <algorithm>
#include <cerrno>
#include <cinttypes>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace s5b {

// ---- method maps (slow5_press.c:58-161) ----------------------------------------------------------
int record_press_from_byte(uint8_t b) {
    switch (b) {
        case 0: return PRESS_NONE;
        case 1: return PRESS_ZLIB;
        case 2: return PRESS_ZSTD;
        case 250: return PRESS_SVB_ZD;
        default: return PRESS_BAD;
    }
}
int signal_press_from_byte(uint8_t b) {
    switch (b) {
        case 0: return PRESS_NONE;
        case 1: return PRESS_SVB_ZD;
        case 2: return PRESS_EX_ZD;
        case 250: return PRESS_ZLIB;
        case 251: return PRESS_ZSTD;
        default: return PRESS_BAD;
    }
}
uint8_t record_press_to_byte(int m) {
    switch (m) {
        case PRESS_NONE: return 0;
        case PRESS_ZLIB: return 1;
        case PRESS_ZSTD: return 2;
        case PRESS_SVB_ZD: return 250;
        default: return 255;
    }
}
uint8_t signal_press_to_byte(int m) {
    switch (m) {
        case PRESS_NONE: return 0;
        case PRESS_SVB_ZD: return 1;
        case PRESS_EX_ZD: return 2;
        case PRESS_ZLIB: return 250;
        case PRESS_ZSTD: return 251;
        default: return 255;
    }
}
int press_from_name(const char *name) {
    if (!strcmp(name, "none")) return PRESS_NONE;
    if (!strcmp(name, "zlib")) return PRESS_ZLIB;
    if (!strcmp(name, "svb-zd")) return PRESS_SVB_ZD;
    if (!strcmp(name, "zstd")) return PRESS_ZSTD;
    if (!strcmp(name, "ex-zd")) return PRESS_EX_ZD;
    return PRESS_BAD;
}

Fmt fmt_from_name(const char *name) {
    if (!strcmp(name, "slow5")) return FMT_ASCII;
    if (!strcmp(name, "blow5")) return FMT_BINARY;
    return FMT_UNKNOWN;
}
Fmt fmt_from_path(const char *path) {
    const char *dot = strrchr(path, '.');
    if (!dot) return FMT_UNKNOWN;
    return fmt_from_name(dot + 1);
}

// ---- aux types ---------------------------------------------------------------------------------
static const struct {
    const char *name;
    int type;
    uint8_t size;
} k_prim[] = {{"int8_t", AUX_INT8, 1},   {"int16_t", AUX_INT16, 2},   {"int32_t", AUX_INT32, 4},   {"int64_t", AUX_INT64, 8},
              {"uint8_t", AUX_UINT8, 1}, {"uint16_t", AUX_UINT16, 2}, {"uint32_t", AUX_UINT32, 4}, {"uint64_t", AUX_UINT64, 8},
              {"float", AUX_FLOAT, 4},   {"double", AUX_DOUBLE, 8},   {"char", AUX_CHAR, 1}};

static bool aux_type_from_str(const std::string &s, AuxField &f) {
    f.type_str = s;
    if (s.compare(0, 4, "enum") == 0) {
        // enum{a,b,c} or enum*{a,b,c}
        const bool arr = s.size() > 4 && s[4] == '*';
        f.type = arr ? AUX_ENUM_ARRAY : AUX_ENUM;
        f.size = 1;
        return true;
    }
    std::string base = s;
    bool arr = false;
    if (!base.empty() && base.back() == '*') {
        arr = true;
        base.pop_back();
    }
    for (const auto &p : k_prim) {
        if (base == p.name) {
            f.size = p.size;
            f.type = arr ? (p.type == AUX_CHAR ? AUX_STRING : p.type + AUX_INT8_ARRAY) : p.type;
            return true;
        }
    }
    return false;
}

// ---- small helpers -----------------------------------------------------------------------------
static void split(const std::string &s, char sep, std::vector<std::string> &out) {
    out.clear();
    size_t a = 0;
    for (;;) {
        size_t b = s.find(sep, a);
        if (b == std::string::npos) {
            out.push_back(s.substr(a));
            return;
        }
        out.push_back(s.substr(a, b - a));
        a = b + 1;
    }
}

// "%f" without the zeros a fixed six-digit fraction leaves behind, and without a bare trailing point; a value that
// rounds to minus zero prints as "0" (what slow5_double_to_str produces, slow5_misc.c:379-406)
std::string double_to_str(double x) {
    char buf[512];
    const int n = snprintf(buf, sizeof buf, "%f", x);
    if (n < 0) return "";
    std::string s(buf, (size_t)(n < (int)sizeof buf ? n : (int)sizeof buf - 1));
    const size_t point = s.find('.');
    if (point != std::string::npos) {
        const size_t last = s.find_last_not_of('0');   // at the latest the point itself
        s.erase(last == point ? point : last + 1);
    }
    return s == "-0" ? std::string("0") : s;
}

static const char *MAIN_TYPES = "#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*";
static const char *MAIN_COLS = "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal";

// Parses the text block shared by both formats: "@key\tv..." lines, then the "#types" and "#columns" lines.
static bool parse_header_text(const std::string &text, Header &h, std::string &err) {
    std::vector<std::string> lines;
    split(text, '\n', lines);
    if (!lines.empty() && lines.back().empty()) lines.pop_back();
    if (lines.size() < 2) {
        err = "malformed header: missing type/column lines";
        return false;
    }
    std::vector<std::string> tok;
    for (size_t i = 0; i + 2 < lines.size(); ++i) {
        const std::string &l = lines[i];
        if (l.empty() || l[0] != '@') {
            err = "malformed header: expected '@' attribute line";
            return false;
        }
        split(l.substr(1), '\t', tok);
        if (tok.size() != (size_t)h.num_read_groups + 1) {
            err = "malformed header: attribute '" + tok[0] + "' does not have one value per read group";
            return false;
        }
        std::vector<std::string> vals(tok.begin() + 1, tok.end());
        for (auto &v : vals)
            if (v == ".") v.clear();  // slow5.c:1749-1751
        h.attrs.emplace_back(tok[0], std::move(vals));
    }
    std::vector<std::string> types, names;
    split(lines[lines.size() - 2], '\t', types);
    split(lines[lines.size() - 1], '\t', names);
    if (types.size() != names.size() || types.size() < 8 || lines[lines.size() - 2].compare(0, strlen(MAIN_TYPES), MAIN_TYPES) != 0 ||
        lines[lines.size() - 1].compare(0, strlen(MAIN_COLS), MAIN_COLS) != 0) {
        err = "malformed header: unexpected type / column line";
        return false;
    }
    for (size_t i = 8; i < types.size(); ++i) {
        AuxField f;
        f.name = names[i];
        if (!aux_type_from_str(types[i], f)) {
            err = "malformed header: unknown auxiliary type '" + types[i] + "'";
            return false;
        }
        h.aux.push_back(f);
    }
    return true;
}

bool reader_open(Reader &r, const char *path, Fmt fmt) {
    r.path = path;
    r.fmt = fmt == FMT_UNKNOWN ? fmt_from_path(path) : fmt;
    if (r.fmt == FMT_UNKNOWN) {
        r.err = "cannot tell the format of '" + r.path + "' from its extension";
        return false;
    }
    r.fp = fopen(path, "rb");
    if (!r.fp) {
        r.err = "cannot open '" + r.path + "': " + strerror(errno);
        return false;
    }
    setvbuf(r.fp, nullptr, _IOFBF, 1 << 20);
    Header &h = r.hdr;
    if (r.fmt == FMT_BINARY) {
        uint8_t fixed[68];
        if (fread(fixed, 1, sizeof fixed, r.fp) != sizeof fixed) {
            r.err = "malformed blow5 header: file too short";
            return false;
        }
        if (memcmp(fixed, "BLOW5\1", 6) != 0) {
            r.err = "malformed blow5 header: invalid magic number";
            return false;
        }
        memcpy(h.version, fixed + 6, 3);
        h.record_method = record_press_from_byte(fixed[9]);
        memcpy(&h.num_read_groups, fixed + 10, 4);
        const bool has_sig = h.version[0] > 0 || h.version[1] >= 2;  // signal method byte exists from 0.2.0
        h.signal_method = has_sig ? signal_press_from_byte(fixed[14]) : PRESS_NONE;
        if (h.version[0] > 1 || (h.version[0] == 1 && h.version[1] > 0)) {
            r.err = "file version is newer than this implementation supports (max 1.0.0)";
            return false;
        }
        if (h.record_method == PRESS_BAD || h.signal_method == PRESS_BAD) {
            r.err = "unknown compression method in blow5 header";
            return false;
        }
        uint32_t hsize;
        memcpy(&hsize, fixed + 64, 4);
        {
            struct stat fst;
            if (fstat(fileno(r.fp), &fst) == 0 && S_ISREG(fst.st_mode) && (uint64_t)hsize + 68 > (uint64_t)fst.st_size) {
                r.err = "malformed blow5 header: truncated";
                return false;
            }
        }
        std::string text;
        try {
            text.assign(hsize, '\0');
        } catch (const std::bad_alloc &) {
            r.err = "malformed blow5 header: implausible size";
            return false;
        }
        if (hsize && fread(&text[0], 1, hsize, r.fp) != hsize) {
            r.err = "malformed blow5 header: truncated";
            return false;
        }
        return parse_header_text(text, h, r.err);
    }
    // ASCII: "#slow5_version\tM.m.p\n#num_read_groups\tN\n" then the shared block up to the '#read_id' line
    std::string text;
    char *line = nullptr;
    size_t cap = 0;
    ssize_t n;
    int stage = 0;
    bool ok = false;
    while ((n = getline(&line, &cap, r.fp)) > 0) {
        std::string l(line, n);
        if (stage == 0) {
            unsigned a, b, c;
            if (sscanf(l.c_str(), "#slow5_version\t%u.%u.%u", &a, &b, &c) != 3) {
                r.err = "malformed slow5 header: expected '#slow5_version'";
                break;
            }
            h.version[0] = (uint8_t)a;
            h.version[1] = (uint8_t)b;
            h.version[2] = (uint8_t)c;
            stage = 1;
        } else if (stage == 1) {
            unsigned g;
            if (sscanf(l.c_str(), "#num_read_groups\t%u", &g) != 1) {
                r.err = "malformed slow5 header: expected '#num_read_groups'";
                break;
            }
            h.num_read_groups = g;
            stage = 2;
        } else {
            text += l;
            if (l.compare(0, 8, "#read_id") == 0) {
                ok = true;
                break;
            }
        }
    }
    free(line);
    if (!ok) {
        if (r.err.empty()) r.err = "malformed slow5 header: no column line";
        return false;
    }
    h.record_method = h.signal_method = PRESS_NONE;
    return parse_header_text(text, h, r.err);
}

void reader_close(Reader &r) {
    if (r.fp) fclose(r.fp);
    r.fp = nullptr;
}

int reader_next_mem(Reader &r, std::vector<uint8_t> &mem) {
    if (r.fmt == FMT_BINARY) {
        uint8_t pre[8];
        const size_t got = fread(pre, 1, 8, r.fp);
        if (got != 8) {
            // the end-of-file marker "5WOLB" is 5 bytes (slow5.c:3237-3259, :4644-4682)
            if (got == 5 && memcmp(pre, "5WOLB", 5) == 0 && feof(r.fp)) return 0;
            r.err = "blow5 file is truncated or has no end-of-file marker";
            return -1;
        }
        uint64_t size;
        memcpy(&size, pre, 8);
        // a size prefix cannot exceed what is left of a regular file: a corrupt prefix must not turn into a huge
        // allocation (the reference reports the failed read and exits, slow5.c:3262-3271)
        uint64_t limit = 1ull << 40;
        struct stat fst;
        const off_t here = ftello(r.fp);
        if (here >= 0 && fstat(fileno(r.fp), &fst) == 0 && S_ISREG(fst.st_mode) && (uint64_t)fst.st_size >= (uint64_t)here)
            limit = (uint64_t)fst.st_size - (uint64_t)here;
        if (size > limit) {
            r.err = "implausible record size (corrupt file?)";
            return -1;
        }
        try {
            mem.resize(size);
        } catch (const std::bad_alloc &) {
            r.err = "cannot allocate memory for a record (corrupt size prefix?)";
            return -1;
        }
        if (size && fread(mem.data(), 1, size, r.fp) != size) {
            r.err = "blow5 record is truncated";
            return -1;
        }
        return 1;
    }
    char *line = nullptr;
    size_t cap = 0;
    ssize_t n = getline(&line, &cap, r.fp);
    if (n <= 0) {
        free(line);
        return 0;
    }
    if (line[n - 1] == '\n') --n;
    mem.assign(reinterpret_cast<uint8_t *>(line), reinterpret_cast<uint8_t *>(line) + n);
    free(line);
    return 1;
}

std::string header_to_mem(const Header &h, Fmt fmt, int record_method, int signal_method) {
    std::string out;
    if (fmt == FMT_ASCII) {
        char buf[128];
        snprintf(buf, sizeof buf, "#slow5_version\t%u.%u.%u\n#num_read_groups\t%" PRIu32 "\n", h.version[0], h.version[1],
                 h.version[2], h.num_read_groups);
        out = buf;
    } else {
        uint8_t ver[3] = {h.version[0], h.version[1], h.version[2]};
        const bool below_020 = ver[0] == 0 && ver[1] < 2;
        if (below_020 && (signal_method != PRESS_NONE || (record_method != PRESS_NONE && record_method != PRESS_ZLIB))) {
            ver[0] = 0;
            ver[1] = 2;
            ver[2] = 0;
        }
        out.assign(68, '\0');
        memcpy(&out[0], "BLOW5\1", 6);
        memcpy(&out[6], ver, 3);
        out[9] = (char)record_press_to_byte(record_method);
        memcpy(&out[10], &h.num_read_groups, 4);
        out[14] = (char)signal_press_to_byte(signal_method);
    }
    const size_t text_at = out.size();
    // attributes sorted by key (slow5.c:917-925), one value per read group, missing/empty as "."
    std::vector<size_t> order(h.attrs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return strcmp(h.attrs[a].first.c_str(), h.attrs[b].first.c_str()) < 0; });
    for (size_t k : order) {
        out += '@';
        out += h.attrs[k].first;
        for (uint32_t g = 0; g < h.num_read_groups; ++g) {
            out += '\t';
            const std::string &v = g < h.attrs[k].second.size() ? h.attrs[k].second[g] : std::string();
            out += v.empty() ? std::string(".") : v;
        }
        out += '\n';
    }
    out += MAIN_TYPES;
    for (const auto &f : h.aux) {
        out += '\t';
        out += f.type_str;
    }
    out += '\n';
    out += MAIN_COLS;
    for (const auto &f : h.aux) {
        out += '\t';
        out += f.name;
    }
    out += '\n';
    if (fmt == FMT_BINARY) {
        const uint32_t hsize = (uint32_t)(out.size() - text_at);
        memcpy(&out[64], &hsize, 4);
    }
    return out;
}

// ---- binary record ------------------------------------------------------------------------------
bool record_parse_binary(const uint8_t *mem, uint64_t n, const Header &h, int signal_method, Record &rec, std::string &err) {
    uint64_t off = 0;
    auto need = [&](uint64_t k) { return off + k <= n; };
    uint16_t rid_len;
    if (!need(2)) goto bad;
    memcpy(&rid_len, mem + off, 2);
    off += 2;
    if (!need(rid_len)) goto bad;
    rec.read_id.assign(reinterpret_cast<const char *>(mem + off), rid_len);
    off += rid_len;
    if (!need(4 + 8 * 4 + 8)) goto bad;
    memcpy(&rec.read_group, mem + off, 4);
    off += 4;
    memcpy(&rec.digitisation, mem + off, 8);
    memcpy(&rec.offset, mem + off + 8, 8);
    memcpy(&rec.range, mem + off + 16, 8);
    memcpy(&rec.sampling_rate, mem + off + 24, 8);
    off += 32;
    uint64_t lrs;
    memcpy(&lrs, mem + off, 8);
    off += 8;
    // the field counts samples when the signal is stored raw and bytes when it is compressed (slow5.c:3983-3987)
    rec.sig_nbytes = signal_method == PRESS_NONE ? lrs * 2 : lrs;
    if (rec.sig_nbytes / 2 > (1ull << 40) || !need(rec.sig_nbytes)) goto bad;
    rec.len_raw_signal = signal_method == PRESS_NONE ? lrs : 0;  // unknown until the signal is decompressed
    rec.sig_bytes = mem + off;
    off += rec.sig_nbytes;
    rec.aux_bytes = mem + off;
    rec.aux_nbytes = n - off;
    // walk the aux section to validate it (slow5.c:3088-3166)
    {
        uint64_t a = 0;
        for (const auto &f : h.aux) {
            uint64_t len = 1;
            if (f.is_array()) {
                if (a + 8 > rec.aux_nbytes) goto bad;
                memcpy(&len, rec.aux_bytes + a, 8);
                a += 8;
            }
            if (len > rec.aux_nbytes || a + len * f.size > rec.aux_nbytes) goto bad;
            a += len * f.size;
        }
        if (a != rec.aux_nbytes) goto bad;
    }
    return true;
bad:
    err = "record does not parse (inconsistent field sizes)";
    return false;
}

// ---- ASCII record ---------------------------------------------------------------------------------
template <typename T>
static void put(std::string &s, T v) {
    char b[32];
    char *p = b + sizeof b;
    bool neg = false;
    unsigned long long u;
    if (v < 0) {
        neg = true;
        u = 0ull - (unsigned long long)(long long)v;
    } else {
        u = (unsigned long long)v;
    }
    do {
        *--p = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (neg) *--p = '-';
    s.append(p, b + sizeof b - p);
}

static void aux_prim_to_str(const uint8_t *d, int type, std::string &out) {
    switch (type) {
        case AUX_INT8: { int8_t v; memcpy(&v, d, 1); if (v == INT8_MAX) out += '.'; else put(out, (int)v); break; }
        case AUX_INT16: { int16_t v; memcpy(&v, d, 2); if (v == INT16_MAX) out += '.'; else put(out, (int)v); break; }
        case AUX_INT32: { int32_t v; memcpy(&v, d, 4); if (v == INT32_MAX) out += '.'; else put(out, (long long)v); break; }
        case AUX_INT64: { int64_t v; memcpy(&v, d, 8); if (v == INT64_MAX) out += '.'; else put(out, (long long)v); break; }
        case AUX_UINT8: { uint8_t v; memcpy(&v, d, 1); if (v == UINT8_MAX) out += '.'; else put(out, (unsigned)v); break; }
        case AUX_UINT16: { uint16_t v; memcpy(&v, d, 2); if (v == UINT16_MAX) out += '.'; else put(out, (unsigned)v); break; }
        case AUX_UINT32: { uint32_t v; memcpy(&v, d, 4); if (v == UINT32_MAX) out += '.'; else put(out, (unsigned long long)v); break; }
        case AUX_UINT64: { uint64_t v; memcpy(&v, d, 8); if (v == UINT64_MAX) out += '.'; else put(out, (unsigned long long)v); break; }
        case AUX_FLOAT: { float v; memcpy(&v, d, 4); if (std::isnan(v)) out += '.'; else out += double_to_str(v); break; }
        case AUX_DOUBLE: { double v; memcpy(&v, d, 8); if (std::isnan(v)) out += '.'; else out += double_to_str(v); break; }
        case AUX_CHAR: { if (*d == 0) out += '.'; else out += (char)*d; break; }
        case AUX_ENUM: { uint8_t v = *d; if (v == UINT8_MAX) out += '.'; else put(out, (unsigned)v); break; }
        default: out += '.';
    }
}

void record_to_ascii(const Record &rec, const Header &h, std::string &out, const char *sig_text, size_t sig_text_len) {
    out += rec.read_id;
    out += '\t';
    put(out, (unsigned long long)rec.read_group);
    out += '\t';
    out += double_to_str(rec.digitisation);
    out += '\t';
    out += double_to_str(rec.offset);
    out += '\t';
    out += double_to_str(rec.range);
    out += '\t';
    out += double_to_str(rec.sampling_rate);
    out += '\t';
    put(out, (unsigned long long)rec.len_raw_signal);
    out += '\t';
    // signal: comma separated, no trailing comma (slow5.c:3866-3878); already formatted (on the GPU) when sig_text is given
    if (sig_text) {
        out.append(sig_text, sig_text_len);
    } else {
        const size_t n = rec.raw_signal.size();
        const size_t base = out.size();
        out.resize(base + n * 7 + 1);
        char *p = &out[base];
        for (size_t i = 0; i < n; ++i) {
            int v = rec.raw_signal[i];
            if (i) *p++ = ',';
            if (v < 0) {
                *p++ = '-';
                v = -v;
            }
            char t[6];
            int k = 0;
            do {
                t[k++] = (char)('0' + v % 10);
                v /= 10;
            } while (v);
            while (k) *p++ = t[--k];
        }
        out.resize(p - &out[0]);
    }
    // auxiliary fields in header order (slow5.c:3880-3915)
    uint64_t a = 0;
    for (const auto &f : h.aux) {
        out += '\t';
        if (!rec.aux_bytes) {
            out += '.';
            continue;
        }
        if (!f.is_array()) {
            aux_prim_to_str(rec.aux_bytes + a, f.type, out);
            a += f.size;
            continue;
        }
        uint64_t len;
        memcpy(&len, rec.aux_bytes + a, 8);
        a += 8;
        if (len == 0) {
            out += '.';
        } else if (f.type == AUX_STRING) {
            // printed up to the first NUL like strdup() would (slow5.c:4565-4571)
            const char *s = reinterpret_cast<const char *>(rec.aux_bytes + a);
            out.append(s, strnlen(s, len));
        } else {
            const int prim = f.type - AUX_INT8_ARRAY;
            for (uint64_t i = 0; i < len; ++i) {
                if (i) out += ',';
                aux_prim_to_str(rec.aux_bytes + a + i * f.size, prim, out);
            }
        }
        a += len * f.size;
    }
    out += '\n';
}

void record_to_binary(const Record &rec, const uint8_t *signal, uint64_t signal_nbytes, bool signal_is_compressed,
                      std::vector<uint8_t> &out, uint64_t *signal_at) {
    const uint16_t rid_len = (uint16_t)rec.read_id.size();
    out.resize(2 + rid_len + 4 + 32 + 8 + signal_nbytes + rec.aux_nbytes);
    uint8_t *p = out.data();
    memcpy(p, &rid_len, 2);
    p += 2;
    memcpy(p, rec.read_id.data(), rid_len);
    p += rid_len;
    memcpy(p, &rec.read_group, 4);
    p += 4;
    memcpy(p, &rec.digitisation, 8);
    memcpy(p + 8, &rec.offset, 8);
    memcpy(p + 16, &rec.range, 8);
    memcpy(p + 24, &rec.sampling_rate, 8);
    p += 32;
    const uint64_t lrs = signal_is_compressed ? signal_nbytes : signal_nbytes / 2;
    memcpy(p, &lrs, 8);
    p += 8;
    if (signal_at) *signal_at = (uint64_t)(p - out.data());
    if (signal_nbytes) memcpy(p, signal, signal_nbytes);
    p += signal_nbytes;
    if (rec.aux_nbytes) memcpy(p, rec.aux_bytes, rec.aux_nbytes);
}

// ---- ASCII record parse (SLOW5 -> anything) --------------------------------------------------------
bool aux_null_value(int type, uint8_t *dst) {  // the type's NULL representation (slow5.h:139-150)
    switch (type) {
        case AUX_INT8: { int8_t v = INT8_MAX; memcpy(dst, &v, 1); return true; }
        case AUX_INT16: { int16_t v = INT16_MAX; memcpy(dst, &v, 2); return true; }
        case AUX_INT32: { int32_t v = INT32_MAX; memcpy(dst, &v, 4); return true; }
        case AUX_INT64: { int64_t v = INT64_MAX; memcpy(dst, &v, 8); return true; }
        case AUX_UINT8: case AUX_ENUM: { uint8_t v = UINT8_MAX; memcpy(dst, &v, 1); return true; }
        case AUX_UINT16: { uint16_t v = UINT16_MAX; memcpy(dst, &v, 2); return true; }
        case AUX_UINT32: { uint32_t v = UINT32_MAX; memcpy(dst, &v, 4); return true; }
        case AUX_UINT64: { uint64_t v = UINT64_MAX; memcpy(dst, &v, 8); return true; }
        case AUX_FLOAT: { float v = nanf(""); memcpy(dst, &v, 4); return true; }
        case AUX_DOUBLE: { double v = nan(""); memcpy(dst, &v, 8); return true; }
        case AUX_CHAR: { *dst = 0; return true; }
    }
    return false;
}

// The binary auxiliary section of a record laid out for another header: field p of `out` is field src_of_out[p] of `in`
// (copied as stored), or, when that is < 0, the "missing" form slow5_rec_to_mem writes (slow5.c:3993-4044): the type's
// NULL value for a primitive, a zero length for an array.
bool aux_relayout(const uint8_t *aux, uint64_t n, const std::vector<AuxField> &in, const std::vector<AuxField> &out,
                  const std::vector<int> &src_of_out, std::vector<uint8_t> &dst) {
    std::vector<std::pair<uint64_t, uint64_t>> at(in.size());  // (offset, bytes) of every stored field
    uint64_t pos = 0;
    for (size_t f = 0; f < in.size(); ++f) {
        uint64_t len = in[f].size;
        if (in[f].is_array()) {
            if (n - pos < 8) return false;
            uint64_t cnt;
            memcpy(&cnt, aux + pos, 8);
            if (cnt > (n - pos - 8) / (in[f].size ? in[f].size : 1)) return false;
            len = 8 + cnt * in[f].size;
        }
        if (n - pos < len) return false;
        at[f] = {pos, len};
        pos += len;
    }
    if (pos != n) return false;
    dst.clear();
    for (size_t p = 0; p < out.size(); ++p) {
        const int f = src_of_out[p];
        if (f >= 0) {
            dst.insert(dst.end(), aux + at[f].first, aux + at[f].first + at[f].second);
        } else if (out[p].is_array()) {
            dst.insert(dst.end(), 8, 0);
        } else {
            uint8_t v[8] = {0};
            if (!aux_null_value(out[p].type, v)) return false;
            dst.insert(dst.end(), v, v + out[p].size);
        }
    }
    return true;
}

static bool parse_prim(const std::string &tok, int type, uint8_t *dst) {
    if (tok == ".") return aux_null_value(type, dst);  // missing value
    if (tok.empty()) return false;
    char *end = nullptr;
    errno = 0;
    switch (type) {
        case AUX_INT8: case AUX_INT16: case AUX_INT32: case AUX_INT64: {
            const long long v = strtoll(tok.c_str(), &end, 10);
            if (*end || errno) return false;
            if (type == AUX_INT8) { if (v < INT8_MIN || v > INT8_MAX) return false; int8_t x = (int8_t)v; memcpy(dst, &x, 1); }
            else if (type == AUX_INT16) { if (v < INT16_MIN || v > INT16_MAX) return false; int16_t x = (int16_t)v; memcpy(dst, &x, 2); }
            else if (type == AUX_INT32) { if (v < INT32_MIN || v > INT32_MAX) return false; int32_t x = (int32_t)v; memcpy(dst, &x, 4); }
            else { int64_t x = v; memcpy(dst, &x, 8); }
            return true;
        }
        case AUX_UINT8: case AUX_UINT16: case AUX_UINT32: case AUX_UINT64: case AUX_ENUM: {
            if (tok[0] == '-') return false;
            const unsigned long long v = strtoull(tok.c_str(), &end, 10);
            if (*end || errno) return false;
            if (type == AUX_UINT8 || type == AUX_ENUM) { if (v > UINT8_MAX) return false; uint8_t x = (uint8_t)v; memcpy(dst, &x, 1); }
            else if (type == AUX_UINT16) { if (v > UINT16_MAX) return false; uint16_t x = (uint16_t)v; memcpy(dst, &x, 2); }
            else if (type == AUX_UINT32) { if (v > UINT32_MAX) return false; uint32_t x = (uint32_t)v; memcpy(dst, &x, 4); }
            else { uint64_t x = v; memcpy(dst, &x, 8); }
            return true;
        }
        case AUX_FLOAT: { const float v = strtof(tok.c_str(), &end); if (*end) return false; memcpy(dst, &v, 4); return true; }
        case AUX_DOUBLE: { const double v = strtod(tok.c_str(), &end); if (*end) return false; memcpy(dst, &v, 8); return true; }
        case AUX_CHAR: { *dst = (uint8_t)tok[0]; return true; }
    }
    return false;
}

bool record_parse_ascii(const char *line, uint64_t n, const Header &h, Record &rec, std::vector<uint8_t> &aux_store,
                        std::string &err, bool defer_signal) {
    std::vector<std::string> col;
    split(std::string(line, n), '\t', col);
    if (col.size() != 8 + h.aux.size()) {
        err = "slow5 record has the wrong number of columns";
        return false;
    }
    char *end = nullptr;
    rec.read_id = col[0];
    rec.read_group = (uint32_t)strtoul(col[1].c_str(), &end, 10);
    if (*end) goto bad;
    rec.digitisation = strtod(col[2].c_str(), &end);
    if (*end) goto bad;
    rec.offset = strtod(col[3].c_str(), &end);
    if (*end) goto bad;
    rec.range = strtod(col[4].c_str(), &end);
    if (*end) goto bad;
    rec.sampling_rate = strtod(col[5].c_str(), &end);
    if (*end) goto bad;
    rec.len_raw_signal = strtoull(col[6].c_str(), &end, 10);
    if (*end) goto bad;
    if (defer_signal) {
        // the caller converts the column in a batch (s5b_ascii_to_signal_batch_host): hand out where it sits in `line`
        rec.raw_signal.clear();
        uint64_t at = 0;
        for (int tabs = 0; tabs < 7 && at < n; ++at)
            if (line[at] == '\t') ++tabs;
        rec.sig_bytes = reinterpret_cast<const uint8_t *>(line + at);
        rec.sig_nbytes = col[7].size();
    } else {
        rec.raw_signal.clear();
        rec.raw_signal.reserve(rec.len_raw_signal);
        const char *p = col[7].c_str();
        while (*p) {
            const long v = strtol(p, &end, 10);
            if (end == p || v < INT16_MIN || v > INT16_MAX) goto bad;
            rec.raw_signal.push_back((int16_t)v);
            p = end;
            if (*p == ',') ++p;
            else if (*p) goto bad;
        }
        if (rec.raw_signal.size() != rec.len_raw_signal) {
            err = "slow5 record: len_raw_signal does not match the number of samples";
            return false;
        }
    }
    aux_store.clear();
    for (size_t i = 0; i < h.aux.size(); ++i) {
        const AuxField &f = h.aux[i];
        const std::string &tok = col[8 + i];
        if (!f.is_array()) {
            const size_t at = aux_store.size();
            aux_store.resize(at + f.size);
            if (!parse_prim(tok, f.type, &aux_store[at])) goto bad;
            continue;
        }
        uint64_t len = 0;
        const size_t len_at = aux_store.size();
        aux_store.resize(len_at + 8);
        if (tok == ".") {
            len = 0;
        } else if (f.type == AUX_STRING) {
            len = tok.size();
            aux_store.insert(aux_store.end(), tok.begin(), tok.end());
        } else {
            std::vector<std::string> el;
            split(tok, ',', el);
            const int prim = f.type - AUX_INT8_ARRAY;
            for (const auto &e : el) {
                const size_t at = aux_store.size();
                aux_store.resize(at + f.size);
                if (!parse_prim(e, prim, &aux_store[at])) goto bad;
            }
            len = el.size();
        }
        memcpy(&aux_store[len_at], &len, 8);
    }
    rec.aux_bytes = aux_store.data();
    rec.aux_nbytes = aux_store.size();
    if (!defer_signal) {
        rec.sig_bytes = reinterpret_cast<const uint8_t *>(rec.raw_signal.data());
        rec.sig_nbytes = rec.raw_signal.size() * 2;
    }
    return true;
bad:
    err = "slow5 record does not parse";
    return false;
}

}  // namespace s5b
