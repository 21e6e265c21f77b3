// merge_split_main.cpp -- `slow5tools-b200 merge` and `slow5tools-b200 split`: the two other heavy callers of the reference's
// batch worker.  Both are the view loop with one more step per record:
//   merge (src/merge.c:72-452)   several inputs -> one output; read groups are united by run_id (merge.c:283-317), a record's
//                                read_group is renumbered (merge.c:52) and its auxiliary section is laid out for the union
//                                of the inputs' columns (enum columns first, the others sorted by name: merge.c:231-271,325-332);
//   split (src/split.c:110-660)  one input -> several outputs: by read group (-g: every record goes to the file of its group
//                                with read_group 0, split.c:88-89,520), by record count (-r), into -f files (split.c:379-456),
//                                or by the categories a table gives the read ids (-x: demultiplexing, src/demux.c).
// Records go through convert_records (view_main.cpp): decompression and compression are batch calls on the GPU; there is no
// CPU codec here.  What is not carried over: lossy output is (--lossless false), and an
// auxiliary column that changes TYPE between inputs is refused (the reference writes such records with the bytes of one type
// under the header of another, merge.c:262 / test 4.3).
#include <dirent.h>
#include <getopt.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "cli_common.hpp"

using namespace s5b;

#define MS_ERROR(fmt, ...) fprintf(stderr, "[%s::ERROR]\033[1;31m " fmt "\033[0m\n", __func__, __VA_ARGS__)
#define MS_WARNING(fmt, ...) fprintf(stderr, "[%s::WARNING]\033[1;33m " fmt "\033[0m\n", __func__, __VA_ARGS__)
#define MS_INFO(fmt, ...) fprintf(stderr, "[%s::INFO]\033[1;34m " fmt "\033[0m\n", __func__, __VA_ARGS__)

namespace {

bool is_directory(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool has_ext(const std::string &p, const char *ext) {
    const size_t n = strlen(ext);
    return p.size() >= n && p.compare(p.size() - n, n, ext) == 0;
}
// every *.slow5 / *.blow5 under path, directories searched recursively in readdir order (list_all_items, src/read_fast5.c:62-130)
void list_inputs(const std::string &path, std::vector<std::string> &files) {
    if (is_directory(path)) {
        DIR *d = opendir(path.c_str());
        if (!d) return;
        std::vector<std::string> names;
        while (struct dirent *e = readdir(d)) names.push_back(e->d_name);
        closedir(d);
        for (const std::string &fn : names) {
            if (fn == "." || fn == "..") continue;
            list_inputs(path + "/" + fn, files);
        }
    } else if (has_ext(path, ".slow5") || has_ext(path, ".blow5")) {
        files.push_back(path);
    }
}

struct Opts {
    const char *arg_out = nullptr, *arg_to = nullptr, *arg_rec = nullptr, *arg_sig = nullptr, *arg_dir = nullptr;
    int threads = 8;
    long batch = 4096;
    bool lossy = false, allow = false;
    Fmt fmt_out = FMT_UNKNOWN;
    int rec_out = PRESS_ZLIB, sig_out = PRESS_SVB_ZD;  // misc.c:54-55
};

bool parse_lossless(const char *v, bool &lossy) {  // parse_arg_lossless, src/misc.c
    if (!strcmp(v, "true")) lossy = false;
    else if (!strcmp(v, "false")) lossy = true;
    else {
        MS_ERROR("Incorrect argument '%s' for --lossless: true or false", v);
        return false;
    }
    return true;
}

// output format and compression from --to / -o / -c / -s (parse_format_args, auto_detect_formats, parse_compression_opts: src/misc.c)
bool resolve_output(Opts &o, Fmt fallback) {
    if (o.arg_to && (o.fmt_out = fmt_from_name(o.arg_to)) == FMT_UNKNOWN) {
        MS_ERROR("invalid output format '%s'", o.arg_to);
        return false;
    }
    if (o.arg_out) {
        const Fmt by_ext = fmt_from_path(o.arg_out);
        if (o.fmt_out == FMT_UNKNOWN) {
            o.fmt_out = by_ext;
            if (o.fmt_out == FMT_UNKNOWN) {
                MS_ERROR("cannot detect the output format from the file extension of '%s'", o.arg_out);
                return false;
            }
        } else if (by_ext != FMT_UNKNOWN && by_ext != o.fmt_out) {
            MS_ERROR("output file extension '%s' does not match the output format '%s'", o.arg_out, o.arg_to);
            return false;
        }
    }
    if (o.fmt_out == FMT_UNKNOWN) o.fmt_out = fallback;
    if (o.fmt_out == FMT_ASCII && (o.arg_rec || o.arg_sig)) {
        MS_ERROR("%s", "compression options (-c / -s) are only valid for blow5 output");
        return false;
    }
    if (o.arg_rec && (o.rec_out = press_from_name(o.arg_rec)) == PRESS_BAD) {
        MS_ERROR("invalid record compression method '%s'", o.arg_rec);
        return false;
    }
    if (o.arg_sig && (o.sig_out = press_from_name(o.arg_sig)) == PRESS_BAD) {
        MS_ERROR("invalid signal compression method '%s'", o.arg_sig);
        return false;
    }
    if (o.fmt_out == FMT_ASCII) o.rec_out = o.sig_out = PRESS_NONE;
    if ((o.rec_out != PRESS_NONE && o.rec_out != PRESS_ZLIB && o.rec_out != PRESS_ZSTD) ||
        (o.sig_out != PRESS_NONE && o.sig_out != PRESS_SVB_ZD && o.sig_out != PRESS_EX_ZD)) {
        MS_ERROR("%s", "this build supports record compression none/zlib/zstd and signal compression none/svb-zd/ex-zd only");
        return false;
    }
    return true;
}

bool input_supported(const Header &h, const std::string &path) {
    if ((h.record_method != PRESS_NONE && h.record_method != PRESS_ZLIB && h.record_method != PRESS_ZSTD) ||
        (h.signal_method != PRESS_NONE && h.signal_method != PRESS_SVB_ZD && h.signal_method != PRESS_EX_ZD)) {
        MS_ERROR("%s uses a compression method this build does not support (zlib/zstd as signal method)", path.c_str());
        return false;
    }
    return true;
}

s5b_ctx_t *open_gpu() {
    s5b_ctx_t *gpu = nullptr;
    const int rc = s5b_ctx_create(-1, &gpu);
    if (rc != S5B_OK) {
        MS_ERROR("cannot initialise the GPU codec: %s", s5b_strerror(rc));
        return nullptr;
    }
    return gpu;
}

// ---- header attributes -------------------------------------------------------------------------------------------------------
// The reference keeps, per read group, a hash map key -> value that holds only the keys that group was given (slow5_hdr_get_data),
// next to the set of all keys; an emptied value ("") is still a key of its group.  Header::attrs is key -> value per group with
// "" for "nothing to print", so the output header carries a second table saying which (key, group) pairs exist.  A header read
// from a file lists every key for every group: there every pair exists.
struct OutHeader {
    Header h;
    std::vector<std::vector<char>> has;  // parallel to h.attrs
};
int attr_find(const Header &h, const std::string &key) {
    for (size_t i = 0; i < h.attrs.size(); ++i)
        if (h.attrs[i].first == key) return (int)i;
    return -1;
}
const std::string *attr_get(const Header &h, const std::string &key, uint32_t rg) {
    const int i = attr_find(h, key);
    return i >= 0 && rg < h.attrs[i].second.size() ? &h.attrs[i].second[rg] : nullptr;
}
// slow5_hdr_add_attr: a new key, without a value in any read group
int attr_add(OutHeader &o, const std::string &key) {
    const int i = attr_find(o.h, key);
    if (i >= 0) return i;
    o.h.attrs.emplace_back(key, std::vector<std::string>(o.h.num_read_groups));
    o.has.emplace_back(o.h.num_read_groups, 0);
    return (int)o.h.attrs.size() - 1;
}
// slow5_hdr_add_rg_data: a new read group holding the attributes of read group `rg` of `src`
uint32_t rg_append(OutHeader &o, const Header &src, uint32_t rg) {
    const uint32_t g = o.h.num_read_groups++;
    for (auto &kv : o.h.attrs) kv.second.resize(o.h.num_read_groups);
    for (auto &v : o.has) v.resize(o.h.num_read_groups, 0);
    for (const auto &kv : src.attrs) {
        const int i = attr_add(o, kv.first);
        if (rg < kv.second.size()) o.h.attrs[i].second[g] = kv.second[rg];
        o.has[i][g] = 1;
    }
    return g;
}

// compare_headers, src/merge.c:454-544: read group `og` of the output against read group `ig` of an input with the same run_id.
// Differences are warned about and the attribute is emptied in the output; returns whether there were any.
bool compare_rg(OutHeader &o, uint32_t og, const Header &in, uint32_t ig, const char *path, const char *run_id) {
    bool warned = false;
    Header &out = o.h;
    size_t n_out = 0;
    for (size_t a = 0; a < out.attrs.size(); ++a) n_out += o.has[a][og] != 0;
    const size_t n_in = in.attrs.size();
    if (n_in != n_out) {
        warned = true;
        MS_WARNING("Input file %s (run_id-%s) has a different number of attributes (%zu) than seen in the previous files processed so far (%zu)",
                   path, run_id, n_in, n_out);
    }
    for (size_t a = 0; a < out.attrs.size(); ++a) {
        if (!o.has[a][og]) continue;
        const std::string &key = out.attrs[a].first;
        const int ia = attr_find(in, key);
        bool clear = false;
        if (ia < 0) {
            MS_WARNING("Attribute '%s' is not available in input file %s (run_id-%s)", key.c_str(), path, run_id);
            clear = true;
        } else if (in.attrs[ia].second[ig] != out.attrs[a].second[og]) {
            MS_WARNING("Attribute '%s' in input file %s (run_id-%s) has a different value (%s) than what has been seen so far (%s)",
                       key.c_str(), path, run_id, in.attrs[ia].second[ig].c_str(), out.attrs[a].second[og].c_str());
            clear = true;
        }
        if (clear) {
            warned = true;
            MS_INFO("Setting output header's attribute '%s' (run_id-%s) to empty", key.c_str(), run_id);
            out.attrs[a].second[og].clear();
        }
    }
    for (size_t a = 0; a < in.attrs.size(); ++a) {
        const int oa = attr_find(out, in.attrs[a].first);
        if (oa < 0 || !o.has[oa][og]) {
            warned = true;
            MS_INFO("Attribute '%s' in input file %s (run_id-%s) is not seen in previous files. It will be added to the output header but its value (%s) will not be set in the output header.",
                    in.attrs[a].first.c_str(), path, run_id, in.attrs[a].second[ig].c_str());
            attr_add(o, in.attrs[a].first);
        }
    }
    return warned;
}

bool is_enum(const AuxField &f) { return f.type == AUX_ENUM || f.type == AUX_ENUM_ARRAY; }

struct Output {
    FILE *fp = nullptr;
    std::string path;
};
bool finish_output(FILE *fp, Fmt fmt, bool close_it) {
    bool ok = fflush(fp) == 0;
    if (ok && fmt == FMT_BINARY && fwrite("5WOLB", 1, 5, fp) != 5) ok = false;  // slow5_eof_fwrite
    if (close_it) {
        if (fclose(fp) != 0) ok = false;
    } else if (fflush(fp) != 0) {
        ok = false;
    }
    return ok;
}

}  // namespace

int merge_main(int argc, char **argv) {
    static const struct option long_opts[] = {
        {"help", no_argument, nullptr, 'h'},           {"threads", required_argument, nullptr, 't'},
        {"to", required_argument, nullptr, 'b'},       {"compress", required_argument, nullptr, 'c'},
        {"sig-compress", required_argument, nullptr, 's'}, {"lossless", required_argument, nullptr, 'l'},
        {"allow", no_argument, nullptr, 'a'},          {"output", required_argument, nullptr, 'o'},
        {"batchsize", required_argument, nullptr, 'K'}, {nullptr, 0, nullptr, 0}};
    Opts o;
    int opt;
    optind = 1;
    while ((opt = getopt_long(argc, argv, "c:s:ht:o:aK:l:b:", long_opts, nullptr)) != -1) {
        switch (opt) {
            case 'c': o.arg_rec = optarg; break;
            case 's': o.arg_sig = optarg; break;
            case 'a':
                o.allow = true;
                MS_WARNING("%s", "You have requested to merge files despite attribute differences in same read ID. Generated files are for intermediate analysis and are not recommended for archiving.");
                break;
            case 't': o.threads = atoi(optarg); break;
            case 'o': o.arg_out = optarg; break;
            case 'K': o.batch = atol(optarg); break;
            case 'b': o.arg_to = optarg; break;
            case 'l':
                if (!parse_lossless(optarg, o.lossy)) return 1;
                break;
            case 'h':
                printf("Usage: slow5tools-b200 merge [OPTIONS] [SLOW5_FILE/DIR] ...\nMerge multiple SLOW5/BLOW5 files to a single file (GPU codec).\n\n"
                       "OPTIONS:\n    --to FORMAT, -o FILE, -c REC_MTD, -s SIG_MTD, -t INT, -K INT as for view\n"
                       "    -l, --lossless STR            retain information in auxiliary fields during the conversion [true]\n"
                       "    -a, --allow                   allow merging despite attribute differences in the same run_id\n");
                return 0;
            default: return 1;
        }
    }
    if (o.threads < 1 || o.batch < 1) {
        MS_ERROR("%s", "invalid -t / -K value");
        return 1;
    }
    if (!resolve_output(o, FMT_BINARY)) return 1;  // merge.c: stdout defaults to blow5
    if (optind >= argc) {
        MS_ERROR("Not enough arguments. Enter one or more slow5/blow5 files or directories as arguments.%s", "");
        return 1;
    }
    std::vector<std::string> files;
    for (int i = optind; i < argc; ++i) list_inputs(argv[i], files);
    if (files.empty()) {
        MS_ERROR("No slow5/blow5 files found. Exiting.%s", "");
        return 1;
    }
    FILE *fout = stdout;
    if (o.arg_out && !(fout = fopen(o.arg_out, "wb"))) {
        MS_ERROR("File '%s' could not be opened - %s.", o.arg_out, strerror(errno));
        return 1;
    }
    setvbuf(fout, nullptr, _IOFBF, 1 << 20);

    // ---- pass 1 over the headers: the output's read groups, attributes and auxiliary columns (merge.c:208-332)
    OutHeader oh;
    Header &out = oh.h;
    out.version[0] = 0, out.version[1] = 2, out.version[2] = 0;  // slow5_init_empty: the library's own version
    out.num_read_groups = 0;
    std::vector<std::vector<uint32_t>> rg_map;   // per input: its read group j -> the output's
    std::vector<std::string> inputs;
    std::vector<AuxField> enum_cols;             // in the order met
    std::map<std::string, AuxField> other_cols;  // sorted by name
    bool warned = false;
    for (const std::string &path : files) {
        Reader rd;
        if (!reader_open(rd, path.c_str(), FMT_UNKNOWN)) {
            MS_ERROR("[Skip file]: cannot open %s. skipping.\n", path.c_str());
            continue;
        }
        const Header &h = rd.hdr;
        if (!input_supported(h, path)) return 1;
        if (!o.lossy && h.aux.empty()) {
            MS_ERROR("%s has no auxiliary fields. Specify -l false to merge files with no auxiliary fields.", path.c_str());
            return 1;
        }
        if (!o.lossy) {
            for (const AuxField &f : h.aux) {
                if (is_enum(f)) {
                    const AuxField *seen = nullptr;
                    for (const AuxField &e : enum_cols)
                        if (e.name == f.name) seen = &e;
                    if (!seen) {
                        enum_cols.push_back(f);
                    } else if (seen->type_str != f.type_str) {
                        // the label lists are the part of the type string between the braces (slow5.c:1159-1258)
                        const size_t na = std::count(seen->type_str.begin(), seen->type_str.end(), ',');
                        const size_t nb = std::count(f.type_str.begin(), f.type_str.end(), ',');
                        if (na != nb) MS_ERROR("Attribute %s has different number of enum labels in different files", f.name.c_str());
                        else MS_ERROR("Attribute %s has different order/name of the enum labels in different files", f.name.c_str());
                        return 1;
                    }
                } else {
                    auto it = other_cols.find(f.name);
                    if (it == other_cols.end()) {
                        other_cols.emplace(f.name, f);
                    } else if (it->second.type != f.type) {
                        MS_ERROR("Auxiliary field '%s' has type %s in %s and %s in an earlier file: not supported", f.name.c_str(),
                                 f.type_str.c_str(), path.c_str(), it->second.type_str.c_str());
                        return 1;
                    }
                }
            }
        }
        std::vector<uint32_t> map(h.num_read_groups);
        for (uint32_t j = 0; j < h.num_read_groups; ++j) {
            const std::string *rid = attr_get(h, "run_id", j);
            if (!rid || rid->empty()) {
                MS_ERROR("No run_id found in %s.", path.c_str());
                return 1;
            }
            bool found = false;
            for (uint32_t k = 0; k < out.num_read_groups; ++k) {
                const std::string *rk = attr_get(out, "run_id", k);
                if (rk && *rk == *rid) {
                    found = true;
                    map[j] = k;
                    warned = compare_rg(oh, k, h, j, path.c_str(), rid->c_str());  // (the last comparison decides, merge.c:300)
                    break;
                }
            }
            if (!found) map[j] = rg_append(oh, h, j);
        }
        rg_map.push_back(map);
        inputs.push_back(path);
        reader_close(rd);
    }
    if (warned && !o.allow) {
        MS_ERROR("Attributes are different for the same run_id(s). Set -a of you still want to merge files%s", ".");
        return 1;
    }
    if (inputs.empty()) {
        MS_ERROR("No slow5/blow5 files found for conversion. Exiting.%s", "");
        return 1;
    }
    for (const AuxField &e : enum_cols) {
        if (other_cols.count(e.name)) {  // slow5_aux_meta_add fails on the second definition
            MS_ERROR("Could not initialize the record attribute '%s'", e.name.c_str());
            return 1;
        }
        out.aux.push_back(e);
    }
    for (const auto &kv : other_cols) out.aux.push_back(kv.second);

    {
        const std::string h = header_to_mem(out, o.fmt_out, o.rec_out, o.sig_out);
        if (fwrite(h.data(), 1, h.size(), fout) != h.size()) {
            MS_ERROR("Could not write the header to %s\n", o.arg_out ? o.arg_out : "stdout");
            return 1;
        }
    }

    // ---- pass 2: the records, one input after the other
    s5b_ctx_t *gpu = nullptr;
    int ret = 0;
    for (size_t fi = 0; fi < inputs.size() && ret == 0; ++fi) {
        Reader rd;
        if (!reader_open(rd, inputs[fi].c_str(), FMT_UNKNOWN)) {
            MS_ERROR("File '%s' could not be opened - %s.", inputs[fi].c_str(), rd.err.c_str());
            return 1;
        }
        const Header &h = rd.hdr;
        const bool need_gpu = h.record_method != PRESS_NONE || h.signal_method != PRESS_NONE || o.rec_out != PRESS_NONE ||
                              o.sig_out != PRESS_NONE;
        if (need_gpu && !gpu && !(gpu = open_gpu())) return 1;
        // where every output column comes from in this input
        std::vector<int> src(out.aux.size(), -1);
        for (size_t p = 0; p < out.aux.size(); ++p)
            for (size_t f = 0; f < h.aux.size(); ++f)
                if (h.aux[f].name == out.aux[p].name) src[p] = (int)f;
        bool same_layout = src.size() == h.aux.size();
        for (size_t p = 0; same_layout && p < src.size(); ++p) same_layout = src[p] == (int)p;
        const std::vector<uint32_t> &map = rg_map[fi];
        if (rd.fmt == FMT_BINARY && o.fmt_out == FMT_BINARY && need_gpu && !o.lossy && same_layout && h.aux.size() <= 64 &&
            !getenv("S5B_VIEW_SLOW_PATH")) {
            // blow5 -> blow5 and nothing to re-lay: whole batches stay on the device, which renumbers the read groups and
            // checks every record against the header (read group in range, auxiliary section = the header's columns)
            std::vector<uint8_t> sz(h.aux.size() + 1), arr(h.aux.size() + 1);
            for (size_t f = 0; f < h.aux.size(); ++f) {
                sz[f] = h.aux[f].size;
                arr[f] = h.aux[f].is_array() ? 1 : 0;
            }
            if (s5b_ctx_set_aux_layout(gpu, sz.data(), arr.data(), (uint32_t)h.aux.size()) != S5B_OK ||
                s5b_ctx_set_rg_map(gpu, map.data(), (uint32_t)map.size()) != S5B_OK) {
                MS_ERROR("%s", "cannot hand the header's layout to the GPU codec");
                return 1;
            }
            ret = blow5_fast_convert(rd, fout, gpu, o.rec_out, o.sig_out);
            s5b_ctx_set_rg_map(gpu, nullptr, 0);
            reader_close(rd);
            continue;
        }
        ConvertHooks hooks;
        hooks.hdr_out = &out;
        hooks.transform = [&](size_t, Record &rec, std::vector<uint8_t> &aux_store) {
            if (rec.read_group >= map.size()) {
                MS_ERROR("read group %u of record %s is not in the header of %s", rec.read_group, rec.read_id.c_str(), inputs[fi].c_str());
                return false;
            }
            rec.read_group = map[rec.read_group];
            if (o.lossy) {
                rec.aux_bytes = nullptr;
                rec.aux_nbytes = 0;
            } else if (!same_layout) {
                std::vector<uint8_t> laid;
                if (!aux_relayout(rec.aux_bytes, rec.aux_nbytes, h.aux, out.aux, src, laid)) {
                    MS_ERROR("auxiliary section of record %s does not match the header of %s", rec.read_id.c_str(), inputs[fi].c_str());
                    return false;
                }
                aux_store.swap(laid);
                rec.aux_bytes = aux_store.data();
                rec.aux_nbytes = aux_store.size();
            }
            return true;
        };
        ret = convert_records(h, rd.fmt, [&](std::vector<uint8_t> &mem) {
            const int rc = reader_next_mem(rd, mem);
            if (rc < 0) MS_ERROR("%s", rd.err.c_str());
            return rc;
        }, fout, gpu, o.fmt_out, o.rec_out, o.sig_out, o.batch, o.threads, &hooks);
        reader_close(rd);
    }
    if (!finish_output(fout, o.fmt_out, fout != stdout)) ret = 1;
    return ret;
}

namespace {

// ---- demultiplexing (src/demux.c) ---------------------------------------------------------------------------------------------
// A tab-separated table (guppy / dorado barcoding summary, or any table with a read id column and a category column) says which
// category -- barcode -- every read belongs to.  Each category gets its own output file, named after the input with "_<category>"
// in front of the extension (path_spawn, demux.c:261-279), created when its first record arrives (demux_write, :562-590) and
// carrying the input's whole header (slow5_birth, :1100-1130).  A read listed under several categories is written to each of
// them, or only to the -u category when one is named (update_db, :815-836); a read the table does not list is dropped with a
// warning, or goes to the -m category (:842-892); a table that lists more reads than the file holds is an error (:520-523).
struct DemuxOpts {
    const char *table = nullptr;
    const char *code_col = "barcode_arrangement", *rid_col = "parent_read_id";  // demux.h:8-9
    const char *multi = nullptr, *missing = nullptr;
};
struct DemuxPlan {
    std::vector<std::string> names;                        // categories in order of first appearance, then multi, then missing
    std::map<std::string, std::vector<uint16_t>> of_read;  // read id -> categories, each once, in table order
    int multi = -1, missing = -1;                          // their indices in names, -1: not asked for
};

// the non-empty tab-separated fields of a line (the reference walks the line with strtok: runs of tabs count once)
void tab_fields(const std::string &line, std::vector<std::string> &out) {
    out.clear();
    size_t at = 0;
    while (at < line.size()) {
        const size_t end = std::min(line.find('\t', at), line.size());
        if (end > at) out.push_back(line.substr(at, end - at));
        at = end + 1;
    }
}

bool demux_read_table(const DemuxOpts &d, DemuxPlan &plan) {
    FILE *fp = fopen(d.table, "r");
    if (!fp) {
        MS_ERROR("Failed to open '%s': %s", d.table, strerror(errno));
        return false;
    }
    std::string text;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, got);
    fclose(fp);
    std::vector<std::string> f;
    size_t at = 0, rid_pos = 0, code_pos = 0;  // 1-based field numbers, 0: not found
    bool first = true;
    std::map<std::string, uint16_t> index_of;
    while (at < text.size() || first) {
        const size_t end = std::min(text.find('\n', at), text.size());
        tab_fields(text.substr(at, end - at), f);
        at = end + 1;
        if (first) {  // bsum_parsehdr, demux.c:431-470: the first column of either name counts
            first = false;
            for (size_t i = 0; i < f.size() && (!rid_pos || !code_pos); ++i) {
                if (!rid_pos && f[i] == d.rid_col) rid_pos = i + 1;
                else if (!code_pos && f[i] == d.code_col) code_pos = i + 1;
            }
            if (!rid_pos) {
                MS_ERROR("Invalid demux TSV header: missing '%s'", d.rid_col);
                return false;
            }
            if (!code_pos) {
                MS_ERROR("Invalid demux TSV header: missing '%s'", d.code_col);
                return false;
            }
            continue;
        }
        if (f.size() < rid_pos || f.size() < code_pos) continue;  // (a blank or short line names nothing)
        const std::string &rid = f[rid_pos - 1], &code = f[code_pos - 1];
        auto it = index_of.find(code);
        if (it == index_of.end()) {
            if (plan.names.size() >= 65533) {
                MS_ERROR("Too many categories (%zu)", plan.names.size() + 3);
                return false;
            }
            it = index_of.emplace(code, (uint16_t)plan.names.size()).first;
            plan.names.push_back(code);
        }
        std::vector<uint16_t> &v = plan.of_read[rid];
        if (std::find(v.begin(), v.end(), it->second) == v.end()) v.push_back(it->second);
    }
    // the two special categories must not collide with a category of the table or with each other (getcodes, demux.c:226-255)
    if (d.multi) {
        if (index_of.count(d.multi)) {
            MS_ERROR("Multi-category '%s' already exists in demux TSV", d.multi);
            return false;
        }
        plan.multi = (int)plan.names.size();
        plan.names.push_back(d.multi);
    }
    if (d.missing) {
        if (index_of.count(d.missing) || (d.multi && !strcmp(d.multi, d.missing))) {
            MS_ERROR("Uncategorised reads category '%s' already exists", d.missing);
            return false;
        }
        plan.missing = (int)plan.names.size();
        plan.names.push_back(d.missing);
    }
    return true;
}

// <dir>/<input file name with "_<category>" in front of its extension>, extension = the output format's
std::string demux_path(const std::string &in_path, const char *dir, const std::string &category, Fmt fmt) {
    const size_t slash = in_path.find_last_of('/');
    std::string name = in_path.substr(slash == std::string::npos ? 0 : slash + 1);
    const size_t dot = name.find_last_of('.');
    name = name.substr(0, dot) + "_" + category + (fmt == FMT_ASCII ? ".slow5" : ".blow5");
    return dir && *dir ? std::string(dir) + "/" + name : name;
}

int demux_file(const std::string &path, Reader &rd, const Opts &o, const DemuxPlan &plan_in, s5b_ctx_t *gpu) {
    DemuxPlan plan = plan_in;  // (-m adds the reads it catches to the table, demux.c:876-885)
    const Header &h = rd.hdr;
    Header hdr_out = h;
    if (o.lossy) hdr_out.aux.clear();
    std::vector<Output> outs(plan.names.size());
    std::vector<std::vector<FILE *>> dest;  // of the records of the batch being converted
    uint64_t n_records = 0;
    bool failed = false;
    auto file_of = [&](uint16_t k) -> FILE * {
        Output &out = outs[k];
        if (out.fp) return out.fp;
        out.path = demux_path(path, o.arg_dir, plan.names[k], o.fmt_out);
        out.fp = fopen(out.path.c_str(), "wb");
        if (!out.fp) {
            MS_ERROR("Failed to open '%s' for writing: %s", out.path.c_str(), strerror(errno));
            return nullptr;
        }
        setvbuf(out.fp, nullptr, _IOFBF, 1 << 20);
        const std::string hm = header_to_mem(hdr_out, o.fmt_out, o.rec_out, o.sig_out);
        if (fwrite(hm.data(), 1, hm.size(), out.fp) != hm.size()) {
            MS_ERROR("Failed to write the header to '%s'", out.path.c_str());
            return nullptr;
        }
        return out.fp;
    };
    ConvertHooks hooks;
    hooks.hdr_out = &hdr_out;
    hooks.transform = [&](size_t i, Record &rec, std::vector<uint8_t> &) {
        if (dest.size() <= i) dest.resize(i + 1);
        dest[i].clear();
        ++n_records;
        if (o.lossy) rec.aux_bytes = nullptr, rec.aux_nbytes = 0;
        auto it = plan.of_read.find(rec.read_id);
        std::vector<uint16_t> cats;
        if (it == plan.of_read.end()) {
            if (plan.missing < 0) {
                MS_WARNING("Read ID '%s' is missing from demux TSV", rec.read_id.c_str());
                return true;  // written nowhere
            }
            cats = plan.of_read[rec.read_id] = {(uint16_t)plan.missing};
        } else if (plan.multi >= 0 && it->second.size() > 1) {
            cats = {(uint16_t)plan.multi};
        } else {
            cats = it->second;
        }
        for (uint16_t k : cats) {
            FILE *f = file_of(k);
            if (!f) {
                failed = true;
                return false;
            }
            dest[i].push_back(f);
        }
        return true;
    };
    hooks.route_many = [&](size_t i) { return &dest[i]; };
    int ret = convert_records(h, rd.fmt, [&](std::vector<uint8_t> &mem) {
        const int rc = reader_next_mem(rd, mem);
        if (rc < 0) MS_ERROR("Could not read file %s", path.c_str());
        return rc;
    }, nullptr, gpu, o.fmt_out, o.rec_out, o.sig_out, o.batch, o.threads, &hooks);
    if (ret == 0 && n_records < plan.of_read.size()) {
        MS_ERROR("Extra read(s) in demux TSV%s", "");
        ret = 1;
    }
    for (Output &out : outs)
        if (out.fp && !finish_output(out.fp, o.fmt_out, true)) ret = 1;
    return ret || failed ? 1 : 0;
}

}  // namespace

int split_main(int argc, char **argv) {
    static const struct option long_opts[] = {
        {"help", no_argument, nullptr, 'h'},           {"to", required_argument, nullptr, 'b'},
        {"compress", required_argument, nullptr, 'c'}, {"sig-compress", required_argument, nullptr, 's'},
        {"out-dir", required_argument, nullptr, 'd'},  {"threads", required_argument, nullptr, 't'},
        {"lossless", required_argument, nullptr, 'l'}, {"groups", no_argument, nullptr, 'g'},
        {"files", required_argument, nullptr, 'f'},    {"reads", required_argument, nullptr, 'r'},
        {"batchsize", required_argument, nullptr, 'K'}, {"demux", required_argument, nullptr, 'x'},
        {"demux-code", required_argument, nullptr, 1000}, {"demux-rid", required_argument, nullptr, 1001},
        {"demux-uniq", required_argument, nullptr, 'u'}, {"demux-missing", required_argument, nullptr, 'm'},
        {nullptr, 0, nullptr, 0}};
    Opts o;
    DemuxOpts dx;
    enum { BY_READS, BY_FILES, BY_GROUPS, BY_TABLE } how = BY_READS;
    long count = 0;
    int opt;
    optind = 1;
    while ((opt = getopt_long(argc, argv, "hb:c:s:gl:f:r:d:t:K:x:u:m:", long_opts, nullptr)) != -1) {
        switch (opt) {
            case 'b': o.arg_to = optarg; break;
            case 'c': o.arg_rec = optarg; break;
            case 's': o.arg_sig = optarg; break;
            case 'd': o.arg_dir = optarg; break;
            case 't': o.threads = atoi(optarg); break;
            case 'K': o.batch = atol(optarg); break;
            case 'g': how = BY_GROUPS; break;
            case 'f': how = BY_FILES; count = atol(optarg); break;
            case 'r': how = BY_READS; count = atol(optarg); break;
            case 'l':
                if (!parse_lossless(optarg, o.lossy)) return 1;
                break;
            case 'x': how = BY_TABLE; dx.table = optarg; break;
            case 'u': dx.multi = optarg; break;
            case 'm': dx.missing = optarg; break;
            case 1000: dx.code_col = optarg; break;
            case 1001: dx.rid_col = optarg; break;
            case 'h':
                printf("Usage: slow5tools-b200 split [OPTIONS] [SLOW5_FILE/DIR] ...\nSplit a single SLOW5/BLOW5 file into multiple separate files (GPU codec).\n\n"
                       "OPTIONS:\n    -d, --out-dir DIR             output to directory DIR\n    -g, --groups                  split multi read group file into single read group files\n"
                       "    -r, --reads INT               split into INT reads per file\n    -f, --files INT               split reads into INT files evenly\n"
                       "    -x, --demux TSV_PATH          split reads according to TSV file\n        --demux-code STR          categories column name ['barcode_arrangement']\n"
                       "        --demux-rid STR           read IDs column name ['parent_read_id']\n    -m, --demux-missing STR       uncategorised reads to category named STR\n"
                       "    -u, --demux-uniq STR          multi-category reads to category named STR\n"
                       "    --to FORMAT, -c REC_MTD, -s SIG_MTD, -t INT, -K INT, -l STR as for merge\n");
                return 0;
            default: return 1;
        }
    }
    if (o.threads < 1 || o.batch < 1) {
        MS_ERROR("%s", "invalid -t / -K value");
        return 1;
    }
    if (!resolve_output(o, FMT_BINARY)) return 1;
    if (how == BY_READS && count <= 0) {
        MS_ERROR("Default splitting method - reads split is used. Specify the number of reads to include in a slow5 file%s", "");
        return 1;
    }
    if (how == BY_FILES && count <= 0) {
        MS_ERROR("Splitting method - files split is used. Specify the number of files to create from a slow5 file%s", "");
        return 1;
    }
    if (!o.arg_dir) {
        MS_ERROR("The output directory must be specified %s", "");
        return 1;
    }
    {
        struct stat st;
        if (stat(o.arg_dir, &st) == -1) {
            mkdir(o.arg_dir, 0700);
        } else {
            DIR *d = opendir(o.arg_dir);
            int entries = 0;
            if (d) {
                while (readdir(d)) ++entries;
                closedir(d);
            }
            if (entries > 2) {
                MS_ERROR("Output directory %s is not empty. Please remove it or specify another directory.", o.arg_dir);
                return 1;
            }
        }
    }
    std::vector<std::string> files;
    for (int i = optind; i < argc; ++i) list_inputs(argv[i], files);
    if (files.empty()) {
        MS_ERROR("No slow5/blow5 files found. Exiting...%s", "");
        return 1;
    }
    const std::string ext = o.fmt_out == FMT_ASCII ? ".slow5" : ".blow5";
    s5b_ctx_t *gpu = nullptr;
    DemuxPlan plan;
    bool have_plan = false;
    for (const std::string &path : files) {
        Reader rd;
        if (!reader_open(rd, path.c_str(), FMT_UNKNOWN)) {
            MS_ERROR("Cannot open %s. Skipping.\n", path.c_str());
            return 1;
        }
        const Header &h = rd.hdr;
        if (!input_supported(h, path)) return 1;
        if (h.num_read_groups == 1 && how == BY_GROUPS) {
            MS_ERROR("The file %s already has a single read group", path.c_str());
            return 1;
        }
        if (h.num_read_groups > 1 && how != BY_GROUPS && how != BY_TABLE) {
            MS_ERROR("The file %s contains multiple read groups. You must first separate the read groups using -g. See https://slow5.bioinf.science/faq for more info.",
                     path.c_str());
            return 1;
        }
        if (!o.lossy && h.aux.empty()) {
            MS_ERROR("%s has no auxiliary fields. Specify -l false to merge files with no auxiliary fields.", path.c_str());
            return 1;
        }
        const bool need_gpu = h.record_method != PRESS_NONE || h.signal_method != PRESS_NONE || o.rec_out != PRESS_NONE ||
                              o.sig_out != PRESS_NONE;
        if (need_gpu && !gpu && !(gpu = open_gpu())) return 1;
        if (how == BY_TABLE) {
            if (!have_plan) {
                if (!demux_read_table(dx, plan)) return 1;
                have_plan = true;
            }
            const int rc = demux_file(path, rd, o, plan, gpu);
            reader_close(rd);
            if (rc) return 1;
            continue;
        }

        // <out-dir>/<input name without its extension>_<index><ext>, header = one read group of the input (create_output_slow5,
        // split.c:586-653)
        const size_t slash = path.find_last_of('/');
        const std::string base = path.substr(slash == std::string::npos ? 0 : slash + 1);
        const std::string stem = base.substr(0, base.size() - ext.size());  // (the reference cuts the OUTPUT extension's length)
        std::vector<Output> outs;
        auto open_out = [&](uint32_t index, uint32_t rg) -> bool {
            Output out;
            out.path = std::string(o.arg_dir) + "/" + stem + "_" + std::to_string(index) + ext;
            out.fp = fopen(out.path.c_str(), "wb");
            if (!out.fp) {
                MS_ERROR("Output file %s could not be opened - %s.", out.path.c_str(), strerror(errno));
                return false;
            }
            setvbuf(out.fp, nullptr, _IOFBF, 1 << 20);
            OutHeader oh;
            Header &ho = oh.h;
            ho.version[0] = 0, ho.version[1] = 2, ho.version[2] = 0;
            ho.num_read_groups = 0;
            if (!o.lossy) ho.aux = h.aux;
            rg_append(oh, h, rg);
            const std::string hm = header_to_mem(ho, o.fmt_out, o.rec_out, o.sig_out);
            if (fwrite(hm.data(), 1, hm.size(), out.fp) != hm.size()) {
                MS_ERROR("Could not write the header to %s\n", out.path.c_str());
                return false;
            }
            outs.push_back(out);
            return true;
        };
        Header hdr_out = h;
        if (o.lossy) hdr_out.aux.clear();
        std::vector<uint32_t> old_rg;  // of the records of the batch being converted
        ConvertHooks hooks;
        hooks.hdr_out = &hdr_out;
        int ret = 0;
        if (how == BY_GROUPS) {
            for (uint32_t j = 0; j < h.num_read_groups; ++j)
                if (!open_out(j, j)) return 1;
            hooks.transform = [&](size_t i, Record &rec, std::vector<uint8_t> &) {
                if (rec.read_group >= outs.size()) {
                    MS_ERROR("read group %u of record %s is not in the header of %s", rec.read_group, rec.read_id.c_str(), path.c_str());
                    return false;
                }
                if (old_rg.size() <= i) old_rg.resize(i + 1);
                old_rg[i] = rec.read_group;
                rec.read_group = 0;  // split.c:88-89
                if (o.lossy) rec.aux_bytes = nullptr, rec.aux_nbytes = 0;
                return true;
            };
            hooks.route = [&](size_t i) { return outs[old_rg[i]].fp; };
            ret = convert_records(h, rd.fmt, [&](std::vector<uint8_t> &mem) {
                const int rc = reader_next_mem(rd, mem);
                if (rc < 0) MS_ERROR("Could not read file %s", path.c_str());
                return rc;
            }, nullptr, gpu, o.fmt_out, o.rec_out, o.sig_out, o.batch, o.threads, &hooks);
        } else {
            // records per output file: -r as given; -f: the count divided evenly, the first (count % files) files get one more
            // (split.c:379-400).  Counting takes one pass over the record sizes.
            long per_file = count, rem = 0;
            if (how == BY_FILES) {
                const off_t here = ftello(rd.fp);
                std::vector<uint8_t> mem;
                long total = 0;
                int rc;
                while ((rc = reader_next_mem(rd, mem)) > 0) ++total;
                if (rc < 0) {
                    MS_ERROR("Could not read file %s", path.c_str());
                    return 1;
                }
                clearerr(rd.fp);
                fseeko(rd.fp, here, SEEK_SET);
                per_file = total / count;
                rem = total % count;
            }
            hooks.transform = [&](size_t, Record &rec, std::vector<uint8_t> &) {
                rec.read_group = 0;
                if (o.lossy) rec.aux_bytes = nullptr, rec.aux_nbytes = 0;
                return true;
            };
            bool eof = false;
            uint32_t file_index = 0;
            while (!eof && ret == 0) {
                long limit = per_file;
                if (how == BY_FILES) {
                    limit = per_file + (rem > 0 ? 1 : 0);
                    --rem;
                }
                if (!open_out(file_index, 0)) return 1;
                Output &out = outs.back();
                if (limit <= 0) eof = true;  // more files asked for than there are records
                long taken = 0;
                ret = convert_records(h, rd.fmt, [&](std::vector<uint8_t> &mem) {
                    if (taken >= limit) return 0;
                    const int rc = reader_next_mem(rd, mem);
                    if (rc < 0) MS_ERROR("Could not read file %s", path.c_str());
                    if (rc == 0) eof = true;
                    if (rc > 0) ++taken;
                    return rc;
                }, out.fp, gpu, o.fmt_out, o.rec_out, o.sig_out, o.batch, o.threads, &hooks);
                if (!finish_output(out.fp, o.fmt_out, true)) ret = 1;
                out.fp = nullptr;
                if (eof && taken == 0) remove(out.path.c_str());  // the file opened after the last record (split.c:438-446)
                ++file_index;
            }
        }
        for (Output &out : outs)
            if (out.fp && !finish_output(out.fp, o.fmt_out, true)) ret = 1;
        reader_close(rd);
        if (ret) return 1;
    }
    return 0;
}
