// degrade_main.cpp -- `slow5tools-b200 degrade`: irreversible (lossy) conversion, the remaining conversion-type caller of the
// reference's batch worker (src/degrade.c).  It is `view` with one more per-sample step: between decoding a
// record and re-encoding it the b least significant bits of every sample are rounded away (slow5_rec_qts_round,
// src/degrade.c:255 -> slow5_press.c:1965-2019), which is what makes the ex-zd / svb-zd streams behind it so much smaller.
//   * blow5 -> blow5 stays on the device: the transcoder (recode_engine.cu) runs qts_round_kernel on the decoded sample slab
//     of every chunk (s5b_ctx_set_degrade) -- also when the signal method does not change;
//   * anything involving SLOW5 text goes through convert_records (view_main.cpp) with the rounding as one batch call;
//   * -b auto (the default) picks b from the header like the reference (device type, kit, experiment type and sampling
//     frequency of every read group must name one of the datasets below; src/degrade.h, src/degrade.c:58-148) and then holds
//     every record to that dataset's digitisation and sampling rate (src/degrade.c:195-211, :249-253).
// There is no CPU path for the samples: without a CUDA device the command fails.
#include <getopt.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cli_common.hpp"

using namespace s5b;

#define DG_ERROR(fmt, ...) fprintf(stderr, "[%s::ERROR]\033[1;31m " fmt "\033[0m\n", __func__, __VA_ARGS__)
#define DG_WARNING(fmt, ...) fprintf(stderr, "[%s::WARNING]\033[1;33m " fmt "\033[0m\n", __func__, __VA_ARGS__)
#define DG_INFO(fmt, ...) fprintf(stderr, "[%s::INFO]\033[1;34m " fmt "\033[0m\n", __func__, __VA_ARGS__)

namespace {

// The datasets a bit count is known for (src/degrade.h:37-92).  A chemistry comes on a family of instruments that share a
// digitisation; every (chemistry, instrument) pair is one dataset, tried in this order.
struct Instrument {
    const char *label, *device_type;
    float digitisation;
};
struct Chemistry {
    const char *label, *kit, *experiment, *frequency;
    float sampling_rate;
    int family;  // 0: MinION / GridION, 1: PromethION / PromethION 2 Solo
    int bits;
};
const Instrument INSTRUMENTS[2][2] = {
    {{"MinION", "minion", 8192.f}, {"GridION", "gridion", 8192.f}},
    {{"PromethION", "promethion", 2048.f}, {"PromethION 2 Solo", "p2_solo", 2048.f}},
};
const Chemistry CHEMISTRIES[] = {
    {"DNA lsk114 5kHz", "sqk-lsk114", "genomic_dna", "5000", 5000.f, 0, 3},
    {"DNA lsk109 4kHz", "sqk-lsk109", "genomic_dna", "4000", 4000.f, 1, 2},
    {"DNA lsk114 4kHz", "sqk-lsk114", "genomic_dna", "4000", 4000.f, 1, 3},
    {"DNA lsk114 5kHz", "sqk-lsk114", "genomic_dna", "5000", 5000.f, 1, 3},
    {"RNA rna002 3kHz", "sqk-rna002", "rna", "3000", 3000.f, 1, 2},
    {"RNA rna004 4kHz", "sqk-rna004", "rna", "4000", 4000.f, 1, 3},
    {"DNA ulk114 5kHz", "sqk-ulk114", "genomic_dna", "5000", 5000.f, 1, 3},
};

struct Dataset {
    std::string name;
    float digitisation = 0, sampling_rate = 0;
    int bits = 0;
};

const std::vector<std::string> *attr_values(const Header &h, const char *key) {
    for (const auto &kv : h.attrs)
        if (kv.first == key) return &kv.second;
    return nullptr;  // the header does not have the attribute at all
}
// every read group carries `want` under `key` (slow5_hdrcmp, src/degrade.c:154-168)
bool all_groups_have(const Header &h, const char *key, const char *want) {
    const std::vector<std::string> *v = attr_values(h, key);
    if (!v) return false;
    for (uint32_t g = 0; g < h.num_read_groups; ++g)
        if (g >= v->size() || (*v)[g] != want) return false;
    return true;
}
// sample_frequency and sample_rate name the same thing; one of them may be absent (slow5_hdrcmp_sample_freq, :175-184)
bool frequency_is(const Header &h, const char *want) {
    const bool has_freq = attr_values(h, "sample_frequency") != nullptr, has_rate = attr_values(h, "sample_rate") != nullptr;
    if (!has_freq) return all_groups_have(h, "sample_rate", want);
    return all_groups_have(h, "sample_frequency", want) && (!has_rate || all_groups_have(h, "sample_rate", want));
}
bool detect_dataset(const Header &h, Dataset &out) {
    for (const Chemistry &c : CHEMISTRIES)
        for (const Instrument &ins : INSTRUMENTS[c.family]) {
            if (all_groups_have(h, "device_type", ins.device_type) && frequency_is(h, c.frequency) &&
                all_groups_have(h, "sequencing_kit", c.kit) && all_groups_have(h, "experiment_type", c.experiment)) {
                out.name = std::string(c.label) + " " + ins.label;
                out.digitisation = ins.digitisation;
                out.sampling_rate = c.sampling_rate;
                out.bits = c.bits;
                return true;
            }
        }
    return false;
}

void usage(FILE *f) {
    fprintf(f,
            "Usage: slow5tools-b200 degrade [OPTIONS] [FILE]\n"
            "Irreversibly degrade and convert slow5/blow5 FILEs (GPU codec).\n\n"
            "OPTIONS:\n"
            "    --to FORMAT                   specify output file format (slow5 or blow5)\n"
            "    -o, --output [FILE]           output contents to FILE [stdout]\n"
            "    -c, --compress REC_MTD        record compression method [zlib] (only for blow5 format)\n"
            "    -s, --sig-compress SIG_MTD    signal compression method [ex-zd] (only for blow5 format)\n"
            "    -t, --threads INT             number of host threads for parsing/formatting [8]\n"
            "    -K, --batchsize INT           number of records loaded to the memory at once [4096]\n"
            "    --from FORMAT                 specify input file format (slow5 or blow5)\n"
            "    -b, --bits INT                specify the number of least significant bits to eliminate [auto]\n"
            "    -h, --help                    display this message and exit\n"
            "REC_MTD: none, zlib, zstd      SIG_MTD: none, svb-zd, ex-zd\n");
}

// 1..16, -1 for "auto", -2 for anything else (parse_bits, src/degrade.c:217-238)
int bits_from_arg(const char *s) {
    if (!s || !*s) {
        DG_ERROR("Invalid bits argument '%s'", s ? s : "");
        return -2;
    }
    if (!strcmp(s, "auto")) return -1;
    char *end = nullptr;
    const long v = strtol(s, &end, 10);
    if (*end) {
        DG_ERROR("Invalid bits argument '%s'", s);
        return -2;
    }
    if (v < 1 || v > 16) {
        DG_ERROR("Invalid bits argument '%ld': outside of range 1-16", v);
        return -2;
    }
    return (int)v;
}

}  // namespace

int degrade_main(int argc, char **argv) {
    static const struct option long_opts[] = {
        {"sig-compress", required_argument, nullptr, 's'}, {"compress", required_argument, nullptr, 'c'},
        {"from", required_argument, nullptr, 'f'},         {"help", no_argument, nullptr, 'h'},
        {"output", required_argument, nullptr, 'o'},       {"to", required_argument, nullptr, 'T'},
        {"threads", required_argument, nullptr, 't'},      {"batchsize", required_argument, nullptr, 'K'},
        {"bits", required_argument, nullptr, 'b'},         {nullptr, 0, nullptr, 0}};
    if (argc <= 1) {
        usage(stderr);
        return 1;
    }
    const char *arg_sig = nullptr, *arg_rec = nullptr, *arg_from = nullptr, *arg_to = nullptr, *arg_out = nullptr;
    int threads = 8, bits = -1;
    long batch = 4096;
    int opt;
    optind = 1;
    while ((opt = getopt_long(argc, argv, "s:c:f:ho:T:t:K:b:", long_opts, nullptr)) != -1) {
        switch (opt) {
            case 's': arg_sig = optarg; break;
            case 'c': arg_rec = optarg; break;
            case 'f': arg_from = optarg; break;
            case 'T': arg_to = optarg; break;
            case 'o': arg_out = optarg; break;
            case 't': threads = atoi(optarg); break;
            case 'K': batch = atol(optarg); break;
            case 'b':
                bits = bits_from_arg(optarg);
                if (bits == -2) return 1;
                if (bits > 4) DG_WARNING("%s", "bits > 4: basecalling accuracy may be adversely affected!");
                break;
            case 'h': usage(stdout); return 0;
            default: usage(stderr); return 1;
        }
    }
    if (threads < 1 || batch < 1) {
        DG_ERROR("%s", "invalid -t / -K value");
        return 1;
    }
    if (optind >= argc) {
        DG_ERROR("missing input file%s", "");
        usage(stderr);
        return 1;
    }
    if (optind != argc - 1) {
        DG_ERROR("more than 1 input file is given%s", "");
        return 1;
    }
    const char *in_path = argv[optind];
    Fmt fmt_in = FMT_UNKNOWN, fmt_out = FMT_UNKNOWN;
    if (arg_from && (fmt_in = fmt_from_name(arg_from)) == FMT_UNKNOWN) {
        DG_ERROR("invalid input format '%s'", arg_from);
        return 1;
    }
    if (arg_to && (fmt_out = fmt_from_name(arg_to)) == FMT_UNKNOWN) {
        DG_ERROR("invalid output format '%s'", arg_to);
        return 1;
    }
    if (arg_out) {
        const Fmt by_ext = fmt_from_path(arg_out);
        if (fmt_out == FMT_UNKNOWN) {
            if ((fmt_out = by_ext) == FMT_UNKNOWN) {
                DG_ERROR("cannot detect the output format from the file extension of '%s'", arg_out);
                return 1;
            }
        } else if (by_ext != FMT_UNKNOWN && by_ext != fmt_out) {
            DG_ERROR("output file extension '%s' does not match the output format '%s'", arg_out, arg_to);
            return 1;
        }
    }
    if (fmt_out == FMT_UNKNOWN) fmt_out = FMT_ASCII;  // src/degrade.c:391-393
    if (fmt_out == FMT_ASCII && (arg_rec || arg_sig)) {
        DG_ERROR("%s", "compression options (-c / -s) are only valid for blow5 output");
        return 1;
    }
    int rec_out = PRESS_ZLIB, sig_out = PRESS_EX_ZD;  // src/degrade.c:292
    if (arg_rec && (rec_out = press_from_name(arg_rec)) == PRESS_BAD) {
        DG_ERROR("invalid record compression method '%s'", arg_rec);
        return 1;
    }
    if (arg_sig && (sig_out = press_from_name(arg_sig)) == PRESS_BAD) {
        DG_ERROR("invalid signal compression method '%s'", arg_sig);
        return 1;
    }
    if (fmt_out == FMT_ASCII) rec_out = sig_out = PRESS_NONE;
    if ((rec_out != PRESS_NONE && rec_out != PRESS_ZLIB && rec_out != PRESS_ZSTD) ||
        (sig_out != PRESS_NONE && sig_out != PRESS_SVB_ZD && sig_out != PRESS_EX_ZD)) {
        DG_ERROR("%s", "this build supports record compression none/zlib/zstd and signal compression none/svb-zd/ex-zd only");
        return 1;
    }

    Reader rd;
    if (!reader_open(rd, in_path, fmt_in)) {
        DG_ERROR("File '%s' could not be opened - %s.", in_path, rd.err.c_str());
        return 1;
    }
    const Header &hdr = rd.hdr;
    if ((hdr.record_method != PRESS_NONE && hdr.record_method != PRESS_ZLIB && hdr.record_method != PRESS_ZSTD) ||
        (hdr.signal_method != PRESS_NONE && hdr.signal_method != PRESS_SVB_ZD && hdr.signal_method != PRESS_EX_ZD)) {
        DG_ERROR("%s", "input uses a compression method this build does not support (zlib/zstd as signal method)");
        reader_close(rd);
        return 1;
    }
    DG_WARNING("This tool performs lossy compression which is an irreversible operation. Just making sure it is intended. %s", "");
    Dataset ds;
    bool hold_records = false;
    if (bits == -1) {
        if (!detect_dataset(hdr, ds)) {
            DG_ERROR("No suitable bits suggestion%s", "");
            DG_ERROR("%s", "Use option -b to manually specify");
            reader_close(rd);
            return 1;
        }
        DG_INFO("Detected: %s", ds.name.c_str());
        bits = ds.bits;
        hold_records = true;
        DG_INFO("Eliminating %d bits", bits);
    }

    FILE *fout = stdout;
    if (arg_out && !(fout = fopen(arg_out, "wb"))) {
        DG_ERROR("File '%s' could not be opened - %s.", arg_out, strerror(errno));
        reader_close(rd);
        return 1;
    }
    setvbuf(fout, nullptr, _IOFBF, 1 << 20);
    s5b_ctx_t *gpu = nullptr;
    {
        const int rc = s5b_ctx_create(-1, &gpu);  // the rounding itself runs on the device: always needed
        if (rc != S5B_OK) {
            DG_ERROR("cannot initialise the GPU codec: %s", s5b_strerror(rc));
            return 1;
        }
    }
    {
        const std::string h = header_to_mem(hdr, fmt_out, rec_out, sig_out);
        if (fwrite(h.data(), 1, h.size(), fout) != h.size()) {
            DG_ERROR("%s", "could not write the header");
            return 1;
        }
    }
    int ret = 0;
    if (rd.fmt == FMT_BINARY && fmt_out == FMT_BINARY && hdr.aux.size() <= 64 && !getenv("S5B_VIEW_SLOW_PATH")) {
        std::vector<uint8_t> sz(hdr.aux.size() + 1), arr(hdr.aux.size() + 1);
        for (size_t f = 0; f < hdr.aux.size(); ++f) {
            sz[f] = hdr.aux[f].size;
            arr[f] = hdr.aux[f].is_array() ? 1 : 0;
        }
        s5b_ctx_set_aux_layout(gpu, sz.data(), arr.data(), (uint32_t)hdr.aux.size());
        s5b_ctx_set_degrade(gpu, bits, hold_records, ds.digitisation, ds.sampling_rate);
        ret = blow5_fast_convert(rd, fout, gpu, rec_out, sig_out);
        if (ret && hold_records) DG_ERROR("a record may not match %s", ds.name.c_str());
    } else {
        ConvertHooks hooks;
        hooks.qts_bits = bits;
        if (hold_records)
            hooks.transform = [&](size_t, Record &rec, std::vector<uint8_t> &) {
                // slow5_reccmp (src/degrade.c:195-211): the record's doubles against the dataset's floats
                if (rec.digitisation != (double)ds.digitisation || rec.sampling_rate != (double)ds.sampling_rate) {
                    DG_ERROR("Read with ID '%s' does not match %s", rec.read_id.c_str(), ds.name.c_str());
                    return false;
                }
                return true;
            };
        ret = convert_records(hdr, rd.fmt, [&](std::vector<uint8_t> &mem) {
            const int rc = reader_next_mem(rd, mem);
            if (rc < 0) DG_ERROR("%s", rd.err.c_str());
            return rc;
        }, fout, gpu, fmt_out, rec_out, sig_out, batch, threads, &hooks);
    }
    fflush(fout);
    if (ret == 0 && fmt_out == FMT_BINARY && write(fileno(fout), "5WOLB", 5) != 5) ret = 1;  // src/degrade.c:556-560
    if (fout != stdout && fclose(fout) != 0) ret = 1;
    reader_close(rd);
    if (ret) DG_ERROR("File conversion failed.%s", "");
    if (getenv("S5B_ORDERLY_EXIT")) s5b_ctx_destroy(gpu);
    return ret;
}
