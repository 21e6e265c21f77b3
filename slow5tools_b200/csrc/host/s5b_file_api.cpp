// s5b_file_api.cpp -- the slow5lib low-level API slice (include/slow5b200_file.h) over blow5_io + the GPU
// batch codec.  Every codec call goes through the C-ABI of include/slow5b200.h; nothing is computed here.
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/slow5b200_file.h"
#include "blow5_io.hpp"

using namespace s5b;

// The auxiliary columns of a file (name, type, element size in header order), shared by the file and the records decoded from
// it: a record may outlive its file (slow5lib's records carry their own field map), so the table is reference counted.
struct AuxTable {
    std::atomic<long> refs{1};
    std::vector<AuxField> fields;
};
static void aux_table_unref(AuxTable *t) {
    if (t && --t->refs == 0) delete t;
}
// What s5b_rec_t::aux_meta points at: where every field of this record's auxiliary section lies.  Arrays are kept as aligned
// copies (strings NUL-terminated, slow5.c:3112-3121) because slow5_aux_get_*_array hands out typed pointers.
struct AuxView {
    AuxTable *table = nullptr;
    std::vector<uint64_t> len;   // elements of field f (1 for a primitive)
    std::vector<uint64_t> at;    // primitives: offset of the value inside s5b_rec_t::aux
    std::vector<void *> arr;     // arrays: the copy, nullptr when len == 0
};
// what slow5_aux_set* has given a record a caller is building (slow5_rec_set*, slow5.c:3363-3492): field name -> element count and
// bytes; the record's binary section is laid out from it for the header's columns after every call
struct AuxPending {
    int type;
    uint64_t len;
    std::vector<uint8_t> data;
};
// what s5b_rec_t::aux_meta points at
struct RecPriv {
    AuxView *view = nullptr;                            // decoded records
    std::map<std::string, AuxPending> *pending = nullptr;  // records under construction
};
struct S5bFile;
// the public header struct first, the way back to the file behind it (s5b_hdr_get)
struct HdrPriv {
    s5b_hdr pub;
    S5bFile *owner;
};

// the public struct s5b_file (slow5b200_file.h) is the first member: a s5b_file_t* points at it and at this object
struct S5bFile {
    s5b_file pub;
    HdrPriv hdr_priv;
    s5b_hdr &hdr_pub = hdr_priv.pub;
    AuxTable *aux_table = nullptr;   // "r" files with auxiliary columns
    std::string path, mode;
    Reader rd;          // "r"
    FILE *out = nullptr;  // "w"
    bool writing = false;
    Fmt out_fmt = FMT_BINARY;   // "w": what the path's extension says (.blow5 / .slow5)
    bool hdr_written = false;
    Header hdr;         // "w": header to emit
    int rec_press = PRESS_ZLIB, sig_press = PRESS_SVB_ZD;
    s5b_ctx_t *gpu = nullptr;
    // slow5_get / slow5_decode may be called from several threads on one file (slow5lib/examples/random_read_pthreads.c): the file's
    // GPU context serves one batch at a time
    std::mutex gpu_mu;
    // read id -> (offset of the record's size prefix, bytes including the prefix): FILE.idx (slow5_idx.c:360-520)
    struct Where {
        uint64_t offset, size;
    };
    std::unordered_map<std::string, Where> index;
    bool index_loaded = false;
    // what the header / index introspection calls hand out (pointers into this object, valid while the file is open)
    std::vector<std::string> rid_order;                // read ids in index order
    std::vector<char *> rids_c, aux_names_c;
    std::vector<int> aux_types_c;
    std::vector<std::vector<std::string>> enum_labels; // per auxiliary column, empty for non-enum columns
    std::vector<std::vector<char *>> enum_labels_c;
};
static inline S5bFile *impl(s5b_file_t *f) { return reinterpret_cast<S5bFile *>(f); }
static inline const S5bFile *impl(const s5b_file_t *f) { return reinterpret_cast<const S5bFile *>(f); }

// keeps the public fields in step with the private state
static void publish(S5bFile *f) {
    f->pub.fp = f->writing ? f->out : f->rd.fp;
    const Fmt fmt = f->writing ? f->out_fmt : f->rd.fmt;
    f->pub.format = fmt == FMT_BINARY ? 2 : fmt == FMT_ASCII ? 1 : 0;
    const Header &h = f->writing ? f->hdr : f->rd.hdr;
    f->hdr_pub.version.major = h.version[0];
    f->hdr_pub.version.minor = h.version[1];
    f->hdr_pub.version.patch = h.version[2];
    f->hdr_pub.num_read_groups = h.num_read_groups;
    f->hdr_priv.owner = f;
    f->pub.header = &f->hdr_priv.pub;
    f->pub.index = nullptr;
    f->pub.meta.pathname = f->path.c_str();
    f->pub.meta.mode = f->mode.c_str();
    f->pub.meta.fd = f->pub.fp ? fileno(f->pub.fp) : -1;
    f->pub.meta.fread_buffer = nullptr;
    if (f->pub.compress) {
        s5b_press_free(f->pub.compress);
        f->pub.compress = nullptr;
    }
    if (fmt == FMT_BINARY) {  // the reference has no press object for ASCII files (slow5.c:420-431)
        const s5b_press_method_t m = {f->rec_press, f->sig_press};
        f->pub.compress = s5b_press_init(m);
    }
}

namespace {
thread_local int tl_errno = 0;
int fail(int code) {
    tl_errno = code;
    return code;
}
int ensure_gpu(S5bFile *f) {
    if (f->gpu) return S5B_OK;
    return s5b_ctx_create(-1, &f->gpu);
}
}  // namespace

extern "C" {

int s5b_errno_value(void) { return tl_errno; }

s5b_file_t *s5b_open(const char *pathname, const char *mode) {
    if (!pathname || !mode) {
        fail(S5B_ERR_ARG);
        return nullptr;
    }
    S5bFile *f = new S5bFile();
    memset(&f->pub, 0, sizeof f->pub);
    f->path = pathname;
    f->mode = mode;
    if (mode[0] == 'r') {
        if (!reader_open(f->rd, pathname, FMT_UNKNOWN)) {
            fail(S5B_ERR_IO);
            reader_close(f->rd);
            delete f;
            return nullptr;
        }
        f->rec_press = f->rd.hdr.record_method;
        f->sig_press = f->rd.hdr.signal_method;
        if (!f->rd.hdr.aux.empty()) {
            f->aux_table = new AuxTable();
            f->aux_table->fields = f->rd.hdr.aux;
        }
        publish(f);
        f->pub.meta.start_rec_offset = f->rd.fp ? (uint64_t)ftello(f->rd.fp) : 0;
        return &f->pub;
    }
    if (mode[0] == 'a' && fmt_from_path(pathname) == FMT_BINARY) {
        // append (slow5_open_with "a", slow5.c:330-417): header, columns and compression come from the file; new records go where
        // its end-of-file marker is, and s5b_close writes the marker behind them
        Reader rd;
        if (reader_open(rd, pathname, FMT_BINARY)) {
            f->hdr = rd.hdr;
            f->rec_press = rd.hdr.record_method;
            f->sig_press = rd.hdr.signal_method;
            reader_close(rd);
            f->out = fopen(pathname, "r+b");
            char tail[5];
            if (f->out && fseeko(f->out, -5, SEEK_END) == 0 && fread(tail, 1, 5, f->out) == 5 && memcmp(tail, "5WOLB", 5) == 0 &&
                fseeko(f->out, -5, SEEK_END) == 0) {
                f->writing = true;
                f->hdr_written = true;
                publish(f);
                return &f->pub;
            }
            if (f->out) fclose(f->out);
        } else {
            reader_close(rd);
        }
        fail(S5B_ERR_IO);
        delete f;
        return nullptr;
    }
    if (mode[0] == 'w' && fmt_from_path(pathname) != FMT_UNKNOWN) {
        f->out = fopen(pathname, "wb");
        if (f->out) {
            f->writing = true;
            f->out_fmt = fmt_from_path(pathname);
            if (f->out_fmt == FMT_ASCII) f->rec_press = f->sig_press = PRESS_NONE;  // text records are not compressed
            f->hdr.version[0] = 0, f->hdr.version[1] = 2, f->hdr.version[2] = 0;  // SLOW5_VERSION_STRUCT, slow5_defs.h:51-53
            publish(f);
            return &f->pub;
        }
    }
    fail(S5B_ERR_IO);
    delete f;
    return nullptr;
}

int s5b_close(s5b_file_t *fpub) {
    if (!fpub) return fail(S5B_ERR_ARG);
    S5bFile *f = impl(fpub);
    int rc = 0;
    if (f->writing) {
        if (f->out_fmt == FMT_BINARY && fwrite("5WOLB", 1, 5, f->out) != 5) rc = S5B_ERR_IO;  // slow5.c:522-531
        if (fclose(f->out) != 0) rc = S5B_ERR_IO;
    } else {
        reader_close(f->rd);
    }
    if (f->gpu) s5b_ctx_destroy(f->gpu);
    if (f->pub.compress) s5b_press_free(f->pub.compress);
    aux_table_unref(f->aux_table);
    delete f;
    return rc ? fail(rc) : 0;
}

int s5b_hdr_copy(s5b_file_t *dstp, const s5b_file_t *srcp) {
    if (!dstp || !srcp || !impl(dstp)->writing) return fail(S5B_ERR_ARG);
    impl(dstp)->hdr = impl(srcp)->rd.hdr;
    publish(impl(dstp));
    return 0;
}
int s5b_set_press(s5b_file_t *fpub, int rec_press, int sig_press) {
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!f || !f->writing || f->hdr_written) return fail(S5B_ERR_ARG);
    if (f->out_fmt != FMT_BINARY) return fail(S5B_ERR_ARG);  // "File should be in binary format (blow5)", slow5.c:585-589
    if ((rec_press != PRESS_NONE && rec_press != PRESS_ZLIB && rec_press != PRESS_ZSTD) ||
        (sig_press != PRESS_NONE && sig_press != PRESS_SVB_ZD && sig_press != PRESS_EX_ZD))
        return fail(S5B_ERR_ARG);
    f->rec_press = rec_press;
    f->sig_press = sig_press;
    publish(f);
    return 0;
}
int s5b_hdr_write(s5b_file_t *fpub) {
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!f || !f->writing) return fail(S5B_ERR_ARG);
    const std::string h = header_to_mem(f->hdr, f->out_fmt, f->rec_press, f->sig_press);
    if (fwrite(h.data(), 1, h.size(), f->out) != h.size()) return fail(S5B_ERR_IO);
    f->hdr_written = true;
    return (int)h.size();
}
int s5b_file_record_press(const s5b_file_t *f) { return f ? impl(f)->rec_press : S5B_ERR_ARG; }
int s5b_file_signal_press(const s5b_file_t *f) { return f ? impl(f)->sig_press : S5B_ERR_ARG; }

void *s5b_get_next_mem(size_t *n, s5b_file_t *fpub) {
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!f || !n || f->writing) {
        fail(S5B_ERR_ARG);
        return nullptr;
    }
    std::vector<uint8_t> mem;
    const int rc = reader_next_mem(f->rd, mem);
    if (rc <= 0) {
        fail(rc == 0 ? S5B_ERR_EOF : S5B_ERR_IO);
        *n = 0;
        return nullptr;
    }
    void *out = malloc(mem.size() + 1);  // (+1: an ASCII line is handed out NUL-terminated, slow5.c:3216-3231)
    if (!out) {
        fail(S5B_ERR_MEM);
        return nullptr;
    }
    memcpy(out, mem.data(), mem.size());
    static_cast<char *>(out)[mem.size()] = '\0';
    *n = mem.size();
    return out;
}

int s5b_get_next_bytes(char **mem, size_t *bytes, s5b_file_t *f) {
    if (!mem || !bytes) return fail(S5B_ERR_ARG);
    *mem = static_cast<char *>(s5b_get_next_mem(bytes, f));
    return *mem ? 0 : tl_errno;
}

static void aux_view_free(s5b_rec_t *r) {
    RecPriv *pr = static_cast<RecPriv *>(r->aux_meta);
    if (!pr) return;
    if (AuxView *v = pr->view) {
        for (void *p : v->arr) free(p);
        aux_table_unref(v->table);
        delete v;
    }
    delete pr->pending;
    delete pr;
    r->aux_meta = nullptr;
}
// walks the record's auxiliary section against the file's columns (slow5_rec_aux_parse, slow5.c:3088-3166); a section that does
// not fit them leaves the record without a view (the accessors then report S5B_ERR_NOAUX)
static void aux_view_build(s5b_rec_t *r, AuxTable *t) {
    if (!t) return;
    AuxView *v = new AuxView();
    const size_t nf = t->fields.size();
    v->len.assign(nf, 1);
    v->at.assign(nf, 0);
    v->arr.assign(nf, nullptr);
    uint64_t at = 0;
    bool ok = true;
    for (size_t i = 0; i < nf && ok; ++i) {
        const AuxField &fd = t->fields[i];
        uint64_t cnt = 1;
        if (fd.is_array()) {
            if (at + 8 > r->aux_len) {
                ok = false;
                break;
            }
            memcpy(&cnt, r->aux + at, 8);
            at += 8;
        }
        if (cnt > r->aux_len || at + cnt * fd.size > r->aux_len) {
            ok = false;
            break;
        }
        v->len[i] = cnt;
        v->at[i] = at;
        if (fd.is_array() && cnt) {
            const uint64_t bytes = cnt * fd.size;
            uint8_t *copy = static_cast<uint8_t *>(malloc(bytes + 1));
            if (!copy) {
                ok = false;
                break;
            }
            memcpy(copy, r->aux + at, bytes);
            copy[bytes] = 0;
            v->arr[i] = copy;
        }
        at += cnt * fd.size;
    }
    if (!ok || at != r->aux_len) {
        for (void *p : v->arr) free(p);
        delete v;
        return;
    }
    ++t->refs;
    v->table = t;
    RecPriv *pr = new RecPriv();
    pr->view = v;
    r->aux_meta = pr;
}

void s5b_rec_free(s5b_rec_t *r) {
    if (!r) return;
    aux_view_free(r);
    free(r->read_id);
    free(r->raw_signal);
    free(r->aux);
    free(r);
}

// fills (or allocates) a caller-visible record from a parsed one; `sig` is handed over
static void fill_rec(s5b_rec_t **slot, const Record &rec, void *sig, size_t sig_bytes, AuxTable *table) {
    s5b_rec_t *r = *slot;
    if (!r) {
        r = static_cast<s5b_rec_t *>(calloc(1, sizeof *r));
        *slot = r;
    } else {  // reuse the struct, rebuild its members (slow5.c:2626-2639)
        aux_view_free(r);
        free(r->read_id);
        free(r->raw_signal);
        free(r->aux);
    }
    r->read_id_len = (uint16_t)rec.read_id.size();
    r->read_id = strndup(rec.read_id.data(), rec.read_id.size());
    r->read_group = rec.read_group;
    r->digitisation = rec.digitisation;
    r->offset = rec.offset;
    r->range = rec.range;
    r->sampling_rate = rec.sampling_rate;
    r->len_raw_signal = sig_bytes / 2;
    r->raw_signal = static_cast<int16_t *>(sig);
    r->aux_len = rec.aux_nbytes;
    r->aux = static_cast<uint8_t *>(malloc(r->aux_len ? r->aux_len : 1));
    if (r->aux_len) memcpy(r->aux, rec.aux_bytes, r->aux_len);
    aux_view_build(r, table);
}

int s5b_decode_batch(s5b_file_t *fpub, char **mems, size_t *bytes, size_t n, s5b_rec_t **reads) {
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!f || !mems || !bytes || !reads) return fail(S5B_ERR_ARG);
    if (n == 0) return 0;
    const Header &h = f->rd.hdr;
    if (f->rd.fmt == FMT_ASCII) {  // SLOW5 text lines: nothing is compressed, the parse is host work (slow5.c:2641-2810)
        for (size_t i = 0; i < n; ++i) {
            Record rec;
            std::vector<uint8_t> aux_store;
            std::string err;
            if (!record_parse_ascii(mems[i], bytes[i], h, rec, aux_store, err)) return fail(S5B_ERR_RECPARSE);
            const size_t nb = rec.raw_signal.size() * 2;
            void *sig = malloc(nb ? nb : 1);
            if (!sig) return fail(S5B_ERR_MEM);
            memcpy(sig, rec.raw_signal.data(), nb);
            fill_rec(&reads[i], rec, sig, nb, f->aux_table);
        }
        return 0;
    }
    std::unique_lock<std::mutex> gpu_lock(f->gpu_mu, std::defer_lock);
    if (h.record_method != PRESS_NONE || h.signal_method != PRESS_NONE) {
        gpu_lock.lock();
        const int rc = ensure_gpu(f);
        if (rc != S5B_OK) return fail(rc);
    }
    std::vector<const void *> ptrs(n);
    std::vector<size_t> counts(n);
    if (h.record_method == PRESS_ZLIB || h.record_method == PRESS_ZSTD) {
        std::vector<void *> out(n, nullptr);
        std::vector<size_t> out_n(n, 0);
        for (size_t i = 0; i < n; ++i) {
            ptrs[i] = mems[i];
            counts[i] = bytes[i];
        }
        const int rc = s5b_depress_batch_host(f->gpu, h.record_method == PRESS_ZLIB ? S5B_COMPRESS_ZLIB : S5B_COMPRESS_ZSTD,
                                              ptrs.data(), counts.data(), n, out.data(), out_n.data());
        if (rc != S5B_OK) {
            for (void *p : out) free(p);
            return fail(rc == S5B_ERR_PRESS ? S5B_ERR_PRESS : rc);
        }
        for (size_t i = 0; i < n; ++i) {  // the decompressed record replaces *mem (slow5.c:2595-2597)
            free(mems[i]);
            mems[i] = static_cast<char *>(out[i]);
            bytes[i] = out_n[i];
        }
    } else if (h.record_method != PRESS_NONE) {
        return fail(S5B_ERR_ARG);
    }
    std::vector<Record> rec(n);
    std::string err;
    for (size_t i = 0; i < n; ++i)
        if (!record_parse_binary(reinterpret_cast<const uint8_t *>(mems[i]), bytes[i], h, h.signal_method, rec[i], err))
            return fail(S5B_ERR_RECPARSE);
    std::vector<void *> sig(n, nullptr);
    std::vector<size_t> sig_n(n, 0);
    if (h.signal_method == PRESS_SVB_ZD || h.signal_method == PRESS_EX_ZD) {  // PRESS_* == S5B_COMPRESS_*
        for (size_t i = 0; i < n; ++i) {
            ptrs[i] = rec[i].sig_bytes;
            counts[i] = rec[i].sig_nbytes;
        }
        const int rc = s5b_depress_batch_host(f->gpu, h.signal_method, ptrs.data(), counts.data(), n, sig.data(), sig_n.data());
        if (rc != S5B_OK) {
            for (void *p : sig) free(p);
            return fail(rc);
        }
    } else if (h.signal_method == PRESS_NONE) {
        for (size_t i = 0; i < n; ++i) {
            sig_n[i] = rec[i].sig_nbytes;
            sig[i] = malloc(sig_n[i] ? sig_n[i] : 1);
            if (sig[i]) memcpy(sig[i], rec[i].sig_bytes, sig_n[i]);
        }
    } else {
        return fail(S5B_ERR_ARG);
    }
    for (size_t i = 0; i < n; ++i) fill_rec(&reads[i], rec[i], sig[i], sig_n[i], f->aux_table);
    return 0;
}

int s5b_decode(char **mem, size_t *bytes, s5b_rec_t **read, s5b_file_t *f) {
    if (!mem || !*mem || !bytes || !read) return fail(S5B_ERR_ARG);
    return s5b_decode_batch(f, mem, bytes, 1, read);
}

int s5b_encode_batch(s5b_file_t *fpub, s5b_rec_t **reads, size_t n, char **mems, size_t *bytes) {
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!f || !reads || !mems || !bytes) return fail(S5B_ERR_ARG);
    if (n == 0) return 0;
    if (f->writing && f->out_fmt == FMT_ASCII) {  // SLOW5 text lines (slow5.c:3837-3926): host formatting, nothing to compress
        for (size_t i = 0; i < n; ++i) {
            const s5b_rec_t *r = reads[i];
            Record rec;
            rec.read_id.assign(r->read_id ? r->read_id : "", r->read_id ? r->read_id_len : 0);
            rec.read_group = r->read_group;
            rec.digitisation = r->digitisation;
            rec.offset = r->offset;
            rec.range = r->range;
            rec.sampling_rate = r->sampling_rate;
            rec.len_raw_signal = r->len_raw_signal;
            rec.raw_signal.assign(r->raw_signal, r->raw_signal + r->len_raw_signal);
            rec.aux_bytes = r->aux;
            rec.aux_nbytes = r->aux_len;
            std::string line;
            record_to_ascii(rec, f->hdr, line);
            char *m = static_cast<char *>(malloc(line.size() + 1));
            if (!m) return fail(S5B_ERR_MEM);
            memcpy(m, line.data(), line.size());
            m[line.size()] = '\0';
            mems[i] = m;
            bytes[i] = line.size();
        }
        return 0;
    }
    std::unique_lock<std::mutex> gpu_lock(f->gpu_mu, std::defer_lock);
    if (f->rec_press != PRESS_NONE || f->sig_press != PRESS_NONE) {
        gpu_lock.lock();
        const int rc = ensure_gpu(f);
        if (rc != S5B_OK) return fail(rc);
    }
    std::vector<const void *> ptrs(n);
    std::vector<size_t> counts(n);
    std::vector<void *> svb(n, nullptr);
    std::vector<size_t> svb_n(n, 0);
    if (f->sig_press != PRESS_NONE) {
        for (size_t i = 0; i < n; ++i) {
            ptrs[i] = reads[i]->raw_signal;
            counts[i] = reads[i]->len_raw_signal * 2;
        }
        const int rc = s5b_compress_batch_host(f->gpu, f->sig_press, ptrs.data(), counts.data(), n, svb.data(), svb_n.data());
        if (rc != S5B_OK) {
            for (void *p : svb) free(p);
            return fail(rc);
        }
    }
    std::vector<std::vector<uint8_t>> packed(n);
    std::vector<uint32_t> splits(n, 0);
    for (size_t i = 0; i < n; ++i) {
        s5b_rec_t *r = reads[i];
        Record rec;
        rec.read_id.assign(r->read_id ? r->read_id : "", r->read_id ? r->read_id_len : 0);
        rec.read_group = r->read_group;
        rec.digitisation = r->digitisation;
        rec.offset = r->offset;
        rec.range = r->range;
        rec.sampling_rate = r->sampling_rate;
        rec.aux_bytes = r->aux;
        rec.aux_nbytes = r->aux_len;
        uint64_t at = 0;
        if (f->sig_press != PRESS_NONE) {
            record_to_binary(rec, static_cast<const uint8_t *>(svb[i]), svb_n[i], true, packed[i], &at);
            if (f->sig_press == PRESS_SVB_ZD) splits[i] = (uint32_t)(at + 4 + (r->len_raw_signal + 3) / 4);
            // destructive like the reference: the record now owns the compressed signal (slow5.c:3982-3984)
            free(r->raw_signal);
            r->raw_signal = static_cast<int16_t *>(svb[i]);
            r->len_raw_signal = svb_n[i];
        } else {
            record_to_binary(rec, reinterpret_cast<const uint8_t *>(r->raw_signal), r->len_raw_signal * 2, false, packed[i], &at);
        }
    }
    std::vector<void *> z(n, nullptr);
    std::vector<size_t> z_n(n, 0);
    const bool rec_packed = f->rec_press == PRESS_ZLIB || f->rec_press == PRESS_ZSTD;
    if (rec_packed) {
        for (size_t i = 0; i < n; ++i) {
            ptrs[i] = packed[i].data();
            counts[i] = packed[i].size();
        }
        const int rc = s5b_compress_records_host(f->gpu, f->rec_press == PRESS_ZSTD ? S5B_COMPRESS_ZSTD : S5B_COMPRESS_ZLIB,
                                                 ptrs.data(), counts.data(), splits.data(), n, z.data(), z_n.data());
        if (rc != S5B_OK) {
            for (void *p : z) free(p);
            return fail(rc);
        }
    }
    for (size_t i = 0; i < n; ++i) {
        const void *p = rec_packed ? z[i] : packed[i].data();
        const uint64_t sz = rec_packed ? z_n[i] : packed[i].size();
        char *m = static_cast<char *>(malloc(8 + sz));
        if (!m) return fail(S5B_ERR_MEM);
        memcpy(m, &sz, 8);  // size prefix, slow5.c:4055-4060
        memcpy(m + 8, p, sz);
        mems[i] = m;
        bytes[i] = 8 + sz;
        free(z[i]);
    }
    return 0;
}

int s5b_encode(char **mem, size_t *bytes, s5b_rec_t *read, s5b_file_t *f) {
    if (!mem || !bytes || !read) return fail(S5B_ERR_ARG);
    return s5b_encode_batch(f, &read, 1, mem, bytes) == 0 ? 0 : -1;
}

int s5b_write_bytes(char *mem, size_t bytes, s5b_file_t *fpub) {
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!f || !f->writing || !mem) return fail(S5B_ERR_ARG);
    if (fwrite(mem, 1, bytes, f->out) != bytes) return fail(S5B_ERR_IO);
    return 0;  // slow5.c:3785-3794
}

// The single-record conveniences of slow5.h over the calls above (a batch of one each: correct, latency-bound; callers that care
// about throughput use the batch forms).
int s5b_get_next(s5b_rec_t **read, s5b_file_t *f) {  // slow5_get_next, slow5.c:3339-3361
    if (!read || !f) return fail(S5B_ERR_ARG);
    char *mem = nullptr;
    size_t bytes = 0;
    const int rc = s5b_get_next_bytes(&mem, &bytes, f);
    if (rc < 0) return rc;
    const int rd = s5b_decode(&mem, &bytes, read, f);
    free(mem);
    return rd;
}

int s5b_write(s5b_rec_t *read, s5b_file_t *f) {  // slow5_write -> slow5_rec_fwrite, slow5.c:3765-3799: bytes written or -1
    if (!read || !f) return -1;
    char *mem = nullptr;
    size_t bytes = 0;
    if (s5b_encode(&mem, &bytes, read, f) != 0) return -1;
    const int rc = s5b_write_bytes(mem, bytes, f);
    free(mem);
    return rc < 0 ? -1 : (int)bytes;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// slow5_mt.h twins (slow5lib/src/slow5_mt.c:202-400)
// ---------------------------------------------------------------------------------------------
extern "C" {

s5b_mt_t *s5b_init_mt(int num_thread, s5b_file_t *fp) {  // slow5_mt.c:257-268
    if (!fp) return nullptr;
    s5b_mt_t *mt = static_cast<s5b_mt_t *>(calloc(1, sizeof(s5b_mt_t)));
    if (!mt) return nullptr;
    mt->sf = fp;
    mt->num_thread = num_thread;
    return mt;
}
void s5b_free_mt(s5b_mt_t *mt) { free(mt); }

s5b_batch_t *s5b_init_batch(int cap) {  // slow5_mt.c:270-291
    if (cap <= 0) return nullptr;
    s5b_batch_t *b = static_cast<s5b_batch_t *>(calloc(1, sizeof(s5b_batch_t)));
    if (!b) return nullptr;
    b->capacity_rec = cap;
    b->mem_records = static_cast<char **>(calloc(cap, sizeof(char *)));
    b->mem_bytes = static_cast<size_t *>(calloc(cap, sizeof(size_t)));
    b->slow5_rec = static_cast<s5b_rec_t **>(calloc(cap, sizeof(s5b_rec_t *)));
    if (!b->mem_records || !b->mem_bytes || !b->slow5_rec) {
        s5b_free_batch(b);
        return nullptr;
    }
    return b;
}
static void batch_drop_mem(s5b_batch_t *b) {
    for (int i = 0; i < b->capacity_rec; ++i) {
        free(b->mem_records[i]);
        b->mem_records[i] = nullptr;
        b->mem_bytes[i] = 0;
    }
}
void s5b_free_batch(s5b_batch_t *b) {  // slow5_mt.c:293-316
    if (!b) return;
    if (b->mem_records && b->mem_bytes) batch_drop_mem(b);
    if (b->slow5_rec)
        for (int i = 0; i < b->capacity_rec; ++i) s5b_rec_free(b->slow5_rec[i]);
    free(b->mem_records);
    free(b->mem_bytes);
    free(b->slow5_rec);
    free(b);
}

int s5b_get_next_batch(s5b_mt_t *mt, s5b_batch_t *b, int num_reads) {  // slow5_mt.c:336-351 (+ :84-107)
    if (!mt || !mt->sf || !b || num_reads < 0 || num_reads > b->capacity_rec) return fail(S5B_ERR_ARG);
    batch_drop_mem(b);
    int n = 0;
    while (n < num_reads) {
        if (s5b_get_next_bytes(&b->mem_records[n], &b->mem_bytes[n], mt->sf) < 0) {
            if (s5b_errno_value() != S5B_ERR_EOF) return s5b_errno_value();
            break;
        }
        ++n;
    }
    b->n_rec = n;
    if (n == 0) return 0;
    const int rc = s5b_decode_batch(mt->sf, b->mem_records, b->mem_bytes, (size_t)n, b->slow5_rec);
    if (rc < 0) return rc;
    return n;
}

int s5b_encode_batch_mt(s5b_mt_t *mt, s5b_batch_t *b, int num_reads) {  // slow5_mt.c:353-357
    if (!mt || !mt->sf || !b || num_reads < 0 || num_reads > b->capacity_rec) return fail(S5B_ERR_ARG);
    batch_drop_mem(b);
    b->n_rec = num_reads;
    if (num_reads == 0) return 0;
    const int rc = s5b_encode_batch(mt->sf, b->slow5_rec, (size_t)num_reads, b->mem_records, b->mem_bytes);
    return rc < 0 ? rc : num_reads;
}

int s5b_idx_load(s5b_file_t *fpub) {  // slow5_idx_load (slow5.h:560): reads FILE.idx, written by `slow5tools-b200 index`
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!f || f->writing) return fail(S5B_ERR_ARG);
    if (f->index_loaded) return 0;
    FILE *x = fopen((f->path + ".idx").c_str(), "rb");
    if (!x) {
        // "Creates the index if not found" (slow5.c:4152-4169, slow5_idx.c:60-110): the same builder as `slow5tools-b200 index`
        fprintf(stderr, "[slow5_idx_init::INFO]\033[1;34m Index file not found. Creating an index at '%s.idx'.\033[0m\n", f->path.c_str());
        if (index_build_file(f->path.c_str()) != 0) return fail(S5B_ERR_IO);
        x = fopen((f->path + ".idx").c_str(), "rb");
    }
    if (!x) return fail(S5B_ERR_IO);
    std::vector<uint8_t> b;
    uint8_t tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof tmp, x)) > 0) b.insert(b.end(), tmp, tmp + got);
    fclose(x);
    if (b.size() < 64 + 8 || memcmp(b.data(), "SLOW5IDX\1", 9) != 0 || memcmp(b.data() + b.size() - 8, "XDI5WOLS", 8) != 0)
        return fail(S5B_ERR_IO);
    size_t pos = 64;
    const size_t end = b.size() - 8;
    while (pos + 2 <= end) {
        uint16_t n;
        memcpy(&n, b.data() + pos, 2);
        if (pos + 2 + n + 16 > end) return fail(S5B_ERR_IO);
        S5bFile::Where w;
        memcpy(&w.offset, b.data() + pos + 2 + n, 8);
        memcpy(&w.size, b.data() + pos + 2 + n + 8, 8);
        f->index.emplace(std::string(reinterpret_cast<const char *>(b.data() + pos + 2), n), w);
        f->rid_order.emplace_back(reinterpret_cast<const char *>(b.data() + pos + 2), n);
        pos += 2 + (size_t)n + 16;
    }
    f->index_loaded = true;
    f->pub.index = &f->index;
    return 0;
}

void s5b_idx_unload(s5b_file_t *fpub) {  // slow5_idx_unload, slow5.c:4191-4195
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!f) return;
    f->index.clear();
    f->rid_order.clear();
    f->rids_c.clear();
    f->index_loaded = false;
    f->pub.index = nullptr;
}

// slow5_get_batch (slow5_mt.c:319-333): the records of `num_rid` read ids, fetched with pread() and decoded as one GPU batch
int s5b_get_batch(s5b_mt_t *mt, s5b_batch_t *b, char **rid, int num_rid) {
    if (!mt || !mt->sf || !b || !rid || num_rid < 0 || num_rid > b->capacity_rec) return fail(S5B_ERR_ARG);
    S5bFile *f = impl(mt->sf);
    if (f->writing || f->rd.fmt != FMT_BINARY) return fail(S5B_ERR_ARG);
    if (!f->index_loaded) {
        const int rc = s5b_idx_load(mt->sf);
        if (rc < 0) return rc;
    }
    batch_drop_mem(b);
    b->rid = rid;
    b->n_rec = num_rid;
    const int fd = fileno(f->rd.fp);
    for (int i = 0; i < num_rid; ++i) {
        const auto it = rid[i] ? f->index.find(rid[i]) : f->index.end();
        if (it == f->index.end() || it->second.size < 8) return fail(S5B_ERR_ARG);  // SLOW5_ERR_NOTFOUND in the reference
        const size_t bytes = (size_t)(it->second.size - 8);
        char *m = static_cast<char *>(malloc(bytes ? bytes : 1));
        if (!m) return fail(S5B_ERR_MEM);
        size_t done = 0;
        while (done < bytes) {
            const ssize_t r = pread(fd, m + done, bytes - done, (off_t)(it->second.offset + 8 + done));
            if (r <= 0) {
                free(m);
                return fail(S5B_ERR_IO);
            }
            done += (size_t)r;
        }
        b->mem_records[i] = m;
        b->mem_bytes[i] = bytes;
    }
    if (num_rid == 0) return 0;
    const int rc = s5b_decode_batch(mt->sf, b->mem_records, b->mem_bytes, (size_t)num_rid, b->slow5_rec);
    return rc < 0 ? rc : num_rid;
}

int s5b_get(const char *read_id, s5b_rec_t **read, s5b_file_t *fpub) {  // slow5_get, slow5.c:2517-2578
    S5bFile *f = fpub ? impl(fpub) : nullptr;
    if (!read_id || !read || !f || f->writing) return fail(S5B_ERR_ARG);
    if (!f->index_loaded) return fail(S5B_ERR_NOIDX);
    const auto it = f->index.find(read_id);
    // an index entry covers the record's size prefix (BLOW5) or its line with the newline (SLOW5), slow5_idx.c:207-334
    const bool text = f->rd.fmt == FMT_ASCII;
    const uint64_t skip = text ? 0 : 8, drop = text ? 1 : 8;
    if (it == f->index.end() || it->second.size < drop) return fail(S5B_ERR_NOTFOUND);
    size_t bytes = (size_t)(it->second.size - drop);
    char *m = static_cast<char *>(malloc(bytes + 1));
    if (!m) return fail(S5B_ERR_MEM);
    m[bytes] = '\0';
    const int fd = fileno(f->rd.fp);
    size_t done = 0;
    while (done < bytes) {
        const ssize_t r = pread(fd, m + done, bytes - done, (off_t)(it->second.offset + skip + done));
        if (r <= 0) {
            free(m);
            return fail(S5B_ERR_IO);
        }
        done += (size_t)r;
    }
    const int rc = s5b_decode(&m, &bytes, read, fpub);
    free(m);
    return rc;
}

// the "even lazier" forms pyslow5 uses (slow5_mt.h:59-65, slow5_mt.c:390-447): the mt / batch objects live for one call, the
// records are handed over as a bare array
int s5b_get_next_batch_lazy(s5b_rec_t ***read, s5b_file_t *fp, int batch_size, int num_threads) {
    if (!read) return fail(S5B_ERR_ARG);
    s5b_mt_t *mt = s5b_init_mt(num_threads, fp);
    s5b_batch_t *b = s5b_init_batch(batch_size);
    if (!mt || !b) {
        s5b_free_batch(b);
        s5b_free_mt(mt);
        return fail(S5B_ERR_MEM);
    }
    const int ret = s5b_get_next_batch(mt, b, batch_size);
    *read = b->slow5_rec;
    b->slow5_rec = nullptr;
    s5b_free_batch(b);
    s5b_free_mt(mt);
    return ret;
}
int s5b_get_batch_lazy(s5b_rec_t ***read, s5b_file_t *fp, char **rid, int num_rid, int num_threads) {
    if (!read) return fail(S5B_ERR_ARG);
    s5b_mt_t *mt = s5b_init_mt(num_threads, fp);
    s5b_batch_t *b = s5b_init_batch(num_rid);
    if (!mt || !b) {
        s5b_free_batch(b);
        s5b_free_mt(mt);
        return fail(S5B_ERR_MEM);
    }
    const int ret = s5b_get_batch(mt, b, rid, num_rid);
    *read = b->slow5_rec;
    b->slow5_rec = nullptr;
    s5b_free_batch(b);
    s5b_free_mt(mt);
    return ret;
}
int s5b_write_batch_lazy(s5b_rec_t **read, s5b_file_t *fp, int batch_size, int num_threads) {
    if (!read) return fail(S5B_ERR_ARG);
    s5b_mt_t *mt = s5b_init_mt(num_threads, fp);
    s5b_batch_t *b = s5b_init_batch(batch_size);
    if (!mt || !b) {
        s5b_free_batch(b);
        s5b_free_mt(mt);
        return fail(S5B_ERR_MEM);
    }
    free(b->slow5_rec);
    b->slow5_rec = read;          // the caller's records: written, not taken over
    const int ret = s5b_write_batch(mt, b, batch_size);
    b->slow5_rec = nullptr;
    s5b_free_batch(b);
    s5b_free_mt(mt);
    return ret;
}
void s5b_free_batch_lazy(s5b_rec_t ***read, int num_rec) {
    if (!read || !*read) return;
    for (int i = 0; i < num_rec; ++i) s5b_rec_free((*read)[i]);
    free(*read);
    *read = nullptr;
}

int s5b_write_batch(s5b_mt_t *mt, s5b_batch_t *b, int num_reads) {  // slow5_mt.c:359-378
    const int rc = s5b_encode_batch_mt(mt, b, num_reads);
    if (rc < 0) return rc;
    for (int i = 0; i < num_reads; ++i)
        if (s5b_write_bytes(b->mem_records[i], b->mem_bytes[i], mt->sf) < 0) return s5b_errno_value();
    return num_reads;
}


}  // extern "C"

// ---- auxiliary field accessors and header attributes (slow5.h:396, :469-508; slow5.c:1383-1400, :3493-3660) -------------
namespace {
// the field's position in the record's view, or a negative S5B_ERR_* (slow5.c:3496-3522)
int aux_find(const s5b_rec_t *read, const char *field, int want_type, const AuxView **view) {
    if (!read || !field) return S5B_ERR_ARG;
    const RecPriv *pr = static_cast<const RecPriv *>(read->aux_meta);
    const AuxView *v = pr ? pr->view : nullptr;
    if (!v) return S5B_ERR_NOAUX;
    const std::vector<AuxField> &fs = v->table->fields;
    for (size_t i = 0; i < fs.size(); ++i)
        if (fs[i].name == field) {
            if (fs[i].type != want_type) return S5B_ERR_TYPE;
            *view = v;
            return (int)i;
        }
    return S5B_ERR_NOFLD;
}
template <typename T>
T aux_prim(const s5b_rec_t *read, const char *field, int *err, int type, T null_value) {
    const AuxView *v = nullptr;
    const int i = aux_find(read, field, type, &v);
    T val = null_value;
    if (i >= 0) memcpy(&val, read->aux + v->at[i], sizeof(T));
    else fail(i);
    if (err) *err = i >= 0 ? 0 : i;
    return val;
}
template <typename T>
T *aux_array(const s5b_rec_t *read, const char *field, uint64_t *len, int *err, int type) {
    const AuxView *v = nullptr;
    const int i = aux_find(read, field, type, &v);
    T *val = nullptr;
    if (i >= 0) {
        val = static_cast<T *>(v->arr[i]);   // NULL with *len = 0 for a value marked missing: not an error
        if (len) *len = v->len[i];
    } else {
        fail(i);
    }
    if (err) *err = i >= 0 ? 0 : i;
    return val;
}
}  // namespace

extern "C" {

int8_t s5b_aux_get_int8(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<int8_t>(r, f, err, AUX_INT8, INT8_MAX); }
int16_t s5b_aux_get_int16(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<int16_t>(r, f, err, AUX_INT16, INT16_MAX); }
int32_t s5b_aux_get_int32(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<int32_t>(r, f, err, AUX_INT32, INT32_MAX); }
int64_t s5b_aux_get_int64(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<int64_t>(r, f, err, AUX_INT64, INT64_MAX); }
uint8_t s5b_aux_get_uint8(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<uint8_t>(r, f, err, AUX_UINT8, UINT8_MAX); }
uint16_t s5b_aux_get_uint16(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<uint16_t>(r, f, err, AUX_UINT16, UINT16_MAX); }
uint32_t s5b_aux_get_uint32(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<uint32_t>(r, f, err, AUX_UINT32, UINT32_MAX); }
uint64_t s5b_aux_get_uint64(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<uint64_t>(r, f, err, AUX_UINT64, UINT64_MAX); }
float s5b_aux_get_float(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<float>(r, f, err, AUX_FLOAT, nanf("")); }
double s5b_aux_get_double(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<double>(r, f, err, AUX_DOUBLE, nan("")); }
char s5b_aux_get_char(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<char>(r, f, err, AUX_CHAR, 0); }
uint8_t s5b_aux_get_enum(const s5b_rec_t *r, const char *f, int *err) { return aux_prim<uint8_t>(r, f, err, AUX_ENUM, UINT8_MAX); }

int8_t *s5b_aux_get_int8_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<int8_t>(r, f, len, err, AUX_INT8_ARRAY); }
int16_t *s5b_aux_get_int16_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<int16_t>(r, f, len, err, AUX_INT16_ARRAY); }
int32_t *s5b_aux_get_int32_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<int32_t>(r, f, len, err, AUX_INT32_ARRAY); }
int64_t *s5b_aux_get_int64_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<int64_t>(r, f, len, err, AUX_INT64_ARRAY); }
uint8_t *s5b_aux_get_uint8_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<uint8_t>(r, f, len, err, AUX_UINT8_ARRAY); }
uint16_t *s5b_aux_get_uint16_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<uint16_t>(r, f, len, err, AUX_UINT16_ARRAY); }
uint32_t *s5b_aux_get_uint32_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<uint32_t>(r, f, len, err, AUX_UINT32_ARRAY); }
uint64_t *s5b_aux_get_uint64_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<uint64_t>(r, f, len, err, AUX_UINT64_ARRAY); }
float *s5b_aux_get_float_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<float>(r, f, len, err, AUX_FLOAT_ARRAY); }
double *s5b_aux_get_double_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<double>(r, f, len, err, AUX_DOUBLE_ARRAY); }
char *s5b_aux_get_string(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<char>(r, f, len, err, AUX_STRING); }
uint8_t *s5b_aux_get_enum_array(const s5b_rec_t *r, const char *f, uint64_t *len, int *err) { return aux_array<uint8_t>(r, f, len, err, AUX_ENUM_ARRAY); }

// ---- header / index introspection (slow5.h:633-654; slow5.c:907-928, :1402-1490, :2549-2569) ----------------------------
static S5bFile *owner_of(const s5b_hdr_t *header) {
    return header ? reinterpret_cast<const HdrPriv *>(header)->owner : nullptr;
}

const char **s5b_get_hdr_keys(const s5b_hdr_t *header, uint64_t *len) {   // malloc()'d array (the caller frees it), keys sorted
    S5bFile *f = owner_of(header);
    if (len) *len = 0;
    if (!f) return nullptr;
    const Header &h = f->writing ? f->hdr : f->rd.hdr;
    if (len) *len = h.attrs.size();
    if (h.attrs.empty()) return nullptr;
    const char **keys = static_cast<const char **>(malloc(h.attrs.size() * sizeof *keys));
    if (!keys) {
        fail(S5B_ERR_MEM);
        return nullptr;
    }
    for (size_t i = 0; i < h.attrs.size(); ++i) keys[i] = h.attrs[i].first.c_str();
    std::stable_sort(keys, keys + h.attrs.size(), [](const char *a, const char *b) { return strcmp(a, b) < 0; });
    return keys;
}

char **s5b_get_aux_names(const s5b_hdr_t *header, uint64_t *len) {   // the library's own array: not to be freed
    S5bFile *f = owner_of(header);
    if (len) *len = 0;
    if (!f) return nullptr;
    const Header &h = f->writing ? f->hdr : f->rd.hdr;
    if (len) *len = h.aux.size();
    if (h.aux.empty()) return nullptr;
    f->aux_names_c.clear();
    for (const AuxField &a : h.aux) f->aux_names_c.push_back(const_cast<char *>(a.name.c_str()));
    return f->aux_names_c.data();
}

int *s5b_get_aux_types(const s5b_hdr_t *header, uint64_t *len) {   // enum slow5_aux_type values (slow5.h:104-131)
    S5bFile *f = owner_of(header);
    if (len) *len = 0;
    if (!f) return nullptr;
    const Header &h = f->writing ? f->hdr : f->rd.hdr;
    if (len) *len = h.aux.size();
    if (h.aux.empty()) return nullptr;
    f->aux_types_c.clear();
    for (const AuxField &a : h.aux) f->aux_types_c.push_back(a.type);
    return f->aux_types_c.data();
}

char **s5b_get_aux_enum_labels(const s5b_hdr_t *header, const char *field, uint8_t *n) {
    S5bFile *f = owner_of(header);
    if (!f || !field) {
        fail(S5B_ERR_ARG);
        return nullptr;
    }
    const Header &h = f->writing ? f->hdr : f->rd.hdr;
    if (h.aux.empty()) {
        fail(S5B_ERR_NOAUX);
        return nullptr;
    }
    bool any_enum = false;
    for (const AuxField &a : h.aux) any_enum |= a.type == AUX_ENUM || a.type == AUX_ENUM_ARRAY;
    if (!any_enum) {
        fail(S5B_ERR_TYPE);
        return nullptr;
    }
    for (size_t i = 0; i < h.aux.size(); ++i) {
        if (h.aux[i].name != field) continue;
        if (h.aux[i].type != AUX_ENUM && h.aux[i].type != AUX_ENUM_ARRAY) {
            fail(S5B_ERR_TYPE);
            return nullptr;
        }
        if (f->enum_labels.size() != h.aux.size()) {
            f->enum_labels.assign(h.aux.size(), {});
            f->enum_labels_c.assign(h.aux.size(), {});
        }
        if (f->enum_labels[i].empty()) {  // "enum{a,b,c}" / "enum*{a,b,c}" as written in the header (slow5.c:1159-1258)
            const std::string &t = h.aux[i].type_str;
            const size_t open = t.find('{'), close = t.rfind('}');
            if (open != std::string::npos && close != std::string::npos && close > open) {
                size_t at = open + 1;
                while (at <= close) {
                    const size_t end = std::min(t.find(',', at), close);
                    f->enum_labels[i].push_back(t.substr(at, end - at));
                    at = end + 1;
                }
            }
            for (std::string &l : f->enum_labels[i]) f->enum_labels_c[i].push_back(const_cast<char *>(l.c_str()));
        }
        if (n) *n = (uint8_t)f->enum_labels_c[i].size();
        return f->enum_labels_c[i].data();
    }
    fail(S5B_ERR_NOFLD);
    return nullptr;
}

char **s5b_get_rids(const s5b_file_t *fpub, uint64_t *len) {   // read ids in index order; the library's own array
    S5bFile *f = fpub ? const_cast<S5bFile *>(impl(fpub)) : nullptr;
    if (len) *len = 0;
    if (!f || !f->index_loaded) {
        fail(f ? S5B_ERR_NOIDX : S5B_ERR_ARG);
        return nullptr;
    }
    if (f->rids_c.size() != f->rid_order.size()) {
        f->rids_c.clear();
        for (std::string &r : f->rid_order) f->rids_c.push_back(const_cast<char *>(r.c_str()));
    }
    if (len) *len = f->rids_c.size();
    return f->rids_c.data();
}

// ---- building a file from scratch (slow5.h:404-447, :510-545; slow5.c:1502-1640, :2160-2220, :3363-3492) -----------------
s5b_rec_t *s5b_rec_init(void) { return static_cast<s5b_rec_t *>(calloc(1, sizeof(s5b_rec_t))); }

int s5b_hdr_add(const char *attr, s5b_hdr_t *header) {   // 0, -1 (NULL argument), -2 (the attribute exists)
    S5bFile *f = owner_of(header);
    if (!attr || !f || !f->writing) return -1;
    for (const auto &kv : f->hdr.attrs)
        if (kv.first == attr) return -2;
    f->hdr.attrs.emplace_back(attr, std::vector<std::string>(f->hdr.num_read_groups));
    return 0;
}

int s5b_hdr_set(const char *attr, const char *value, uint32_t read_group, s5b_hdr_t *header) {   // 0 / -1
    S5bFile *f = owner_of(header);
    if (!attr || !value || !f || !f->writing || read_group >= f->hdr.num_read_groups) return -1;
    for (auto &kv : f->hdr.attrs)
        if (kv.first == attr) {
            kv.second.resize(f->hdr.num_read_groups);
            kv.second[read_group] = value;
            return 0;
        }
    return -1;
}

int64_t s5b_hdr_add_rg(s5b_hdr_t *header) {   // the new read group's number, -1 on error
    S5bFile *f = owner_of(header);
    if (!f || !f->writing) return -1;
    const int64_t rg = f->hdr.num_read_groups++;
    for (auto &kv : f->hdr.attrs) kv.second.resize(f->hdr.num_read_groups);
    f->hdr_priv.pub.num_read_groups = f->hdr.num_read_groups;
    return rg;
}

int s5b_aux_add(const char *field, int type, s5b_hdr_t *header) {   // 0, -1 (bad argument), -2 (exists), -3 (enum: needs labels)
    S5bFile *f = owner_of(header);
    if (!field || !f || !f->writing || type < AUX_INT8 || type > AUX_ENUM_ARRAY) return -1;
    if (type == AUX_ENUM || type == AUX_ENUM_ARRAY) return -3;
    for (const AuxField &a : f->hdr.aux)
        if (a.name == field) return -2;
    static const struct {
        const char *name;
        uint8_t size;
    } prim[] = {{"int8_t", 1}, {"int16_t", 2}, {"int32_t", 4}, {"int64_t", 8}, {"uint8_t", 1}, {"uint16_t", 2},
                {"uint32_t", 4}, {"uint64_t", 8}, {"float", 4}, {"double", 8}, {"char", 1}};
    AuxField a;
    a.name = field;
    a.type = type;
    const bool arr = type >= AUX_INT8_ARRAY;
    const int base = arr ? (type == AUX_STRING ? AUX_CHAR : type - AUX_INT8_ARRAY) : type;
    a.size = prim[base].size;
    a.type_str = std::string(prim[base].name) + (arr ? "*" : "");
    f->hdr.aux.push_back(a);
    return 0;
}

// an enum column with its labels (slow5_aux_add_enum, slow5.c:2222-2334): 0, -1 (bad argument), -2 (exists), -4 (a label is not a C identifier)
int s5b_aux_add_enum(const char *field, const char **labels, uint8_t n_labels, s5b_hdr_t *header) {
    S5bFile *f = owner_of(header);
    if (!field || !labels || !f || !f->writing) return -1;
    for (const AuxField &a : f->hdr.aux)
        if (a.name == field) return -2;
    std::string type = "enum{";
    for (uint8_t i = 0; i < n_labels; ++i) {
        const char *l = labels[i];
        if (!l || !*l || (l[0] >= '0' && l[0] <= '9')) return -4;
        for (const char *c = l; *c; ++c)
            if (!((*c >= 'a' && *c <= 'z') || (*c >= 'A' && *c <= 'Z') || (*c >= '0' && *c <= '9') || *c == '_')) return -4;
        if (i) type += ',';
        type += l;
    }
    type += '}';
    AuxField a;
    a.name = field;
    a.type = AUX_ENUM;
    a.size = 1;
    a.type_str = type;
    f->hdr.aux.push_back(a);
    return 0;
}

}  // extern "C"

namespace {
// number of labels of an enum column, from its type as written in the header ("enum{a,b,c}")
size_t enum_label_count(const AuxField &c) {
    const size_t open = c.type_str.find('{'), close = c.type_str.rfind('}');
    if (open == std::string::npos || close == std::string::npos || close <= open + 1) return 0;
    size_t n = 1;
    for (size_t i = open + 1; i < close; ++i) n += c.type_str[i] == ',';
    return n;
}
// lays the record's binary auxiliary section out for the header's columns from what has been set so far: a column without a
// value gets the type's NULL value, an array without one a zero count (slow5.c:3993-4044)
void aux_relay_pending(s5b_rec_t *r, const std::map<std::string, AuxPending> &pending, const std::vector<AuxField> &cols) {
    std::vector<uint8_t> out;
    for (const AuxField &c : cols) {
        const auto it = pending.find(c.name);
        const bool have = it != pending.end() && it->second.type == c.type;
        if (c.is_array()) {
            const uint64_t n = have ? it->second.len : 0;
            const uint8_t *p = reinterpret_cast<const uint8_t *>(&n);
            out.insert(out.end(), p, p + 8);
            if (have) out.insert(out.end(), it->second.data.begin(), it->second.data.end());
        } else if (have) {
            out.insert(out.end(), it->second.data.begin(), it->second.data.end());
        } else {
            uint8_t v[8] = {0};
            aux_null_value(c.type, v);
            out.insert(out.end(), v, v + c.size);
        }
    }
    free(r->aux);
    r->aux_len = out.size();
    r->aux = static_cast<uint8_t *>(malloc(out.size() ? out.size() : 1));
    if (r->aux && !out.empty()) memcpy(r->aux, out.data(), out.size());
}
// 0, -1 (bad argument), -2 (no such column), -3 (array / primitive mismatch)
int aux_set_any(s5b_rec_t *r, const char *field, const void *data, uint64_t len, bool want_array, bool want_string, s5b_hdr_t *header) {
    S5bFile *f = header ? reinterpret_cast<HdrPriv *>(header)->owner : nullptr;
    if (!r || !field || !data || !f) return -1;
    const std::vector<AuxField> &cols = f->writing ? f->hdr.aux : f->rd.hdr.aux;
    if (cols.empty()) return -1;
    const AuxField *col = nullptr;
    for (const AuxField &c : cols)
        if (c.name == field) col = &c;
    if (!col) return -2;
    if (col->is_array() != want_array || (want_string && col->type != AUX_STRING)) return -3;
    if (col->type == AUX_ENUM || col->type == AUX_ENUM_ARRAY) {  // a value beyond the labels (slow5.c:3388-3396, :3440-3451)
        const size_t nl = enum_label_count(*col);
        for (uint64_t i = 0; i < len; ++i)
            if (static_cast<const uint8_t *>(data)[i] >= nl) return -4;
    }
    RecPriv *pr = static_cast<RecPriv *>(r->aux_meta);
    if (pr && pr->view) {  // a decoded record that is being edited: its stored values are the starting point
        aux_view_free(r);
        pr = nullptr;
    }
    if (!pr) {
        pr = new RecPriv();
        r->aux_meta = pr;
    }
    if (!pr->pending) pr->pending = new std::map<std::string, AuxPending>();
    AuxPending &p = (*pr->pending)[field];
    p.type = col->type;
    p.len = len;
    p.data.assign(static_cast<const uint8_t *>(data), static_cast<const uint8_t *>(data) + len * col->size);
    aux_relay_pending(r, *pr->pending, cols);
    return r->aux ? 0 : -1;
}
}  // namespace

extern "C" {

int s5b_aux_set(s5b_rec_t *read, const char *field, const void *data, s5b_hdr_t *header) {
    return aux_set_any(read, field, data, 1, false, false, header);
}
int s5b_aux_set_array(s5b_rec_t *read, const char *field, const void *data, size_t len, s5b_hdr_t *header) {
    return aux_set_any(read, field, data, len, true, false, header);
}
int s5b_aux_set_string(s5b_rec_t *read, const char *field, const char *data, s5b_hdr_t *header) {
    return aux_set_any(read, field, data, data ? strlen(data) : 0, true, true, header);
}

char *s5b_hdr_get(const char *attr, uint32_t read_group, const s5b_hdr_t *header) {
    if (!attr || !header || read_group >= header->num_read_groups) return nullptr;
    const S5bFile *f = reinterpret_cast<const HdrPriv *>(header)->owner;
    const Header &h = f->writing ? f->hdr : f->rd.hdr;
    for (const auto &kv : h.attrs)
        if (kv.first == attr) return read_group < kv.second.size() ? const_cast<char *>(kv.second[read_group].c_str()) : nullptr;
    return nullptr;
}

}  // extern "C"
