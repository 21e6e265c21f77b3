// view_main.cpp -- `slow5tools-b200 view`: the reference's `slow5tools view` (src/view.c:59-323) with the
// per-record codec work of its worker (depress_parse_rec_to_mem, src/view.c:35-57) moved out of the pthread
// pool (src/thread.c:114 work_db) into batch calls on the GPU through the C-ABI (include/slow5b200.h).
//
// Per batch of -K records:   serial read (as the reference, view.c:265-278)
//   -> s5b_depress_batch_host(ZLIB)      record decompression         (slow5.c:2586)
//   -> host: parse fixed fields, locate the signal bytes               (slow5.c:2811-2950), -t threads
//   -> s5b_depress_batch_host(SVB_ZD)    signal decompression          (slow5.c:2915)
//   -> SLOW5 ASCII formatting on -t threads, or
//      s5b_compress_batch_host(SVB_ZD) -> host: pack records -> s5b_compress_records_host (zlib) (slow5.c:3973,4050)
//   -> serial write (view.c:296-299), end-of-file marker (view.c:313).
// Same flags, defaults and failure behaviour (any record error -> message on stderr, exit status 1).
// There is no CPU codec here: without a CUDA device the command fails.
#include <getopt.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <cerrno>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cinttypes>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/slow5b200.h"
#include "blow5_io.hpp"
#include "cli_common.hpp"

using namespace s5b;

namespace {

#define ERROR(fmt, ...) fprintf(stderr, "[%s::ERROR]\033[1;31m " fmt "\033[0m\n", __func__, __VA_ARGS__)

void parallel_for(size_t n, int threads, const std::function<void(size_t)> &fn) {
    if (threads <= 1 || n < 64) {
        for (size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                const size_t i0 = next.fetch_add(16);
                if (i0 >= n) break;
                const size_t i1 = i0 + 16 < n ? i0 + 16 : n;
                for (size_t i = i0; i < i1; ++i) fn(i);
            }
        });
    for (auto &th : pool) th.join();
}

void usage(FILE *f) {
    fprintf(f,
            "Usage: slow5tools-b200 view [OPTIONS] [SLOW5_FILE/BLOW5_FILE]\n"
            "View a SLOW5/BLOW5 file or convert between the two (GPU codec).\n\n"
            "OPTIONS:\n"
            "    --to FORMAT                   specify output file format (slow5 or blow5)\n"
            "    -o, --output [FILE]           output contents to FILE [stdout]\n"
            "    -c, --compress REC_MTD        record compression method [zlib] (only for blow5 format)\n"
            "    -s, --sig-compress SIG_MTD    signal compression method [svb-zd] (only for blow5 format)\n"
            "    -t, --threads INT             number of host threads for parsing/formatting [8]\n"
            "    -K, --batchsize INT           number of records loaded to the memory at once [4096]\n"
            "    --from FORMAT                 specify input file format (slow5 or blow5)\n"
            "    -h, --help                    display this message and exit\n"
            "REC_MTD: none, zlib, zstd      SIG_MTD: none, svb-zd, ex-zd\n");
}

struct Batch {
    std::vector<std::vector<uint8_t>> mem;     // as read from the file
    std::vector<void *> inflated;              // malloc'd by the library (record method zlib)
    std::vector<size_t> inflated_n;
    std::vector<Record> rec;
    std::vector<std::vector<uint8_t>> aux_store;  // ASCII input: binary form of the aux columns
    std::vector<void *> sig;                   // malloc'd decoded signals
    std::vector<size_t> sig_n;
    void free_all() {
        for (void *p : inflated) free(p);
        for (void *p : sig) free(p);
        inflated.clear();
        sig.clear();
    }
};

// ---- blow5 -> blow5: three-stage pipeline (read | GPU transcode | write) over pinned chunks -------------------
// The reader fills a pinned chunk with packed records straight from the file (one record-boundary walk, no
// per-record allocation); the GPU stage is one s5b_blow5_recode_host call per chunk (one H2D, kernels, one D2H of
// the finished file image); the writer issues one fwrite per chunk.  Replaces the per-record
// malloc/fread ... fwrite/free loops of src/view.c:265-299.
struct Chunk {
    uint8_t *in = nullptr, *out = nullptr;
    uint64_t in_cap = 0, out_cap = 0, in_bytes = 0, out_bytes = 0;
    std::vector<uint64_t> off;
    std::vector<uint32_t> len;
    bool eof = false;
    int err = 0;
};
struct ChunkQueue {
    std::mutex m;
    std::condition_variable cv;
    std::deque<Chunk *> q;
    void push(Chunk *c) {
        {
            std::lock_guard<std::mutex> l(m);
            q.push_back(c);
        }
        cv.notify_one();
    }
    Chunk *pop() {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [&] { return !q.empty(); });
        Chunk *c = q.front();
        q.pop_front();
        return c;
    }
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int view_fast_binary(Reader &rd, FILE *fout, s5b_ctx_t *gpu, int rec_out, int sig_out, long batch) {
    (void)batch;
    const Header &hdr = rd.hdr;
    const bool timing = getenv("S5B_TIMING") != nullptr;
    const double t_begin = now_s();
    double t_gpu = 0, t_wait = 0;
    uint64_t target = 32ull << 20, slack = 8ull << 20;
    if (const char *e = getenv("S5B_VIEW_CHUNK_KB")) {  // test hook: small chunks exercise the carry / grow paths
        const long kb = atol(e);
        if (kb > 0) {
            target = (uint64_t)kb << 10;
            slack = target / 4 + 64;
        }
    }
    const int NCH = 3;
    Chunk chunks[NCH];
    ChunkQueue free_q, full_q, done_q;
    // output is at most ~2.4x the input when decompressing zlib+svb-zd records and smaller when compressing; a
    // chunk whose image does not fit is retried with a larger buffer
    const bool expanding = hdr.record_method != PRESS_NONE || hdr.signal_method != PRESS_NONE;
    auto alloc_chunk = [&](int i) -> bool {
        chunks[i].in_cap = target + slack;
        chunks[i].in = static_cast<uint8_t *>(s5b_host_alloc(chunks[i].in_cap));
        chunks[i].out_cap = (expanding ? 3 : 1) * target + slack;
        chunks[i].out = static_cast<uint8_t *>(s5b_host_alloc(chunks[i].out_cap));
        return chunks[i].in && chunks[i].out;
    };
    // page-locking ~140 MB per chunk takes ~0.1 s: the first chunk is made here, the others while it is being read and transcoded
    // (a chunk that cannot be had just leaves the pipeline with fewer buffers)
    if (!alloc_chunk(0)) {
        ERROR("%s", "cannot allocate pinned staging memory");
        return 1;
    }
    free_q.push(&chunks[0]);
    std::thread more_chunks([&] {
        for (int i = 1; i < NCH; ++i)
            if (alloc_chunk(i)) free_q.push(&chunks[i]);
        if (timing) fprintf(stderr, "[timing] pinned staging complete %.3f s after the start\n", now_s() - t_begin);
    });
    struct Joiner {
        std::thread &t;
        ~Joiner() {
            if (t.joinable()) t.join();
        }
    } join_more_chunks{more_chunks};
    if (timing) fprintf(stderr, "[timing] first pinned chunk %.3f s\n", now_s() - t_begin);
    // the header was consumed through stdio; continue with plain read() from the same position
    const int fd = fileno(rd.fp);
    const off_t start = ftello(rd.fp);
    if (lseek(fd, start, SEEK_SET) < 0) {
        ERROR("%s", "cannot seek in the input file");
        return 1;
    }
    // size of a regular input file (-1: a pipe or the like, read sequentially) and the read position in it
    int64_t in_size = -1;
    uint64_t rpos = (uint64_t)start;
    {
        struct stat ist;
        if (fstat(fd, &ist) == 0 && S_ISREG(ist.st_mode)) in_size = (int64_t)ist.st_size;
    }
    std::string rerr;
    std::thread reader([&] {
        // big sequential read()s straight into the pinned chunk; records are walked in place (u64 size chain,
        // slow5.c:3237-3271) and a record cut by the end of the chunk is carried over to the next one
        std::vector<uint8_t> carry;
        bool eof = false;
        while (!eof) {
            Chunk *c = free_q.pop();
            c->off.clear();
            c->len.clear();
            c->eof = false;
            c->err = 0;
            uint64_t filled = carry.size();
            // The carry starts at a record boundary: its size prefix says how much room that record needs.  A record
            // larger than the chunk (ultra-long reads, uncompressed input) grows the buffer to hold it whole.
            uint64_t need = filled;
            if (filled >= 8) {
                uint64_t size;
                memcpy(&size, carry.data(), 8);
                if (size <= (1ull << 32) - 64 && 8 + size > need) need = 8 + size;
            }
            if (need >= c->in_cap) {
                s5b_host_free(c->in);
                c->in_cap = need + target;
                c->in = static_cast<uint8_t *>(s5b_host_alloc(c->in_cap));
            }
            if (!c->in) {
                rerr = "cannot allocate pinned staging memory";
                c->err = 1;
                c->eof = true;
                full_q.push(c);
                break;
            }
            if (filled) memcpy(c->in, carry.data(), filled);
            carry.clear();
            bool file_end = false;
            if (in_size >= 0 && rpos <= (uint64_t)in_size) {
                // a regular file: the chunk's bytes are known in advance and come in through a few positioned reads side by side
                // (one read() copies out of the page cache at ~3-5 GB/s on one core)
                const uint64_t want = std::min<uint64_t>(c->in_cap - filled, (uint64_t)in_size - rpos);
                const int parts = want >= (8u << 20) ? 4 : 1;
                std::atomic<int> bad(0);
                std::thread th[4];
                for (int k = 0; k < parts; ++k) {
                    const uint64_t a0 = want * k / parts, a1 = want * (k + 1) / parts;
                    auto work = [&, a0, a1] {
                        uint64_t done = a0;
                        while (done < a1) {
                            const ssize_t got = pread(fd, c->in + filled + done, a1 - done, (off_t)(rpos + done));
                            if (got <= 0) {  // (0: the file shrank under us)
                                bad = got < 0 ? errno : EIO;
                                return;
                            }
                            done += (uint64_t)got;
                        }
                    };
                    if (k + 1 < parts) th[k] = std::thread(work);
                    else work();
                }
                for (int k = 0; k + 1 < parts; ++k) th[k].join();
                if (bad) {
                    rerr = std::string("read failed: ") + strerror(bad);
                    c->err = 1;
                    file_end = true;
                } else {
                    filled += want;
                    rpos += want;
                    file_end = rpos == (uint64_t)in_size;
                }
            }
            while (in_size < 0 && filled < c->in_cap) {
                const ssize_t got = read(fd, c->in + filled, c->in_cap - filled);
                if (got < 0) {
                    rerr = std::string("read failed: ") + strerror(errno);
                    c->err = 1;
                    file_end = true;
                    break;
                }
                if (got == 0) {
                    file_end = true;
                    break;
                }
                filled += (uint64_t)got;
            }
            uint64_t pos = 0;
            while (!c->err) {
                const uint64_t left = filled - pos;
                if (file_end && left == 5 && memcmp(c->in + pos, "5WOLB", 5) == 0) {  // end-of-file marker
                    pos += 5;
                    eof = true;
                    break;
                }
                if (left < 8) {
                    if (file_end) {
                        rerr = "blow5 file is truncated or has no end-of-file marker";
                        c->err = 1;
                    }
                    break;
                }
                uint64_t size;
                memcpy(&size, c->in + pos, 8);
                if (size > (1ull << 32) - 64) {
                    rerr = "implausible record size (corrupt file?)";
                    c->err = 1;
                    break;
                }
                if (left < 8 + size) {
                    if (file_end) {
                        rerr = "blow5 record is truncated";
                        c->err = 1;
                    }
                    break;
                }
                c->off.push_back(pos + 8);
                c->len.push_back((uint32_t)size);
                pos += 8 + size;
            }
            if (c->err) eof = true;
            if (!eof) {
                if (pos == 0 && !file_end && filled == c->in_cap) {
                    // not even one whole record fits: carry everything, the next round grows the buffer
                    carry.assign(c->in, c->in + filled);
                    uint64_t size;
                    memcpy(&size, c->in, 8);
                    carry.reserve(size + 8);
                } else {
                    carry.assign(c->in + pos, c->in + filled);
                }
                if (file_end && carry.empty() && c->len.empty()) {  // file ended without a marker
                    rerr = "blow5 file is truncated or has no end-of-file marker";
                    c->err = 1;
                    eof = true;
                }
            }
            c->in_bytes = pos;
            c->eof = eof;
            full_q.push(c);
        }
    });
    int wret = 0;
    const int ofd = fileno(fout);
    fflush(fout);
    const off_t wstart = lseek(ofd, 0, SEEK_CUR);
    // positioned writes need a regular file that honours the offset: with O_APPEND (shell `>>`) pwrite() appends
    // wherever the parts happen to finish, so such descriptors take the sequential path like pipes do
    struct stat ost;
    const int oflags = fcntl(ofd, F_GETFL);
    const bool seekable = wstart >= 0 && fstat(ofd, &ost) == 0 && S_ISREG(ost.st_mode) && oflags >= 0 && !(oflags & O_APPEND);
    uint64_t wpos = seekable ? (uint64_t)wstart : 0;
    std::thread writer([&] {
        for (;;) {
            Chunk *c = done_q.pop();
            if (!c) break;
            // a regular file takes the chunk as parallel pwrite()s (page-cache / tmpfs copies scale with threads);
            // pipes and terminals get one sequential write
            const uint64_t nb = c->out_bytes;
            if (seekable && nb > (8u << 20)) {
                const int parts = 4;
                std::thread th[parts];
                std::atomic<int> bad(0);
                for (int k = 0; k < parts; ++k) {
                    const uint64_t a0 = nb * k / parts, a1 = nb * (k + 1) / parts;
                    th[k] = std::thread([&, a0, a1] {
                        uint64_t done = a0;
                        while (done < a1) {
                            const ssize_t w = pwrite(ofd, c->out + done, a1 - done, (off_t)(wpos + done));
                            if (w <= 0) {
                                bad = 1;
                                return;
                            }
                            done += (uint64_t)w;
                        }
                    });
                }
                for (int k = 0; k < parts; ++k) th[k].join();
                if (bad) wret = 1;
                wpos += nb;
            } else {
                // small chunk (or a pipe): one sequential write.  On a seekable file it must be a pwrite at wpos as well:
                // the pwrite()s above never move the descriptor's own offset, so a plain write() here would land at
                // the start of the record area and overwrite the first records of the file
                uint64_t done = 0;
                while (!wret && done < nb) {
                    const ssize_t w = seekable ? pwrite(ofd, c->out + done, nb - done, (off_t)(wpos + done))
                                               : write(ofd, c->out + done, nb - done);
                    if (w <= 0) wret = 1;
                    else done += (uint64_t)w;
                }
                wpos += nb;
            }
            free_q.push(c);
        }
    });
    int ret = 0;
    for (;;) {
        const double tw = now_s();
        Chunk *c = full_q.pop();
        t_wait += now_s() - tw;
        const double tg = now_s();
        const bool last = c->eof;
        if (c->err && ret == 0) {
            ERROR("%s", rerr.c_str());
            ret = 1;
        }
        c->out_bytes = 0;
        if (ret == 0 && !c->len.empty()) {
            int rc = s5b_blow5_recode_host(gpu, hdr.record_method, hdr.signal_method, rec_out, sig_out, c->in, c->in_bytes,
                                           c->off.data(), c->len.data(), c->len.size(), c->out, c->out_cap, &c->out_bytes);
            if (rc == S5B_ERR_NOSPACE && c->out_bytes > c->out_cap) {  // image larger than the staging buffer: grow, retry
                s5b_host_free(c->out);
                c->out_cap = c->out_bytes + (8u << 20);
                c->out = static_cast<uint8_t *>(s5b_host_alloc(c->out_cap));
                rc = c->out ? s5b_blow5_recode_host(gpu, hdr.record_method, hdr.signal_method, rec_out, sig_out, c->in, c->in_bytes,
                                                    c->off.data(), c->len.data(), c->len.size(), c->out, c->out_cap, &c->out_bytes)
                            : S5B_ERR_MEM;
            }
            if (rc != S5B_OK) {
                ERROR("record conversion failed: %s", s5b_strerror(rc));
                ret = 1;
                c->out_bytes = 0;
            }
        }
        t_gpu += now_s() - tg;
        done_q.push(c);
        if (last) break;
    }
    reader.join();
    done_q.push(nullptr);
    writer.join();
    if (seekable && lseek(ofd, (off_t)wpos, SEEK_SET) < 0) wret = 1;  // the end-of-file marker goes after the last chunk
    if (wret) {
        ERROR("%s", "writing the output failed");
        ret = 1;
    }
    if (timing)
        fprintf(stderr, "[timing] fast path total %.3f s: gpu calls %.3f s, waiting for the reader %.3f s\n", now_s() - t_begin,
                t_gpu, t_wait);
    // the pinned chunks are left to process teardown (freeing ~0.5 GB of page-locked memory costs more than it is worth)
    return ret;
}

}  // namespace

int blow5_fast_convert(Reader &rd, FILE *fout, s5b_ctx_t *gpu, int rec_out, int sig_out) {
    return view_fast_binary(rd, fout, gpu, rec_out, sig_out, 0);
}

// The general conversion loop (the body of slow5_convert_parallel, src/view.c:254-301, and of `get`'s per-batch work,
// src/get.c:37-66): records come from `next` -- 1 = one stored record (binary: without its size prefix; text: one line
// without the newline) placed in the buffer, 0 = no more, -1 = error already reported -- are decompressed, parsed,
// re-encoded for (fmt_out, rec_out, sig_out) with the codec calls batched on the GPU, and written to fout in order.
int convert_records(const Header &hdr, Fmt fmt_in, const std::function<int(std::vector<uint8_t> &)> &next, FILE *fout,
                    s5b_ctx_t *gpu, Fmt fmt_out, int rec_out, int sig_out, long batch, int threads, const ConvertHooks *hooks) {
    int ret = 0;
    Batch b;
    const Header &hdr_o = hooks && hooks->hdr_out ? *hooks->hdr_out : hdr;  // what the output records follow
    // record i of the batch goes to fout, to the file its route names, or to every file of route_many (none: it is dropped)
    std::vector<FILE *> one(1);
    auto put = [&](size_t i, const void *head, size_t head_n, const void *body, size_t body_n) -> bool {
        const std::vector<FILE *> *to = &one;
        if (hooks && hooks->route_many) to = hooks->route_many(i);
        else one[0] = hooks && hooks->route ? hooks->route(i) : fout;
        for (FILE *f : *to)
            if ((head_n && fwrite(head, 1, head_n, f) != head_n) || (body_n && fwrite(body, 1, body_n, f) != body_n)) return false;
        return true;
    };
    bool eof = false;
    while (!eof && ret == 0) {
        // ---- load (serial)
        b.mem.clear();
        while ((long)b.mem.size() < batch) {
            b.mem.emplace_back();
            const int rc = next(b.mem.back());
            if (rc <= 0) {
                b.mem.pop_back();
                if (rc < 0) ret = 1;  // the source has reported what went wrong
                eof = true;
                break;
            }
        }
        const size_t n = b.mem.size();
        if (n == 0 || ret) break;
        b.rec.assign(n, Record());
        b.aux_store.assign(n, std::vector<uint8_t>());
        std::vector<const void *> ptrs(n);
        std::vector<size_t> counts(n);

        // ---- record decompression
        std::vector<const uint8_t *> packed(n);
        std::vector<size_t> packed_n(n);
        if (fmt_in == FMT_BINARY && (hdr.record_method == PRESS_ZLIB || hdr.record_method == PRESS_ZSTD)) {
            for (size_t i = 0; i < n; ++i) {
                ptrs[i] = b.mem[i].data();
                counts[i] = b.mem[i].size();
            }
            b.inflated.assign(n, nullptr);
            b.inflated_n.assign(n, 0);
            const int rc = s5b_depress_batch_host(gpu, hdr.record_method == PRESS_ZLIB ? S5B_COMPRESS_ZLIB : S5B_COMPRESS_ZSTD,
                                                  ptrs.data(), counts.data(), n, b.inflated.data(), b.inflated_n.data());
            if (rc != S5B_OK) {
                ERROR("record decompression failed: %s", s5b_strerror(rc));
                ret = 1;
                break;
            }
            for (size_t i = 0; i < n; ++i) {
                packed[i] = static_cast<const uint8_t *>(b.inflated[i]);
                packed_n[i] = b.inflated_n[i];
            }
        } else {
            for (size_t i = 0; i < n; ++i) {
                packed[i] = b.mem[i].data();
                packed_n[i] = b.mem[i].size();
            }
        }
        // ---- parse
        std::atomic<int> bad(0);
        if (fmt_in == FMT_BINARY) {
            parallel_for(n, threads, [&](size_t i) {
                std::string e;
                if (!record_parse_binary(packed[i], packed_n[i], hdr, hdr.signal_method, b.rec[i], e)) bad = 1;
            });
        } else {
            // with a GPU context the raw_signal column is converted there, a batch at a time (slow5.c:2754-2778)
            const bool defer = gpu != nullptr;
            parallel_for(n, threads, [&](size_t i) {
                std::string e;
                if (!record_parse_ascii(reinterpret_cast<const char *>(packed[i]), packed_n[i], hdr, b.rec[i], b.aux_store[i], e, defer))
                    bad = 1;
            });
        }
        if (bad) {
            ERROR("%s", "a record could not be parsed");
            ret = 1;
            break;
        }
        if (hooks && hooks->transform) {
            for (size_t i = 0; i < n && ret == 0; ++i)
                if (!hooks->transform(i, b.rec[i], b.aux_store[i])) ret = 1;
            if (ret) break;
        }
        // ---- signal decompression
        std::vector<const int16_t *> sig(n);
        // text output with a GPU context: the stored signals go straight to the device formatter (decode + sprintf loop of
        // slow5.c:3866-3878 in one trip), nothing is decompressed to the host
        const int qts = hooks ? hooks->qts_bits : 0;  // degrade needs the samples themselves in between
        const bool gpu_text = fmt_out == FMT_ASCII && gpu != nullptr && fmt_in == FMT_BINARY && !qts;
        if (fmt_in == FMT_ASCII && gpu != nullptr) {
            std::vector<const char *> tptrs(n);
            std::vector<uint64_t> expect(n);
            for (size_t i = 0; i < n; ++i) {
                tptrs[i] = reinterpret_cast<const char *>(b.rec[i].sig_bytes);
                counts[i] = b.rec[i].sig_nbytes;
                expect[i] = b.rec[i].len_raw_signal;
            }
            b.sig.assign(n, nullptr);
            b.sig_n.assign(n, 0);
            std::vector<int16_t *> outp(n, nullptr);
            const int rc = s5b_ascii_to_signal_batch_host(gpu, tptrs.data(), counts.data(), expect.data(), n, outp.data(), b.sig_n.data());
            for (size_t i = 0; i < n; ++i) b.sig[i] = outp[i];
            if (rc != S5B_OK) {
                ERROR("%s", "a record could not be parsed");
                ret = 1;
                break;
            }
            for (size_t i = 0; i < n; ++i) {
                sig[i] = static_cast<const int16_t *>(b.sig[i]);
                b.sig_n[i] *= 2;  // bytes, like the codec calls report
                b.rec[i].sig_bytes = reinterpret_cast<const uint8_t *>(b.sig[i]);
                b.rec[i].sig_nbytes = b.sig_n[i];
            }
        } else if (gpu_text) {
            for (size_t i = 0; i < n; ++i) sig[i] = nullptr;
        } else if (fmt_in == FMT_BINARY && hdr.signal_method != PRESS_NONE) {  // PRESS_* == S5B_COMPRESS_* (slow5_press.h:61-67)
            for (size_t i = 0; i < n; ++i) {
                ptrs[i] = b.rec[i].sig_bytes;
                counts[i] = b.rec[i].sig_nbytes;
            }
            b.sig.assign(n, nullptr);
            b.sig_n.assign(n, 0);
            const int rc = s5b_depress_batch_host(gpu, hdr.signal_method, ptrs.data(), counts.data(), n, b.sig.data(),
                                                  b.sig_n.data());
            if (rc != S5B_OK) {
                ERROR("signal decompression failed: %s", s5b_strerror(rc));
                ret = 1;
                break;
            }
            for (size_t i = 0; i < n; ++i) {
                sig[i] = static_cast<const int16_t *>(b.sig[i]);
                b.rec[i].len_raw_signal = b.sig_n[i] / 2;
            }
        } else {
            for (size_t i = 0; i < n; ++i) {
                sig[i] = reinterpret_cast<const int16_t *>(b.rec[i].sig_bytes);  // memcpy'd below where alignment matters
                b.rec[i].len_raw_signal = b.rec[i].sig_nbytes / 2;
            }
        }

        // ---- degrade: slow5_rec_qts_round on every record of the batch (src/degrade.c:255), one trip to the device
        std::vector<void *> degraded;
        struct FreeAll {
            std::vector<void *> &v;
            ~FreeAll() { for (void *p : v) free(p); }
        } free_degraded{degraded};
        if (qts) {
            degraded.assign(n, nullptr);
            std::vector<size_t> got(n, 0);
            for (size_t i = 0; i < n; ++i) {
                ptrs[i] = sig[i];
                counts[i] = b.rec[i].len_raw_signal * 2;
            }
            const int rc = gpu ? s5b_qts_round_batch_host(gpu, qts, ptrs.data(), counts.data(), n, degraded.data(), got.data()) : S5B_ERR_DEVICE;
            if (rc != S5B_OK) {
                ERROR("signal degradation failed: %s", s5b_strerror(rc));
                ret = 1;
                break;
            }
            for (size_t i = 0; i < n; ++i) sig[i] = static_cast<const int16_t *>(degraded[i]);
        }

        // ---- output
        if (qts && fmt_out == FMT_ASCII) {
            // the degraded samples go back to the device formatter as raw int16 (slow5.c:3866-3878)
            std::vector<char *> text(n, nullptr);
            std::vector<size_t> text_n(n, 0);
            for (size_t i = 0; i < n; ++i) {
                ptrs[i] = sig[i];
                counts[i] = b.rec[i].len_raw_signal * 2;
            }
            const int rc = s5b_signal_to_ascii_batch_host(gpu, PRESS_NONE, ptrs.data(), counts.data(), n, text.data(), text_n.data());
            if (rc != S5B_OK) {
                ERROR("signal formatting failed: %s", s5b_strerror(rc));
                for (char *p : text) free(p);
                ret = 1;
                break;
            }
            std::vector<std::string> lines(n);
            parallel_for(n, threads, [&](size_t i) {
                record_to_ascii(b.rec[i], hdr_o, lines[i], text[i], text_n[i]);
                free(text[i]);
            });
            for (size_t i = 0; i < n && ret == 0; ++i)
                if (!put(i, nullptr, 0, lines[i].data(), lines[i].size())) ret = 1;
        } else if (gpu_text) {
            std::vector<char *> text(n, nullptr);
            std::vector<size_t> text_n(n, 0);
            for (size_t i = 0; i < n; ++i) {
                ptrs[i] = b.rec[i].sig_bytes;
                counts[i] = b.rec[i].sig_nbytes;
            }
            const int rc = s5b_signal_to_ascii_batch_host(gpu, hdr.signal_method, ptrs.data(), counts.data(), n, text.data(), text_n.data());
            if (rc != S5B_OK) {
                ERROR("signal decompression failed: %s", s5b_strerror(rc));
                for (char *p : text) free(p);
                ret = 1;
                break;
            }
            std::vector<std::string> lines(n);
            parallel_for(n, threads, [&](size_t i) {
                Record &r = b.rec[i];
                // the sample count of the len_raw_signal column: what the stored stream says it holds
                const uint8_t *sb = r.sig_bytes;
                uint64_t ns = r.sig_nbytes / 2;
                if (hdr.signal_method == PRESS_SVB_ZD) {
                    uint32_t v = 0;
                    if (r.sig_nbytes >= 4) memcpy(&v, sb, 4);
                    ns = v;
                } else if (hdr.signal_method == PRESS_EX_ZD) {
                    ns = 0;
                    if (r.sig_nbytes >= 9) memcpy(&ns, sb + 1, 8);
                }
                r.len_raw_signal = ns;
                record_to_ascii(r, hdr_o, lines[i], text[i], text_n[i]);
                free(text[i]);
            });
            for (size_t i = 0; i < n && ret == 0; ++i)
                if (!put(i, nullptr, 0, lines[i].data(), lines[i].size())) ret = 1;
        } else if (fmt_out == FMT_ASCII) {
            std::vector<std::string> lines(n);
            parallel_for(n, threads, [&](size_t i) {
                Record &r = b.rec[i];
                r.raw_signal.resize(r.len_raw_signal);
                if (r.len_raw_signal) memcpy(r.raw_signal.data(), sig[i], r.len_raw_signal * 2);
                record_to_ascii(r, hdr_o, lines[i]);
                std::vector<int16_t>().swap(r.raw_signal);
            });
            for (size_t i = 0; i < n && ret == 0; ++i)
                if (!put(i, nullptr, 0, lines[i].data(), lines[i].size())) ret = 1;
        } else {
            // signal compression
            std::vector<void *> svb(n, nullptr);
            std::vector<size_t> svb_n(n, 0);
            std::vector<const uint8_t *> store(n);
            std::vector<size_t> store_n(n);
            if (sig_out != PRESS_NONE) {
                for (size_t i = 0; i < n; ++i) {
                    ptrs[i] = sig[i];
                    counts[i] = b.rec[i].len_raw_signal * 2;
                }
                const int rc = s5b_compress_batch_host(gpu, sig_out, ptrs.data(), counts.data(), n, svb.data(), svb_n.data());
                if (rc != S5B_OK) {
                    ERROR("signal compression failed: %s", s5b_strerror(rc));
                    ret = 1;
                }
                for (size_t i = 0; i < n; ++i) {
                    store[i] = static_cast<const uint8_t *>(svb[i]);
                    store_n[i] = svb_n[i];
                }
            } else {
                for (size_t i = 0; i < n; ++i) {
                    store[i] = reinterpret_cast<const uint8_t *>(sig[i]);
                    store_n[i] = b.rec[i].len_raw_signal * 2;
                }
            }
            std::vector<std::vector<uint8_t>> rec_mem(n);
            std::vector<uint32_t> splits(n, 0);
            if (ret == 0) {
                parallel_for(n, threads, [&](size_t i) {
                    uint64_t at = 0;
                    record_to_binary(b.rec[i], store[i], store_n[i], sig_out != PRESS_NONE, rec_mem[i], &at);
                    // Huffman block split for the zlib encoder: where the svb-zd data bytes start
                    if (sig_out == PRESS_SVB_ZD) splits[i] = (uint32_t)(at + 4 + (b.rec[i].len_raw_signal + 3) / 4);
                });
            }
            for (void *p : svb) free(p);
            std::vector<void *> z(n, nullptr);
            std::vector<size_t> z_n(n, 0);
            const bool rec_packed = rec_out == PRESS_ZLIB || rec_out == PRESS_ZSTD;
            if (ret == 0 && rec_packed) {
                for (size_t i = 0; i < n; ++i) {
                    ptrs[i] = rec_mem[i].data();
                    counts[i] = rec_mem[i].size();
                }
                const int rc = s5b_compress_records_host(gpu, rec_out == PRESS_ZSTD ? S5B_COMPRESS_ZSTD : S5B_COMPRESS_ZLIB,
                                                         ptrs.data(), counts.data(), splits.data(), n, z.data(), z_n.data());
                if (rc != S5B_OK) {
                    ERROR("record compression failed: %s", s5b_strerror(rc));
                    ret = 1;
                }
            }
            for (size_t i = 0; i < n && ret == 0; ++i) {
                const void *p = rec_packed ? z[i] : rec_mem[i].data();
                const uint64_t sz = rec_packed ? z_n[i] : rec_mem[i].size();
                if (!put(i, &sz, 8, p, sz)) ret = 1;  // slow5.c:4055-4060
            }
            for (void *p : z) free(p);
        }
        if (ret) ERROR("%s", "writing the output failed");
        b.free_all();
    }
    b.free_all();
    return ret;
}

int view_main(int argc, char **argv) {
    static const struct option long_opts[] = {
        {"sig-compress", required_argument, nullptr, 's'}, {"compress", required_argument, nullptr, 'c'},
        {"from", required_argument, nullptr, 'f'},         {"help", no_argument, nullptr, 'h'},
        {"output", required_argument, nullptr, 'o'},       {"to", required_argument, nullptr, 'b'},
        {"threads", required_argument, nullptr, 't'},      {"batchsize", required_argument, nullptr, 'K'},
        {nullptr, 0, nullptr, 0}};
    const char *arg_sig = nullptr, *arg_rec = nullptr, *arg_from = nullptr, *arg_to = nullptr, *arg_out = nullptr;
    int threads = 8;
    long batch = 4096;
    int opt;
    optind = 1;
    while ((opt = getopt_long(argc, argv, "s:c:f:ho:b:t:K:", long_opts, nullptr)) != -1) {
        switch (opt) {
            case 's': arg_sig = optarg; break;
            case 'c': arg_rec = optarg; break;
            case 'f': arg_from = optarg; break;
            case 'b': arg_to = optarg; break;
            case 'o': arg_out = optarg; break;
            case 't': threads = atoi(optarg); break;
            case 'K': batch = atol(optarg); break;
            case 'h': usage(stdout); return 0;
            default: usage(stderr); return 1;
        }
    }
    if (threads < 1 || batch < 1) {
        ERROR("%s", "invalid -t / -K value");
        return 1;
    }
    if (optind >= argc) {
        ERROR("missing input file%s", "");
        usage(stderr);
        return 1;
    }
    if (optind != argc - 1) {
        ERROR("more than 1 input file is given%s", "");
        return 1;
    }
    const char *in_path = argv[optind];
    Fmt fmt_in = FMT_UNKNOWN, fmt_out = FMT_UNKNOWN;
    if (arg_from && (fmt_in = fmt_from_name(arg_from)) == FMT_UNKNOWN) {
        ERROR("invalid input format '%s'", arg_from);
        return 1;
    }
    if (arg_to && (fmt_out = fmt_from_name(arg_to)) == FMT_UNKNOWN) {
        ERROR("invalid output format '%s'", arg_to);
        return 1;
    }
    if (arg_out) {
        const Fmt by_ext = fmt_from_path(arg_out);
        if (fmt_out == FMT_UNKNOWN) {
            fmt_out = by_ext;
            if (fmt_out == FMT_UNKNOWN) {
                ERROR("cannot detect the output format from the file extension of '%s'", arg_out);
                return 1;
            }
        } else if (by_ext != FMT_UNKNOWN && by_ext != fmt_out) {
            ERROR("output file extension '%s' does not match the output format '%s'", arg_out, arg_to);
            return 1;
        }
    }
    if (fmt_out == FMT_UNKNOWN) fmt_out = FMT_ASCII;  // view.c:160-162
    if (fmt_out == FMT_ASCII && (arg_rec || arg_sig)) {  // misc.c:219-249
        ERROR("%s", "compression options (-c / -s) are only valid for blow5 output");
        return 1;
    }
    int rec_out = PRESS_ZLIB, sig_out = PRESS_SVB_ZD;  // misc.c:54-55
    if (arg_rec && (rec_out = press_from_name(arg_rec)) == PRESS_BAD) {
        ERROR("invalid record compression method '%s'", arg_rec);
        return 1;
    }
    if (arg_sig && (sig_out = press_from_name(arg_sig)) == PRESS_BAD) {
        ERROR("invalid signal compression method '%s'", arg_sig);
        return 1;
    }
    if (fmt_out == FMT_ASCII) rec_out = sig_out = PRESS_NONE;
    if ((rec_out != PRESS_NONE && rec_out != PRESS_ZLIB && rec_out != PRESS_ZSTD) ||
        (sig_out != PRESS_NONE && sig_out != PRESS_SVB_ZD && sig_out != PRESS_EX_ZD)) {
        ERROR("%s", "this build supports record compression none/zlib/zstd and signal compression none/svb-zd/ex-zd only");
        return 1;
    }

    Reader rd;
    if (!reader_open(rd, in_path, fmt_in)) {
        ERROR("File '%s' could not be opened - %s.", in_path, rd.err.c_str());
        return 1;
    }
    const Header &hdr = rd.hdr;
    if ((hdr.record_method != PRESS_NONE && hdr.record_method != PRESS_ZLIB && hdr.record_method != PRESS_ZSTD) ||
        (hdr.signal_method != PRESS_NONE && hdr.signal_method != PRESS_SVB_ZD && hdr.signal_method != PRESS_EX_ZD)) {
        ERROR("%s", "input uses a compression method this build does not support (zlib/zstd as signal method)");
        return 1;
    }
    FILE *fout = stdout;
    if (arg_out && !(fout = fopen(arg_out, "wb"))) {
        ERROR("File '%s' could not be opened - %s.", arg_out, strerror(errno));
        return 1;
    }
    setvbuf(fout, nullptr, _IOFBF, 1 << 20);

    const bool need_gpu = hdr.record_method != PRESS_NONE || hdr.signal_method != PRESS_NONE || rec_out != PRESS_NONE ||
                          sig_out != PRESS_NONE;
    s5b_ctx_t *gpu = nullptr;
    const double t_ctx = now_s();
    if (need_gpu) {
        const int rc = s5b_ctx_create(-1, &gpu);
        if (rc != S5B_OK) {
            ERROR("cannot initialise the GPU codec: %s", s5b_strerror(rc));
            return 1;
        }
    }
    if (getenv("S5B_TIMING")) fprintf(stderr, "[timing] context create %.3f s\n", now_s() - t_ctx);
    {
        const std::string h = header_to_mem(hdr, fmt_out, rec_out, sig_out);
        if (fwrite(h.data(), 1, h.size(), fout) != h.size()) {
            ERROR("%s", "could not write the header");
            return 1;
        }
    }

    int ret = 0;
    bool eof = false;
    if (rd.fmt == FMT_BINARY && fmt_out == FMT_BINARY && need_gpu && !getenv("S5B_VIEW_SLOW_PATH")) {
        // blow5 -> blow5: whole batches stay on the device (pinned chunk pipeline); the device checks every record's auxiliary
        // section against the header's columns like record_parse_binary does on the other path
        if (hdr.aux.size() <= 64) {
            std::vector<uint8_t> sz(hdr.aux.size() + 1), arr(hdr.aux.size() + 1);
            for (size_t f = 0; f < hdr.aux.size(); ++f) {
                sz[f] = hdr.aux[f].size;
                arr[f] = hdr.aux[f].is_array() ? 1 : 0;
            }
            s5b_ctx_set_aux_layout(gpu, sz.data(), arr.data(), (uint32_t)hdr.aux.size());
        }
        ret = view_fast_binary(rd, fout, gpu, rec_out, sig_out, batch);
        eof = true;
    }
    if (!eof && ret == 0)
        ret = convert_records(hdr, rd.fmt, [&](std::vector<uint8_t> &mem) {
            const int rc = reader_next_mem(rd, mem);
            if (rc < 0) ERROR("%s", rd.err.c_str());
            return rc;
        }, fout, gpu, fmt_out, rec_out, sig_out, batch, threads);
    fflush(fout);
    if (ret == 0 && fmt_out == FMT_BINARY && write(fileno(fout), "5WOLB", 5) != 5) ret = 1;  // view.c:313
    if (fout != stdout) {
        if (fclose(fout) != 0) ret = 1;
    } else {
        fflush(fout);
    }
    reader_close(rd);
    // the context (streams, device slabs) is left to process teardown: an orderly destroy costs ~0.1 s for nothing
    if (gpu && getenv("S5B_ORDERLY_EXIT")) s5b_ctx_destroy(gpu);
    return ret;
}

int index_main(int argc, char **argv);  // index_main.cpp
int get_main(int argc, char **argv);    // get_main.cpp
int merge_main(int argc, char **argv);  // merge_split_main.cpp
int split_main(int argc, char **argv);
int degrade_main(int argc, char **argv);  // degrade_main.cpp

static int run_command(int argc, char **argv);

int main(int argc, char **argv) {
    const int rc = run_command(argc, argv);
    // Every output file has been closed by its sub-command.  The CUDA runtime's orderly teardown at exit (primary context,
    // page-locked buffers, device slabs) costs a few tenths of a second and buys nothing here: leave it to the kernel.
    fflush(nullptr);
    if (!getenv("S5B_ORDERLY_EXIT")) _exit(rc);
    return rc;
}

static int run_command(int argc, char **argv) {
    if (argc >= 2 && (!strcmp(argv[1], "--version") || !strcmp(argv[1], "-V"))) {
        printf("slow5tools-b200 %s\n", s5b_version());
        return 0;
    }
    if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
        fprintf(argc < 2 ? stderr : stdout, "Usage: slow5tools-b200 <command> [options]\n\nCOMMANDS:\n    view    view the contents of a SLOW5/BLOW5 file or convert between different formats and compressions\n    index   create a SLOW5/BLOW5 index file\n    get     display the read entry for each specified read id\n    merge   merge multiple SLOW5/BLOW5 files to a single file\n    split   split a SLOW5/BLOW5 file by read group, number of reads or number of files\n    degrade irreversibly degrade the signals (lossy) and convert a SLOW5/BLOW5 file\n");
        return argc < 2 ? 1 : 0;
    }
    if (!strcmp(argv[1], "get")) {
        const int rc = get_main(argc - 1, argv + 1);
        if (rc != 0) {
            fprintf(stderr, "[main::ERROR] get failed\n");
            return EXIT_FAILURE;
        }
        return 0;
    }
    if (!strcmp(argv[1], "index")) {
        const int rc = index_main(argc - 1, argv + 1);
        if (rc != 0) {
            fprintf(stderr, "[main::ERROR] index failed\n");
            return EXIT_FAILURE;
        }
        return 0;
    }
    if (!strcmp(argv[1], "merge") || !strcmp(argv[1], "split")) {
        const int rc = argv[1][0] == 'm' ? merge_main(argc - 1, argv + 1) : split_main(argc - 1, argv + 1);
        if (rc != 0) {
            fprintf(stderr, "[main::ERROR] %s failed\n", argv[1]);
            return EXIT_FAILURE;
        }
        return 0;
    }
    if (!strcmp(argv[1], "degrade")) {
        const int rc = degrade_main(argc - 1, argv + 1);
        if (rc != 0) {
            fprintf(stderr, "[main::ERROR] degrade failed\n");
            return EXIT_FAILURE;
        }
        return 0;
    }
    if (!strcmp(argv[1], "view")) {
        const int rc = view_main(argc - 1, argv + 1);
        if (rc != 0) {
            fprintf(stderr, "[main::ERROR] view failed\n");
            return EXIT_FAILURE;
        }
        return 0;
    }
    fprintf(stderr, "[main::ERROR] unrecognised command '%s' (this build provides the hot path only: view, index, get, merge, split, degrade)\n", argv[1]);
    return EXIT_FAILURE;
}
