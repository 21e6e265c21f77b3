/* oracle/ref_driver.c -- TEST / BASELINE INFRASTRUCTURE (see oracle.h).
 * Times the CPU implementation of the hot path over a batch with a fork-join pthread pool shaped like
 * the reference's own (src/thread.c:69-111 pthread_db: static split + per-thread loop), calling, per
 * read, exactly what the reference's workers call:
 *     slow5_ptr_compress_solo(SLOW5_COMPRESS_SVB_ZD, ...)   slow5lib/src/slow5_press.c:330
 *     slow5_ptr_depress_solo (SLOW5_COMPRESS_SVB_ZD, ...)   slow5lib/src/slow5_press.c:439
 * When `ref_lib_path` names the compiled, UNMODIFIED reference (oracle/_ref/libslow5_ref.so) the
 * entry points are taken from it with dlopen ("reference" baseline); when it is NULL the oracle's own
 * restatement is used ("port" baseline).  Used by bench.py's cpu_baseline leg and --impl reference.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef void *(*solo_fn)(int method, const void *ptr, size_t count, size_t *n);

static void *port_compress(int method, const void *ptr, size_t count, size_t *n) {
    (void) method;
    uint32_t ns = (uint32_t) (count / 2);
    uint8_t *out = (uint8_t *) malloc(orc_svbzd_bound(ns));
    if (!out) { *n = 0; return NULL; }
    *n = orc_svbzd_compress((const int16_t *) ptr, count, out);
    return out;
}
static void *port_depress(int method, const void *ptr, size_t count, size_t *n) {
    (void) method;
    uint32_t ns = 0;
    if (count >= 4) memcpy(&ns, ptr, 4);
    int16_t *out = (int16_t *) malloc((size_t) ns * 2 + 2);
    if (!out) { *n = 0; return NULL; }
    if (orc_svbzd_depress((const uint8_t *) ptr, count, out, ns, &ns) != 0) { free(out); *n = 0; return NULL; }
    *n = (size_t) ns * 2;
    return out;
}

typedef struct {
    solo_fn fn;
    int method;
    const uint8_t *base;       /* input slab */
    const uint64_t *off;       /* byte offsets */
    const uint64_t *len;       /* byte lengths */
    void **out;
    size_t *out_n;
    uint64_t n_reads;
    uint64_t next;             /* shared work index (the reference steals the same way, thread.c:54) */
    uint64_t grain;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *) arg;
    for (;;) {
        uint64_t r0 = __sync_fetch_and_add(&j->next, j->grain);
        if (r0 >= j->n_reads) break;
        uint64_t r1 = r0 + j->grain < j->n_reads ? r0 + j->grain : j->n_reads;
        for (uint64_t r = r0; r < r1; ++r)
            j->out[r] = j->fn(j->method, j->base + j->off[r], (size_t) j->len[r], &j->out_n[r]);
    }
    return NULL;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static double run_pool(job_t *j, int threads) {
    pthread_t *t = (pthread_t *) malloc(sizeof(pthread_t) * (size_t) threads);
    double t0 = now_s();
    for (int i = 0; i < threads; ++i) pthread_create(&t[i], NULL, worker, j);   /* fork-join per batch */
    for (int i = 0; i < threads; ++i) pthread_join(t[i], NULL);
    double dt = now_s() - t0;
    free(t);
    return dt;
}

/* Encodes then decodes every read of the batch; returns 0 and fills the timings, or a negative code:
 * -1 dlopen/dlsym failed, -2 a codec call failed, -3 round trip mismatch. */
int refdrv_svbzd_roundtrip(const char *ref_lib_path, const int16_t *sig, const uint64_t *sig_off,
                           const uint32_t *n_samples, uint64_t n_reads, int threads,
                           double *enc_seconds, double *dec_seconds, uint64_t *svb_bytes) {
    solo_fn cfn = port_compress, dfn = port_depress;
    void *h = NULL;
    if (ref_lib_path) {
        h = dlopen(ref_lib_path, RTLD_NOW | RTLD_LOCAL);
        if (!h) return -1;
        cfn = (solo_fn) dlsym(h, "slow5_ptr_compress_solo");
        dfn = (solo_fn) dlsym(h, "slow5_ptr_depress_solo");
        if (!cfn || !dfn) { dlclose(h); return -1; }
    }
    if (threads < 1) threads = 1;
    int rc = 0;
    uint64_t *off = (uint64_t *) malloc(8 * n_reads), *len = (uint64_t *) malloc(8 * n_reads);
    void **enc = (void **) calloc(n_reads, sizeof(void *)), **dec = (void **) calloc(n_reads, sizeof(void *));
    size_t *enc_n = (size_t *) calloc(n_reads, sizeof(size_t)), *dec_n = (size_t *) calloc(n_reads, sizeof(size_t));
    for (uint64_t r = 0; r < n_reads; ++r) { off[r] = sig_off[r] * 2; len[r] = (uint64_t) n_samples[r] * 2; }
    job_t j = { cfn, 2 /* SLOW5_COMPRESS_SVB_ZD */, (const uint8_t *) sig, off, len, enc, enc_n, n_reads, 0, 16 };
    *enc_seconds = run_pool(&j, threads);
    uint64_t total = 0;
    uint64_t *zoff = (uint64_t *) calloc(n_reads, 8), *zlen = (uint64_t *) malloc(8 * n_reads);
    for (uint64_t r = 0; r < n_reads; ++r) {
        if (!enc[r]) rc = -2;
        total += enc_n[r];
        zlen[r] = enc_n[r];
    }
    *svb_bytes = total;
    if (rc == 0) {
        /* decode from the per-record buffers the encoder returned: base 0 + absolute addresses */
        for (uint64_t r = 0; r < n_reads; ++r) zoff[r] = (uint64_t) (uintptr_t) enc[r];
        job_t d = { dfn, 2, (const uint8_t *) 0, zoff, zlen, dec, dec_n, n_reads, 0, 16 };
        *dec_seconds = run_pool(&d, threads);
        for (uint64_t r = 0; r < n_reads && rc == 0; ++r) {
            if (!dec[r] && len[r]) rc = -2;
            else if (len[r] && (dec_n[r] != len[r] || memcmp(dec[r], (const uint8_t *) sig + off[r], len[r]))) rc = -3;
        }
    }
    for (uint64_t r = 0; r < n_reads; ++r) { free(enc[r]); free(dec[r]); }
    free(off); free(len); free(enc); free(dec); free(enc_n); free(dec_n); free(zoff); free(zlen);
    if (h) dlclose(h);
    return rc;
}

/* ---- the full record path: what `slow5tools view` does to every record (src/view.c:35-57 depress_parse_rec_to_mem) ----
 * Per record, in the pool's worker threads:
 *     mem = malloc + copy of the stored record          (slow5_get_next_mem, slow5.c:3233-3281, minus the fread)
 *     slow5_decode(&mem, &bytes, &rec, from)            (slow5.h:658 -> slow5_rec_depress_parse, slow5.c:2580)
 *     slow5_encode(&out, &out_bytes, rec, to)           (slow5.h:660 -> slow5_press_init + slow5_rec_to_mem, slow5.c:4083)
 *     slow5_rec_free(rec); free(mem)
 * `from` / `to` are slow5_file_t handles opened by the reference itself ("w" mode files in tmpdir with
 * slow5_set_press(in/out methods)), so every byte of the work is the reference's own public API.  With ref_lib_path
 * NULL the oracle's restatement (blow5_oracle.c) stands in ("port").  The outputs are concatenated into `out` (the
 * file image, [u64 size][record] per record) after the clock stops; out may be NULL to time only.
 * Returns 0, -1 (dlopen / dlsym / open failed), -2 (a record failed), -4 (out too small; *out_bytes = bytes needed). */
typedef struct {
    int use_ref;
    int in_rec, in_sig, out_rec, out_sig;
    void *from, *to;
    int (*decode)(char **, size_t *, void **, void *);
    int (*encode)(char **, size_t *, void *, void *);
    void (*rec_free)(void *);
    const uint8_t *in;
    const uint64_t *rec_off;
    const uint32_t *rec_len;
    uint8_t **out;
    size_t *out_n;
    uint64_t n, next, grain;
    int failed;
} rjob_t;

static void *rworker(void *arg) {
    rjob_t *j = (rjob_t *) arg;
    for (;;) {
        uint64_t r0 = __sync_fetch_and_add(&j->next, j->grain);
        if (r0 >= j->n) break;
        uint64_t r1 = r0 + j->grain < j->n ? r0 + j->grain : j->n;
        for (uint64_t r = r0; r < r1; ++r) {
            j->out[r] = NULL;
            j->out_n[r] = 0;
            if (j->use_ref) {
                size_t bytes = j->rec_len[r];
                char *mem = (char *) malloc(bytes ? bytes : 1);
                if (!mem) { j->failed = 1; continue; }
                memcpy(mem, j->in + j->rec_off[r], bytes);
                void *rec = NULL;
                if (j->decode(&mem, &bytes, &rec, j->from) != 0) { free(mem); j->failed = 1; continue; }
                free(mem);
                char *o = NULL;
                size_t on = 0;
                if (j->encode(&o, &on, rec, j->to) != 0) { j->rec_free(rec); j->failed = 1; continue; }
                j->rec_free(rec);
                j->out[r] = (uint8_t *) o;
                j->out_n[r] = on;
            } else {
                if (orc_blow5_recode_record(j->in_rec, j->in_sig, j->out_rec, j->out_sig, j->in + j->rec_off[r], j->rec_len[r],
                                            &j->out[r], &j->out_n[r]) != 0)
                    j->failed = 1;
            }
        }
    }
    return NULL;
}

int refdrv_record_pass(const char *ref_lib_path, const char *tmpdir, int in_rec, int in_sig, int out_rec, int out_sig,
                       const uint8_t *in, const uint64_t *rec_off, const uint32_t *rec_len, uint64_t n, int threads,
                       uint8_t *out, uint64_t out_cap, uint64_t *out_bytes, uint64_t *out_off, double *seconds) {
    rjob_t j;
    memset(&j, 0, sizeof j);
    void *h = NULL;
    int (*closef)(void *) = NULL;
    char pa[4096], pb[4096];
    if (ref_lib_path) {
        h = dlopen(ref_lib_path, RTLD_NOW | RTLD_LOCAL);
        if (!h) return -1;
        void *(*openf)(const char *, const char *) = (void *(*)(const char *, const char *)) dlsym(h, "slow5_open");
        int (*setp)(void *, int, int) = (int (*)(void *, int, int)) dlsym(h, "slow5_set_press");
        void (*loglvl)(int) = (void (*)(int)) dlsym(h, "slow5_set_log_level");
        closef = (int (*)(void *)) dlsym(h, "slow5_close");
        j.decode = (int (*)(char **, size_t *, void **, void *)) dlsym(h, "slow5_decode");
        j.encode = (int (*)(char **, size_t *, void *, void *)) dlsym(h, "slow5_encode");
        j.rec_free = (void (*)(void *)) dlsym(h, "slow5_rec_free");
        if (!openf || !setp || !closef || !j.decode || !j.encode || !j.rec_free) { dlclose(h); return -1; }
        if (loglvl) loglvl(1 /* SLOW5_LOG_ERR */);
        snprintf(pa, sizeof pa, "%s/refdrv_from_%d.blow5", tmpdir, (int) getpid());
        snprintf(pb, sizeof pb, "%s/refdrv_to_%d.blow5", tmpdir, (int) getpid());
        /* `from` must look like a file that was READ (header->aux_meta == NULL without aux columns, methods taken from
         * the header): let the reference write an empty file with the input methods, then open that for reading */
        int (*hdrw)(void *) = (int (*)(void *)) dlsym(h, "slow5_hdr_write");
        void *tmp = openf(pa, "w");
        if (!hdrw || !tmp || setp(tmp, in_rec, in_sig) != 0 || hdrw(tmp) < 0) {
            if (tmp) closef(tmp);
            dlclose(h);
            return -1;
        }
        closef(tmp);
        j.from = openf(pa, "r");
        j.to = openf(pb, "w");
        if (!j.from || !j.to || setp(j.to, out_rec, out_sig) != 0) {
            if (j.from) closef(j.from);
            if (j.to) closef(j.to);
            dlclose(h);
            return -1;
        }
        j.use_ref = 1;
    }
    if (threads < 1) threads = 1;
    j.in_rec = in_rec; j.in_sig = in_sig; j.out_rec = out_rec; j.out_sig = out_sig;
    j.in = in; j.rec_off = rec_off; j.rec_len = rec_len;
    j.n = n; j.next = 0; j.grain = 16;
    j.out = (uint8_t **) calloc(n ? n : 1, sizeof(uint8_t *));
    j.out_n = (size_t *) calloc(n ? n : 1, sizeof(size_t));
    pthread_t *t = (pthread_t *) malloc(sizeof(pthread_t) * (size_t) threads);
    double t0 = now_s();
    for (int i = 0; i < threads; ++i) pthread_create(&t[i], NULL, rworker, &j);
    for (int i = 0; i < threads; ++i) pthread_join(t[i], NULL);
    *seconds = now_s() - t0;
    free(t);
    int rc = j.failed ? -2 : 0;
    uint64_t total = 0;
    for (uint64_t r = 0; r < n; ++r) total += j.out_n[r];
    *out_bytes = total;
    if (rc == 0 && out) {
        if (total > out_cap) rc = -4;
        else {
            uint64_t at = 0;
            for (uint64_t r = 0; r < n; ++r) {
                if (out_off) out_off[r] = at;
                memcpy(out + at, j.out[r], j.out_n[r]);
                at += j.out_n[r];
            }
            if (out_off) out_off[n] = at;
        }
    }
    for (uint64_t r = 0; r < n; ++r) free(j.out[r]);
    free(j.out); free(j.out_n);
    if (j.use_ref) {
        closef(j.from);
        closef(j.to);
        remove(pa);
        remove(pb);
        dlclose(h);
    }
    return rc;
}
