/* oracle/oracle.h -- TEST INFRASTRUCTURE. CPU restatement of the BLOW5 per-record codec hot path.
 *
 * This is the checker, not the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so.  Nothing under slow5tools_b200/ links,
 * imports or executes it.  Every function cites the reference file:line (relative to the reference
 * tree, slow5tools @ c114858 / slow5lib @ c13c4b8) whose behaviour it restates.
 *
 * Parity pinning: tests/test_oracle_svbzd.py checks this restatement against
 *   (1) the reference's own svb-zd test vectors (slow5lib/test/unit_test_press.c:118,149,179),
 *   (2) the known-answer table of SURVEY.md section 8c (tests/golden/svbzd_kat.json),
 *   (3) golden vectors produced by the compiled reference (tests/golden/, made by
 *       tests/golden/make_golden.py from oracle/_ref/libslow5_ref.so), and
 *   (4) when oracle/_ref/libslow5_ref.so is present, that library directly on random inputs.
 */
#ifndef S5B_ORACLE_H
#define S5B_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* streamvbyte.h:31-37 : worst-case svb bytes (keys + 4 B per value), header NOT included */
size_t orc_svb_max_compressedbytes(uint32_t n_values);

/* streamvbyte_zigzag.c:4-6,15-20 */
void orc_zigzag_delta_encode(const int32_t *in, uint32_t *out, size_t n, int32_t prev);
/* streamvbyte_zigzag.c:23-25,34-40 (int32 accumulator, truncating int16 store) */
void orc_zigzag_delta_decode(const uint32_t *in, int16_t *out, size_t n, int32_t prev);

/* streamvbyte_encode.c:31-115 (and the SSSE3 twin streamvbyte_x64_encode.c:6-56, same bytes):
 * returns bytes written (keys + data) */
size_t orc_svb_encode(const uint32_t *in, uint32_t n_values, uint8_t *out);
/* streamvbyte_decode.c:36-105 : returns bytes consumed (keys + data) */
size_t orc_svb_decode(const uint8_t *in, uint32_t *out, uint32_t n_values);

/* slow5_press.c:1082-1115 (ptr_compress_svb_zd) + :1062-1079 (ptr_compress_svb).
 * in: int16 samples, count_bytes = 2*N.  out must hold orc_svbzd_bound(N) bytes.
 * returns total bytes written: 4 (u32 N, LE) + keys + data. */
size_t orc_svbzd_bound(uint32_t n_samples);
size_t orc_svbzd_compress(const int16_t *in, size_t count_bytes, uint8_t *out);

/* slow5_press.c:1143-1173 (ptr_depress_svb_zd) + :1118-1140 (ptr_depress_svb).
 * returns 0 and sets *n_samples on success; -13 (SLOW5_ERR_PRESS) if the stream does not consume
 * exactly count_bytes-4 bytes (slow5_press.c:1130-1136); -2 (SLOW5_ERR_ARG) if count_bytes < 4 or
 * out_cap_samples is too small. */
int orc_svbzd_depress(const uint8_t *in, size_t count_bytes, int16_t *out, size_t out_cap_samples,
                      uint32_t *n_samples);

/* Batch forms used by the parity tests and by bench.py's cpu_baseline leg (one read after another,
 * single thread).  Offsets are in elements of the respective arrays (samples / bytes). */
void orc_svbzd_compress_batch(const int16_t *sig, const uint64_t *sig_off, const uint32_t *n_samples,
                              uint64_t n_reads, uint8_t *out, const uint64_t *out_off, uint32_t *out_len);
int  orc_svbzd_depress_batch(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                             uint64_t n_reads, int16_t *sig, const uint64_t *sig_off,
                             uint32_t *n_samples, int32_t *status);
/* size only (no output written) */
size_t orc_svbzd_size(const int16_t *in, uint32_t n_samples);

/* ---- ex-zd signal codec (oracle/exzd_oracle.c), slow5_press.c:1236-1848 ------------------------------------
 * Pinned by tests/test_oracle_exzd.py against the compiled reference and the reference's ex-zd BLOW5 golden. */
size_t orc_exzd_bound(uint64_t n_samples);
/* ptr_compress_ex_zd (:1778 -> _v0 :1721-1776); returns bytes written, 0 for an empty input (undefined in the reference) */
size_t orc_exzd_compress(const int16_t *in, size_t count_bytes, uint8_t *out);
/* ptr_depress_ex_zd (:1824 -> _v0 :1787-1822); 0 / -13 (SLOW5_ERR_PRESS) / -2 */
int orc_exzd_depress(const uint8_t *in, size_t count, int16_t *out, size_t out_cap_samples, uint64_t *n_samples);

/* ---- the per-record transcoding step of `slow5tools view` (oracle/blow5_oracle.c), slow5.c:2580-2950, :3928-4074 ----
 * one stored record in (no size prefix), one output record out INCLUDING its u64 size prefix (malloc'd).  Methods are
 * enum slow5_press_method values; records none / zlib, signals none / svb-zd / ex-zd. */
int orc_blow5_recode_record(int in_rec, int in_sig, int out_rec, int out_sig, const uint8_t *in, size_t in_len,
                            uint8_t **out, size_t *out_len);

/* ---- lossy degradation (oracle/qts_oracle.c), slow5_press.c:1965-2005 (slow5tools degrade) ----------------------------
 * bits in 1..16 (0: nothing happens); the int result is truncated into the int16 like the reference's store */
int orc_qts_round_sample(int sample, int bits);
void orc_qts_round(int16_t *samples, uint64_t n, int bits);

#ifdef __cplusplus
}
#endif
#endif
