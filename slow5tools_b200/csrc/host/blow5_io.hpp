// blow5_io.hpp -- host-side BLOW5 / SLOW5 file framing, header and record (de)serialisation.
//
// The part of slow5lib's file+record layer (slow5lib/src/slow5.c) that the codec path sits inside and that
// SURVEY 8(a) rows a16-a18 say stays on the host: header parse/emit (slow5.c:638-881, :948-1306), raw record
// fetch (:3206-3281), binary record parse (:2811-2950, :3088-3166), record packing (:3928-4074) and the
// SLOW5 ASCII record form (:3824-3926, :4479-4619; slow5_misc.c:379-406).  Written from the format, not
// transcribed; the codec calls themselves are NOT here -- they go through the batch C-ABI to the GPU.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

namespace s5b {

enum Fmt { FMT_UNKNOWN = 0, FMT_ASCII = 1, FMT_BINARY = 2 };

// enum slow5_press_method (slow5_press.h:61-67)
enum Press { PRESS_NONE = 0, PRESS_ZLIB = 1, PRESS_SVB_ZD = 2, PRESS_ZSTD = 3, PRESS_EX_ZD = 4, PRESS_BAD = 255 };

// file byte <-> library enum (slow5_press.c:58-161)
int record_press_from_byte(uint8_t b);
int signal_press_from_byte(uint8_t b);
uint8_t record_press_to_byte(int m);
uint8_t signal_press_to_byte(int m);
int press_from_name(const char *name);  // none / zlib / svb-zd / zstd / ex-zd (src/misc.c:251-265), PRESS_BAD otherwise

// enum slow5_aux_type order (slow5.h:104-131)
enum AuxType {
    AUX_INT8 = 0, AUX_INT16, AUX_INT32, AUX_INT64, AUX_UINT8, AUX_UINT16, AUX_UINT32, AUX_UINT64, AUX_FLOAT, AUX_DOUBLE,
    AUX_CHAR, AUX_ENUM,
    AUX_INT8_ARRAY, AUX_INT16_ARRAY, AUX_INT32_ARRAY, AUX_INT64_ARRAY, AUX_UINT8_ARRAY, AUX_UINT16_ARRAY,
    AUX_UINT32_ARRAY, AUX_UINT64_ARRAY, AUX_FLOAT_ARRAY, AUX_DOUBLE_ARRAY, AUX_STRING, AUX_ENUM_ARRAY,
    AUX_UNKNOWN = 255
};

struct AuxField {
    std::string name;
    std::string type_str;  // as written in the header, e.g. "uint8_t", "char*", "enum{a,b}"
    int type = AUX_UNKNOWN;
    uint8_t size = 0;      // bytes of the primitive
    bool is_array() const { return type >= AUX_INT8_ARRAY; }
};

struct Header {
    uint8_t version[3] = {0, 1, 0};
    uint32_t num_read_groups = 1;
    int record_method = PRESS_NONE;  // of the file this header was read from
    int signal_method = PRESS_NONE;
    std::vector<std::pair<std::string, std::vector<std::string>>> attrs;  // key -> value per read group ("" = missing)
    std::vector<AuxField> aux;
};

struct Record {
    std::string read_id;
    uint32_t read_group = 0;
    double digitisation = 0, offset = 0, range = 0, sampling_rate = 0;
    uint64_t len_raw_signal = 0;     // samples
    std::vector<int16_t> raw_signal; // decoded samples (empty while the signal is still compressed)
    const uint8_t *sig_bytes = nullptr;  // view into the packed record: the stored signal bytes
    uint64_t sig_nbytes = 0;
    const uint8_t *aux_bytes = nullptr;  // view into the packed record: binary aux section
    uint64_t aux_nbytes = 0;
};

struct Reader {
    FILE *fp = nullptr;
    Fmt fmt = FMT_UNKNOWN;
    Header hdr;
    std::string path;
    std::string err;
};

// opens and parses the header; returns false with r.err set on failure
bool reader_open(Reader &r, const char *path, Fmt fmt);
void reader_close(Reader &r);
// Binary: next packed record WITHOUT the u64 size prefix (slow5_get_next_mem, slow5.c:3233-3281).
// Returns 1 record read, 0 clean EOF ("5WOLB" seen), -1 error (r.err).  ASCII: one line without '\n'.
int reader_next_mem(Reader &r, std::vector<uint8_t> &mem);

// header emit (slow5_hdr_to_mem, slow5.c:948-1157).  For binary the version is bumped 0.1.0 -> 0.2.0 when the
// requested compression needs it (slow5.c:4729-4746).
std::string header_to_mem(const Header &h, Fmt fmt, int record_method, int signal_method);

// binary record parse: fills the fixed fields and the sig/aux views (signal still as stored).
// Returns false when the record is inconsistent (slow5.c:2937-2949).
bool record_parse_binary(const uint8_t *mem, uint64_t n, const Header &h, int signal_method, Record &rec, std::string &err);
// ASCII record parse (one SLOW5 line); aux columns are converted to the binary aux form in aux_store.
// defer_signal: leave the raw_signal column unparsed and point rec.sig_bytes / sig_nbytes at its characters inside `line`
bool record_parse_ascii(const char *line, uint64_t n, const Header &h, Record &rec, std::vector<uint8_t> &aux_store,
                        std::string &err, bool defer_signal = false);

// SLOW5 ASCII line for a record whose raw_signal is decoded (slow5.c:3824-3926), including '\n'
// sig_text (optional): the already formatted raw_signal column (s5b_signal_to_ascii_batch_host); else rec.raw_signal is printed
void record_to_ascii(const Record &rec, const Header &h, std::string &out, const char *sig_text = nullptr, size_t sig_text_len = 0);
// packed binary record WITHOUT record compression and WITHOUT the size prefix (slow5.c:3928-4044);
// `signal` = the bytes to store (raw int16 or svb-zd stream), `signal_is_compressed` selects the meaning of the
// len_raw_signal field (sample count vs byte count, slow5.c:3983-3987).  *signal_at receives the offset of
// the signal bytes inside the record.
void record_to_binary(const Record &rec, const uint8_t *signal, uint64_t signal_nbytes, bool signal_is_compressed,
                      std::vector<uint8_t> &out, uint64_t *signal_at);

// auxiliary section of a record re-laid for another header (merge: union of the inputs' columns); see blow5_io.cpp
bool aux_null_value(int type, uint8_t *dst);
bool aux_relayout(const uint8_t *aux, uint64_t n, const std::vector<AuxField> &in, const std::vector<AuxField> &out,
                  const std::vector<int> &src_of_out, std::vector<uint8_t> &dst);

std::string double_to_str(double x);  // slow5_misc.c:379-406

// FILE.idx for a SLOW5 / BLOW5 file (index_main.cpp; slow5_idx_create, slow5.c:4138-4150); 0 = written, errors are reported on stderr
int index_build_file(const char *path);

Fmt fmt_from_path(const char *path);   // by extension (src/misc.c:178-216)
Fmt fmt_from_name(const char *name);   // "slow5" / "blow5"

}  // namespace s5b
