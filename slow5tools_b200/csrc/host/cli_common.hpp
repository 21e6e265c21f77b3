// cli_common.hpp -- what the sub-commands of slow5tools-b200 share: the general conversion loop of view_main.cpp.
#pragma once
#include <cstdio>
#include <functional>
#include <vector>

#include "../../../include/slow5b200.h"
#include "blow5_io.hpp"

// Optional per-record steps of convert_records for the callers that do more than convert (src/merge.c:43-70,
// src/split.c:81-110: the same decode -> modify -> re-encode worker with a different "modify" and destination).
struct ConvertHooks {
    // header the OUTPUT records follow (their auxiliary columns); nullptr: the input's
    const s5b::Header *hdr_out = nullptr;
    // called once per parsed record, in input order, before it is re-encoded: may change read_group and point aux_bytes /
    // aux_nbytes at a new auxiliary section kept in aux_store; false = error already reported
    std::function<bool(size_t i, s5b::Record &rec, std::vector<uint8_t> &aux_store)> transform;
    // destination of record i of the batch; nullptr: fout
    std::function<FILE *(size_t i)> route;
    // demultiplexing (src/demux.c:562-590): every file record i goes to -- none (the record is dropped), one or several; takes
    // precedence over route
    std::function<const std::vector<FILE *> *(size_t i)> route_many;
    // degrade (src/degrade.c:255): > 0 rounds this many low bits of every sample away before the record is re-encoded
    int qts_bits = 0;
};

int convert_records(const s5b::Header &hdr, s5b::Fmt fmt_in, const std::function<int(std::vector<uint8_t> &)> &next, FILE *fout,
                    s5b_ctx_t *gpu, s5b::Fmt fmt_out, int rec_out, int sig_out, long batch, int threads,
                    const ConvertHooks *hooks = nullptr);

// blow5 -> blow5 with whole batches resident on the device (view_main.cpp: pinned chunk pipeline around s5b_blow5_recode_host):
// converts the records of `rd` (positioned behind its header) to (rec_out, sig_out) and appends them to fout.  The context's
// auxiliary layout / read-group table (s5b_ctx_set_aux_layout, s5b_ctx_set_rg_map) apply.  0 = ok.
int blow5_fast_convert(s5b::Reader &rd, FILE *fout, s5b_ctx_t *gpu, int rec_out, int sig_out);
