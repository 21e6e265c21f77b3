#!/bin/bash
# The blow5 -> blow5 fast path of `view` (reader / transcoder / writer threads, chunk carry, positioned parallel reads, staging
# buffers made on the side) under ThreadSanitizer and AddressSanitizer without a device: the GPU stubs are patched so that the
# context exists and s5b_blow5_recode_host is the identity (records in -> [u64 size][record] image out), and the input is an
# uncompressed file whose header claims zlib records, which sends it down the fast path.  The output must equal the plain file.
# usage: bash tools/dev/asan/pipeline.sh     (needs numpy for the 40 MB test file; prints one line per run)
set -e
R=$(cd "$(dirname "$0")/../../.." && pwd); H=$R/slow5tools_b200/csrc/host; T=$(mktemp -d); cd $T
python3 - "$R" <<'PY'
import re, sys
R = sys.argv[1]
s = open(R + "/tools/dev/asan/gpu_stubs.cpp").read().replace("/root/repo/include", R + "/include")
s = s.replace("int s5b_ctx_create(int, s5b_ctx_t **o) { *o = nullptr; return S5B_ERR_DEVICE; }",
              "int s5b_ctx_create(int, s5b_ctx_t **o) { *o = (s5b_ctx_t *)malloc(8); return S5B_OK; }")
s = s.replace("void s5b_ctx_destroy(s5b_ctx_t *) {}", "void s5b_ctx_destroy(s5b_ctx_t *c) { free(c); }")
s = re.sub(r"int s5b_blow5_recode_host\([^)]*\) \{ return S5B_ERR_DEVICE; \}",
           "int s5b_blow5_recode_host(s5b_ctx_t *, int, int, int, int, const uint8_t *h_in, uint64_t, const uint64_t *off, const uint32_t *len, "
           "uint64_t n, uint8_t *h_out, uint64_t cap, uint64_t *out_bytes) { uint64_t at = 0; for (uint64_t i = 0; i < n; ++i) { "
           "if (at + 8 + len[i] <= cap) { uint64_t sz = len[i]; memcpy(h_out + at, &sz, 8); memcpy(h_out + at + 8, h_in + off[i], len[i]); } "
           "at += 8 + len[i]; } *out_bytes = at; return at > cap ? S5B_ERR_NOSPACE : S5B_OK; }", s)
assert "memcpy(h_out" in s
open("stubs.cpp", "w").write(s)
import numpy as np
sys.path.insert(0, R); sys.path.insert(0, R + "/tools")
import bench_view
sig = np.random.default_rng(3).integers(0, 2000, 5000 * 4096).astype(np.int16)
bench_view.write_blow5("plain.blow5", sig, 5000, 4096)
b = bytearray(open("plain.blow5", "rb").read()); b[9] = 1
open("fake_zlib.blow5", "wb").write(b)
PY
for SAN in thread address; do
  FL="-g -O1 -std=c++11 -fsanitize=$SAN -fno-omit-frame-pointer"; mkdir -p $SAN
  g++ $FL -fPIC -shared $H/blow5_io.cpp $H/s5b_file_api.cpp $H/press_api.cpp $H/index_main.cpp stubs.cpp -o $SAN/libslow5b200.so
  g++ $FL $H/view_main.cpp $H/get_main.cpp $H/merge_split_main.cpp $H/degrade_main.cpp -o $SAN/cli -L $SAN -lslow5b200 -lpthread -Wl,-rpath,$T/$SAN
  for kb in default 4 700; do
    if [ $kb = default ]; then unset S5B_VIEW_CHUNK_KB; else export S5B_VIEW_CHUNK_KB=$kb; fi
    rm -f out.blow5
    ASAN_OPTIONS=detect_leaks=0 ./$SAN/cli view fake_zlib.blow5 -c none -s none -o out.blow5 > o.out 2> o.err && rc=0 || rc=$?
    cmp -s out.blow5 plain.blow5 && same=yes || same=no
    echo "$SAN chunk_kb=$kb rc=$rc identical=$same reports=$(grep -c 'ERROR: AddressSanitizer\|WARNING: ThreadSanitizer' o.err || true)"
  done
done
cd /; rm -rf $T
