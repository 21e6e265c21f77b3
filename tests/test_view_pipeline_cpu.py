"""The blow5 -> blow5 fast path of `view` (csrc/host/view_main.cpp: reader / transcoder / writer threads around
s5b_blow5_recode_host, record carry between chunks, positioned parallel reads, staging buffers made on a helper thread) WITHOUT a
device: the host sources are built against stubs of the GPU entry points (tools/dev/asan/gpu_stubs.cpp) in which the context
exists and the transcoder is the identity, and the input is an uncompressed file whose header claims zlib records -- which sends
it down the fast path.  The output must be the plain file at every chunk size (tools/dev/asan/pipeline.sh runs the same under
ThreadSanitizer / AddressSanitizer).  Test infrastructure only: nothing here is the product library."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "slow5tools_b200", "csrc", "host")


@pytest.fixture(scope="module")
def rig(tmp_path_factory):
    d = tmp_path_factory.mktemp("pipeline")
    s = open(os.path.join(ROOT, "tools", "dev", "asan", "gpu_stubs.cpp")).read().replace("/root/repo/include", os.path.join(ROOT, "include"))
    s = s.replace("int s5b_ctx_create(int, s5b_ctx_t **o) { *o = nullptr; return S5B_ERR_DEVICE; }",
                  "int s5b_ctx_create(int, s5b_ctx_t **o) { *o = (s5b_ctx_t *)malloc(8); return S5B_OK; }")
    s = s.replace("void s5b_ctx_destroy(s5b_ctx_t *) {}", "void s5b_ctx_destroy(s5b_ctx_t *c) { free(c); }")
    s = re.sub(r"int s5b_blow5_recode_host\([^)]*\) \{ return S5B_ERR_DEVICE; \}",
               "int s5b_blow5_recode_host(s5b_ctx_t *, int, int, int, int, const uint8_t *h_in, uint64_t, const uint64_t *off, "
               "const uint32_t *len, uint64_t n, uint8_t *h_out, uint64_t cap, uint64_t *out_bytes) { uint64_t at = 0; "
               "for (uint64_t i = 0; i < n; ++i) { if (at + 8 + len[i] <= cap) { uint64_t sz = len[i]; memcpy(h_out + at, &sz, 8); "
               "memcpy(h_out + at + 8, h_in + off[i], len[i]); } at += 8 + len[i]; } *out_bytes = at; "
               "return at > cap ? S5B_ERR_NOSPACE : S5B_OK; }", s)
    assert "memcpy(h_out" in s
    (d / "stubs.cpp").write_text(s)
    fl = ["-O1", "-std=c++11"]
    subprocess.check_call(["g++"] + fl + ["-fPIC", "-shared"] + [os.path.join(HOST, f) for f in
                          ("blow5_io.cpp", "s5b_file_api.cpp", "press_api.cpp", "index_main.cpp")] + [str(d / "stubs.cpp"), "-o", str(d / "libslow5b200.so")])
    subprocess.check_call(["g++"] + fl + [os.path.join(HOST, f) for f in ("view_main.cpp", "get_main.cpp", "merge_split_main.cpp", "degrade_main.cpp")] +
                          ["-o", str(d / "cli"), "-L", str(d), "-lslow5b200", "-lpthread", "-Wl,-rpath," + str(d)])
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_view
    reads, samples = 2600, 4096          # 21 MB: more than 8 MiB per chunk, so a chunk comes in through four positioned reads
    sig = np.random.default_rng(3).integers(0, 2000, reads * samples).astype(np.int16)
    plain = str(d / "plain.blow5")
    bench_view.write_blow5(plain, sig, reads, samples)
    b = bytearray(open(plain, "rb").read())
    b[9] = 1                              # record method byte: "zlib" (the stub codec is the identity)
    fake = str(d / "fake_zlib.blow5")
    open(fake, "wb").write(b)
    return str(d / "cli"), plain, fake, d


@pytest.mark.parametrize("chunk_kb", [None, 4, 700])
def test_fast_path_pipeline_reproduces_the_file(rig, chunk_kb):
    cli, plain, fake, d = rig
    out = str(d / ("out_%s.blow5" % chunk_kb))
    env = dict(os.environ)
    env.pop("S5B_VIEW_CHUNK_KB", None)
    if chunk_kb:
        env["S5B_VIEW_CHUNK_KB"] = str(chunk_kb)   # 4 KiB: every record is larger than a chunk (the reader grows its buffer)
    r = subprocess.run([cli, "view", fake, "-c", "none", "-s", "none", "-o", out], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=env, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    assert open(out, "rb").read() == open(plain, "rb").read()


def test_fast_path_reports_truncated_input(rig):
    cli, plain, fake, d = rig
    cut = str(d / "cut.blow5")
    open(cut, "wb").write(open(fake, "rb").read()[:-3000])
    r = subprocess.run([cli, "view", cut, "-c", "none", "-s", "none", "-o", str(d / "cut_out.blow5")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 1 and b"truncated" in r.stderr
