"""The drop-in boundary: the slow5lib names and struct layouts a caller of the per-record codec path sees.

* the reference's own example program slow5lib/examples/adv/sequential_read_pthreads.c, UNCHANGED, compiled against
  include/compat (our <slow5/slow5.h>) and linked with libslow5b200.so prints what the reference build prints;
* libslow5b200_compat.so exports the slow5_* symbols themselves;
* the public struct layouts match the reference's headers field by field (checked with offsetof on both sides);
* the stateful press twins (slow5_press_init / slow5_ptr_compress / slow5_compress_footer_next ...) load and validate.
CPU only.  The pieces that need the reference tree (example source, headers) skip without it."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFTREE = "/root/reference"
LIBDIR = os.path.join(ROOT, "slow5tools_b200")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libslow5_ref.so")
have_tree = pytest.mark.skipif(not os.path.exists(os.path.join(REFTREE, "slow5lib", "examples", "adv", "sequential_read_pthreads.c")),
                               reason="reference tree not present")


def _cc(args, cwd=None):
    subprocess.check_call(["gcc"] + args, cwd=cwd)


@have_tree
def test_reference_example_builds_unchanged_and_prints_the_same(tmp_path):
    src = os.path.join(REFTREE, "slow5lib", "examples", "adv", "sequential_read_pthreads.c")
    os.makedirs(tmp_path / "examples")
    shutil.copy(os.path.join(REFTREE, "slow5lib", "examples", "example.slow5"), tmp_path / "examples" / "example.slow5")
    ours = str(tmp_path / "ours")
    _cc(["-O1", "-Wall", "-I", os.path.join(ROOT, "include", "compat"), src, "-o", ours, "-L", LIBDIR, "-lslow5b200", "-lpthread",
         "-Wl,-rpath," + LIBDIR])
    a = subprocess.run([ours], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert a.returncode == 0, a.stderr.decode()
    mine = sorted(a.stderr.decode().splitlines())
    assert any("Successfully decoded the read" in ln for ln in mine) and any("Read 5 raw records" in ln for ln in mine)
    if os.path.exists(REF_SO):
        theirs = str(tmp_path / "theirs")
        _cc(["-O1", "-I", os.path.join(REFTREE, "slow5lib", "include"), src, "-o", theirs, REF_SO, "-lpthread", "-lz",
             "-Wl,-rpath," + os.path.dirname(REF_SO)])
        b = subprocess.run([theirs], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
        assert b.returncode == 0
        assert mine == sorted(b.stderr.decode().splitlines())   # (threads print in any order)


def test_compat_library_exports_the_slow5_names():
    L = C.CDLL(os.path.join(LIBDIR, "libslow5b200_compat.so"))
    for name in ("slow5_open", "slow5_close", "slow5_get_next_mem", "slow5_get_next_bytes", "slow5_decode", "slow5_encode",
                 "slow5_write_bytes", "slow5_rec_free", "slow5_set_press", "slow5_hdr_write", "slow5_init_mt", "slow5_init_batch",
                 "slow5_get_next_batch", "slow5_encode_batch", "slow5_write_batch", "slow5_free_batch", "slow5_free_mt",
                 "slow5_press_init", "__slow5_press_init", "slow5_press_free", "__slow5_press_free", "slow5_ptr_compress",
                 "slow5_ptr_depress", "slow5_ptr_compress_solo", "slow5_ptr_depress_solo", "slow5_compress_footer_next",
                 "slow5_errno_location", "slow5_get_next", "slow5_get", "slow5_write", "slow5_get_batch", "slow5_idx_load",
                 "slow5_hdr_get", "slow5_aux_get_int8", "slow5_aux_get_int16", "slow5_aux_get_int32", "slow5_aux_get_int64",
                 "slow5_aux_get_uint8", "slow5_aux_get_uint16", "slow5_aux_get_uint32", "slow5_aux_get_uint64", "slow5_aux_get_float",
                 "slow5_aux_get_double", "slow5_aux_get_char", "slow5_aux_get_enum", "slow5_aux_get_string",
                 "slow5_aux_get_int8_array", "slow5_aux_get_int16_array", "slow5_aux_get_int32_array", "slow5_aux_get_int64_array",
                 "slow5_aux_get_uint8_array", "slow5_aux_get_uint16_array", "slow5_aux_get_uint32_array",
                 "slow5_aux_get_uint64_array", "slow5_aux_get_float_array", "slow5_aux_get_double_array", "slow5_aux_get_enum_array"):
        assert hasattr(L, name), name
    # a round trip through the names (method NONE needs no GPU): open a reference fixture, read the raw records
    fix = os.path.join(ROOT, "tests", "golden", "fixtures", "exp_1_lossless_zlib_svb_v0.2.0.blow5")
    L.slow5_open.restype = C.c_void_p
    L.slow5_open.argtypes = [C.c_char_p, C.c_char_p]
    L.slow5_get_next_bytes.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p]
    L.slow5_close.argtypes = [C.c_void_p]
    L.slow5_errno_location.restype = C.POINTER(C.c_int)
    sp = L.slow5_open(fix.encode(), b"r")
    assert sp
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    n = 0
    while True:
        mem, nb = C.c_void_p(), C.c_size_t()
        if L.slow5_get_next_bytes(C.byref(mem), C.byref(nb), sp) < 0:
            assert L.slow5_errno_location()[0] == -1          # SLOW5_ERR_EOF
            break
        assert nb.value > 0
        libc.free(mem)
        n += 1
    assert n >= 1
    assert L.slow5_close(sp) == 0


@have_tree
def test_public_struct_layouts_match_the_reference_headers(tmp_path):
    """offsetof / sizeof of every field a caller may touch, printed by one program compiled against the reference's
    headers and one against ours"""
    fields_rec = ["read_id_len", "read_id", "read_group", "digitisation", "offset", "range", "sampling_rate", "len_raw_signal",
                  "raw_signal"]
    fields_file = ["fp", "format", "compress", "header", "index", "meta.pathname", "meta.fd", "meta.start_rec_offset",
                   "meta.fread_buffer", "meta.mode"]
    fields_batch = ["n_rec", "capacity_rec", "mem_records", "mem_bytes", "slow5_rec", "rid"]
    body = ["#include <stdio.h>", "#include <stddef.h>", "#include <slow5/slow5.h>", "#include <slow5/slow5_mt.h>",
            "#include <slow5/slow5_press.h>", "int main(void){"]
    for f in fields_rec:
        body.append('printf("rec.%s %%zu\\n", offsetof(slow5_rec_t, %s));' % (f, f))
    for f in fields_file:
        body.append('printf("file.%s %%zu\\n", offsetof(slow5_file_t, %s));' % (f, f))
    for f in fields_batch:
        body.append('printf("batch.%s %%zu\\n", offsetof(slow5_batch_t, %s));' % (f, f))
    body += ['printf("mt.sf %zu\\n", offsetof(slow5_mt_t, sf));', 'printf("mt.num_thread %zu\\n", offsetof(slow5_mt_t, num_thread));',
             'printf("press.record_press %zu\\n", offsetof(slow5_press_t, record_press));',
             'printf("press.signal_press %zu\\n", offsetof(slow5_press_t, signal_press));',
             'printf("__press.method %zu\\n", offsetof(struct __slow5_press, method));',
             'printf("__press.stream %zu\\n", offsetof(struct __slow5_press, stream));',
             'printf("method_t %zu %zu\\n", sizeof(slow5_press_method_t), offsetof(slow5_press_method_t, signal_method));',
             'printf("hdr.version %zu hdr.num_read_groups %zu\\n", offsetof(struct slow5_hdr, version), offsetof(struct slow5_hdr, num_read_groups));',
             "return 0;}"]
    src = tmp_path / "layout.c"
    ours_body = "\n".join(body).replace("struct slow5_hdr", "struct s5b_hdr")
    src.write_text("\n".join(body))
    (tmp_path / "layout_ours.c").write_text(ours_body)
    _cc(["-DSLOW5_ENABLE_MT", "-I", os.path.join(REFTREE, "slow5lib", "include"), str(src), "-o", str(tmp_path / "ref")])
    _cc(["-I", os.path.join(ROOT, "include", "compat"), str(tmp_path / "layout_ours.c"), "-o", str(tmp_path / "ours")])
    ref = subprocess.check_output([str(tmp_path / "ref")]).decode()
    ours = subprocess.check_output([str(tmp_path / "ours")]).decode()
    assert ref == ours, "\n" + ref + "\n---\n" + ours


def test_stateful_press_objects():
    L = C.CDLL(os.path.join(LIBDIR, "libslow5b200.so"))

    class Method(C.Structure):
        _fields_ = [("record_method", C.c_int), ("signal_method", C.c_int)]

    class Inner(C.Structure):
        _fields_ = [("method", C.c_int), ("stream", C.c_void_p)]

    class Press(C.Structure):
        _fields_ = [("record_press", C.POINTER(Inner)), ("signal_press", C.POINTER(Inner))]

    L.s5b_press_init.restype = C.POINTER(Press)
    L.s5b_press_init.argtypes = [Method]
    L.s5b_press_free.argtypes = [C.POINTER(Press)]
    L.s5b_ptr_compress.restype = C.c_void_p
    L.s5b_ptr_compress.argtypes = [C.POINTER(Inner), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.s5b_ptr_depress.restype = C.c_void_p
    L.s5b_ptr_depress.argtypes = [C.POINTER(Inner), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.s5b_compress_footer_next.argtypes = [C.POINTER(Inner)]
    p = L.s5b_press_init(Method(1, 2))        # zlib records, svb-zd signals: view's default (src/view.c:43)
    assert p and p.contents.record_press.contents.method == 1 and p.contents.signal_press.contents.method == 2
    L.s5b_compress_footer_next(p.contents.record_press)
    L.s5b_press_free(p)
    assert not L.s5b_press_init(Method(9, 0))  # unknown method (slow5_press.c:282-295: SLOW5_ERR_ARG)
    # method NONE is a copy and needs no device (slow5_press.c:340-349)
    q = L.s5b_press_init(Method(0, 0))
    n = C.c_size_t()
    data = b"0123456789"
    out = L.s5b_ptr_compress(q.contents.record_press, data, len(data), C.byref(n))
    assert out and n.value == len(data) and C.string_at(out, n.value) == data
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(out)
    assert not L.s5b_ptr_depress(None, data, len(data), C.byref(n)) and n.value == 0
    L.s5b_press_free(q)


GET_BATCH_C = r"""
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <slow5/slow5.h>
#include <slow5/slow5_mt.h>
/* prints "id len checksum" for every record: first sequentially (slow5_get_next_batch), then by read id in reverse order
 * through slow5_get_batch -- written against the slow5lib names only */
static unsigned long sum(const slow5_rec_t *r) {
    unsigned long s = 0;
    for (uint64_t i = 0; i < r->len_raw_signal; ++i) s = s * 31 + (unsigned short) r->raw_signal[i];
    return s;
}
int main(int argc, char **argv) {
    slow5_file_t *sp = slow5_open(argv[1], "r");
    if (!sp) return 2;
    slow5_mt_t *mt = slow5_init_mt(4, sp);
    slow5_batch_t *b = slow5_init_batch(64);
    int n = slow5_get_next_batch(mt, b, 64);
    if (n <= 0) return 3;
    char **ids = (char **) malloc(sizeof(char *) * n);
    for (int i = 0; i < n; ++i) {
        printf("seq %s %lu %lu\n", b->slow5_rec[i]->read_id, (unsigned long) b->slow5_rec[i]->len_raw_signal, sum(b->slow5_rec[i]));
        ids[n - 1 - i] = strdup(b->slow5_rec[i]->read_id);
    }
    if (slow5_idx_load(sp) < 0) return 4;
    slow5_batch_t *g = slow5_init_batch(64);
    if (slow5_get_batch(mt, g, ids, n) != n) return 5;
    for (int i = n - 1; i >= 0; --i)
        printf("seq %s %lu %lu\n", g->slow5_rec[i]->read_id, (unsigned long) g->slow5_rec[i]->len_raw_signal, sum(g->slow5_rec[i]));
    char *missing[1] = {(char *) "no-such-read"};
    printf("missing %d\n", slow5_get_batch(mt, g, missing, 1) < 0);
    slow5_free_batch(g);
    slow5_free_batch(b);
    slow5_free_mt(mt);
    slow5_close(sp);
    return 0;
}
"""


def _get_batch_program(tmp_path):
    src = tmp_path / "getb.c"
    src.write_text(GET_BATCH_C)
    exe = str(tmp_path / "getb")
    _cc(["-O1", "-I", os.path.join(ROOT, "include", "compat"), str(src), "-o", exe, "-L", LIBDIR, "-lslow5b200", "-Wl,-rpath," + LIBDIR])
    return exe


def _run_get_batch(tmp_path, fixture):
    exe = _get_batch_program(tmp_path)
    f = tmp_path / os.path.basename(fixture)
    shutil.copy(fixture, f)
    cli = os.path.join(LIBDIR, "bin", "slow5tools-b200")
    subprocess.check_call([cli, "index", str(f)], stderr=subprocess.DEVNULL)
    out = subprocess.check_output([exe, str(f)]).decode().splitlines()
    n = (len(out) - 1) // 2
    assert n >= 1 and out[-1] == "missing 1"
    assert out[:n] == out[n:2 * n]          # random access by id returns what the sequential read returned


def test_get_batch_by_read_id_uncompressed(tmp_path):
    _run_get_batch(tmp_path, os.path.join(ROOT, "tests", "golden", "fixtures", "exp_1_lossless.blow5"))


@pytest.mark.gpu
def test_get_batch_by_read_id_compressed(tmp_path):
    _run_get_batch(tmp_path, os.path.join(ROOT, "tests", "golden", "fixtures", "zlib_svb-zd_multi_rg_v0.2.0.blow5"))


# ---- auxiliary field accessors, header attributes, error codes: ours next to the compiled reference (CPU: uncompressed file) ----
AUX_PROG = r"""
#include <stdio.h>
#include <stdlib.h>
#include <inttypes.h>
#include <slow5/slow5.h>
int main(int argc, char **argv) {
    slow5_file_t *sp = slow5_open(argv[1], "r");
    if (!sp) return 2;
    printf("hdr run_id=%s asic_id=%s nosuch=%s rg9=%s\n", slow5_hdr_get("run_id", 0, sp->header), slow5_hdr_get("asic_id", 0, sp->header),
           slow5_hdr_get("no_such_attribute", 0, sp->header) ? "set" : "NULL", slow5_hdr_get("run_id", 9, sp->header) ? "set" : "NULL");
    char *mem = NULL; size_t bytes = 0; slow5_rec_t *rec = NULL; int n = 0;
    while (slow5_get_next_bytes(&mem, &bytes, sp) == 0) {
        if (slow5_decode(&mem, &bytes, &rec, sp) != 0) return 3;
        free(mem);
        int e1, e2, e3, e4, e5, e6, e7, e8; uint64_t len = 77, l2 = 77;
        char *ch = slow5_aux_get_string(rec, "channel_number", &len, &e1);
        double mb = slow5_aux_get_double(rec, "median_before", &e2);
        int32_t rn = slow5_aux_get_int32(rec, "read_number", &e3);
        uint8_t mux = slow5_aux_get_uint8(rec, "start_mux", &e4);
        uint64_t st = slow5_aux_get_uint64(rec, "start_time", &e5);
        int64_t wrong = slow5_aux_get_int64(rec, "read_number", &e6);      /* int32 field asked for as int64 */
        uint16_t none = slow5_aux_get_uint16(rec, "no_such_field", &e7);
        int8_t *arr = slow5_aux_get_int8_array(rec, "channel_number", &l2, &e8);   /* string asked for as int8 array */
        printf("%s ch=%s/%" PRIu64 "/%d mb=%.6f/%d rn=%" PRId32 "/%d mux=%u/%d st=%" PRIu64 "/%d wrong=%" PRId64 "/%d none=%u/%d arr=%s/%" PRIu64 "/%d\n",
               rec->read_id, ch, len, e1, mb, e2, rn, e3, (unsigned)mux, e4, st, e5, wrong, e6, (unsigned)none, e7, arr ? "set" : "NULL", l2, e8);
        ++n;
    }
    slow5_rec_free(rec);
    slow5_close(sp);
    printf("records %d\n", n);
    return 0;
}
"""


@have_tree
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
def test_aux_accessors_and_header_attributes_match_the_reference(tmp_path):
    """slow5_aux_get_* / slow5_hdr_get (slow5.h:396, :469-508) on the reference's own fixture with five auxiliary columns: one
    program, compiled against the reference's headers + library and against include/compat + libslow5b200.so, prints the same --
    values, lengths, and the error codes for a wrong type, a missing field and a missing attribute"""
    fixture = os.path.join(ROOT, "tests", "golden", "fixtures", "exp_1_lossless.blow5")   # none / none: no codec, no device
    src = tmp_path / "aux.c"
    src.write_text(AUX_PROG)
    inc = os.path.join(REFTREE, "slow5lib", "include")
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "theirs")
    _cc(["-O1", "-I", os.path.join(ROOT, "include", "compat"), str(src), "-o", ours, "-L", LIBDIR, "-lslow5b200", "-Wl,-rpath," + LIBDIR])
    refdir = os.path.dirname(REF_SO)
    _cc(["-O1", "-I", inc, str(src), "-o", theirs, "-L", refdir, "-l:libslow5_ref.so", "-Wl,-rpath," + refdir, "-lm", "-lz"])
    a = subprocess.run([ours, fixture], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    b = subprocess.run([theirs, fixture], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr.decode(), b.stderr.decode())
    assert a.stdout == b.stdout
    out = a.stdout.decode()
    assert "records 1" in out or "records" in out
    assert "/-17" in out and "/-12" in out and "nosuch=NULL" in out   # SLOW5_ERR_TYPE, SLOW5_ERR_NOFLD


def test_error_codes_carry_the_reference_values():
    """slow5_defs.h:137-154"""
    text = open(os.path.join(ROOT, "include", "slow5b200_file.h")).read() + open(os.path.join(ROOT, "include", "slow5b200.h")).read()
    want = {"S5B_ERR_EOF": -1, "S5B_ERR_ARG": -2, "S5B_ERR_RECPARSE": -4, "S5B_ERR_IO": -5, "S5B_ERR_NOIDX": -6, "S5B_ERR_NOTFOUND": -7,
            "S5B_ERR_MEM": -10, "S5B_ERR_NOAUX": -11,
            "S5B_ERR_NOFLD": -12, "S5B_ERR_PRESS": -13, "S5B_ERR_TYPE": -17}
    import re
    for name, value in want.items():
        m = re.search(r"#define\s+%s\s+\((-?\d+)\)" % name, text)
        assert m and int(m.group(1)) == value, name
    defs = os.path.join(REFTREE, "slow5lib", "include", "slow5", "slow5_defs.h")
    if os.path.exists(defs):
        ref = open(defs).read()
        for name, value in want.items():
            m = re.search(r"#define\s+%s\s+\((-?\d+)\)" % name.replace("S5B_", "SLOW5_"), ref)
            assert m and int(m.group(1)) == value, name


GET_PROG = r"""
#include <stdio.h>
#include <stdlib.h>
#include <inttypes.h>
#include <slow5/slow5.h>
int main(int argc, char **argv) {
    slow5_file_t *sp = slow5_open(argv[1], "r");
    if (!sp) return 2;
    slow5_rec_t *rec = NULL;
    int rc, n = 0;
    rc = slow5_get(argv[2], &rec, sp);
    printf("get before idx_load -> %d\n", rc);
    while ((rc = slow5_get_next(&rec, sp)) >= 0) {
        printf("next %s %" PRIu64 " %d\n", rec->read_id, rec->len_raw_signal, (int)rec->raw_signal[rec->len_raw_signal - 1]);
        ++n;
    }
    printf("end %d after %d\n", rc, n);
    if (slow5_idx_load(sp) < 0) return 3;
    for (int i = 2; i < argc; ++i) {
        rc = slow5_get(argv[i], &rec, sp);
        if (rc == 0) printf("get %s %" PRIu64 " %d\n", rec->read_id, rec->len_raw_signal, (int)rec->raw_signal[0]);
        else printf("get %s -> %d\n", argv[i], rc);
    }
    slow5_rec_free(rec);
    slow5_close(sp);
    return 0;
}
"""


@have_tree
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
def test_single_record_calls_match_the_reference(tmp_path):
    """slow5_get_next / slow5_get (slow5.h:423, :440) on an uncompressed BLOW5 made from the reference's example.slow5 (5 reads):
    the same program against both libraries prints the same, including SLOW5_ERR_NOIDX, SLOW5_ERR_NOTFOUND and SLOW5_ERR_EOF;
    slow5_write writes the records back byte for byte"""
    cli = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
    text = os.path.join(REFTREE, "slow5lib", "examples", "example.slow5")
    blow = str(tmp_path / "example.blow5")
    subprocess.check_call([cli, "view", text, "-o", blow, "-c", "none", "-s", "none"], stderr=subprocess.DEVNULL)
    subprocess.check_call([cli, "index", blow], stderr=subprocess.DEVNULL)
    ids = [l.split(b"\t")[0].decode() for l in open(text, "rb").read().split(b"\n") if l and l[:1] not in (b"#", b"@")]
    assert len(ids) == 5
    src = tmp_path / "get.c"
    src.write_text(GET_PROG)
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "theirs")
    _cc(["-O1", "-I", os.path.join(ROOT, "include", "compat"), str(src), "-o", ours, "-L", LIBDIR, "-lslow5b200", "-Wl,-rpath," + LIBDIR])
    refdir = os.path.dirname(REF_SO)
    _cc(["-O1", "-I", os.path.join(REFTREE, "slow5lib", "include"), str(src), "-o", theirs, "-L", refdir, "-l:libslow5_ref.so",
         "-Wl,-rpath," + refdir, "-lm", "-lz"])
    args = [blow, ids[3], ids[0], "no-such-read"]
    a = subprocess.run([ours] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    b = subprocess.run([theirs] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr.decode(), b.stderr.decode())
    assert a.stdout == b.stdout
    out = a.stdout.decode()
    assert "get before idx_load -> -6" in out and "end -1 after 5" in out and "get no-such-read -> -7" in out
    # slow5_write: every record back into a new none / none file = the input file
    L = C.CDLL(os.path.join(LIBDIR, "libslow5b200.so"))
    vp = C.c_void_p
    L.s5b_open.restype = vp
    L.s5b_open.argtypes = [C.c_char_p, C.c_char_p]
    L.s5b_get_next.argtypes = [C.POINTER(vp), vp]
    L.s5b_write.argtypes = [vp, vp]
    L.s5b_hdr_copy.argtypes = [vp, vp]
    L.s5b_set_press.argtypes = [vp, C.c_int, C.c_int]
    for f in (L.s5b_close, L.s5b_hdr_write, L.s5b_rec_free):
        f.argtypes = [vp]
    back = str(tmp_path / "back.blow5")
    fin, fout = L.s5b_open(blow.encode(), b"r"), L.s5b_open(back.encode(), b"w")
    assert fin and fout and L.s5b_hdr_copy(fout, fin) == 0 and L.s5b_set_press(fout, 0, 0) == 0 and L.s5b_hdr_write(fout) > 0
    rec = vp()
    total = 0
    while L.s5b_get_next(C.byref(rec), fin) == 0:
        k = L.s5b_write(rec, fout)
        assert k > 8
        total += k
    L.s5b_rec_free(rec)
    assert L.s5b_close(fin) == 0 and L.s5b_close(fout) == 0
    assert open(back, "rb").read() == open(blow, "rb").read() and total > 0


INTRO_PROG = r"""
#include <stdio.h>
#include <stdlib.h>
#include <inttypes.h>
#include <slow5/slow5.h>
int main(int argc, char **argv) {
    slow5_file_t *sp = slow5_open(argv[1], "r");
    if (!sp) return 2;
    uint64_t n = 0;
    const char **keys = slow5_get_hdr_keys(sp->header, &n);
    printf("keys %" PRIu64 ":", n);
    for (uint64_t i = 0; i < n; ++i) printf(" %s", keys[i]);
    printf("\n");
    free(keys);
    char **names = slow5_get_aux_names(sp->header, &n);
    enum slow5_aux_type *types = slow5_get_aux_types(sp->header, &n);
    printf("aux %" PRIu64 ":", n);
    for (uint64_t i = 0; i < n; ++i) printf(" %s=%d%s", names[i], (int)types[i], SLOW5_IS_PTR(types[i]) ? "*" : "");
    printf("\n");
    if (argc > 2) {
        uint8_t k = 0;
        char **labels = slow5_get_aux_enum_labels(sp->header, argv[2], &k);
        printf("labels of %s: %u", argv[2], (unsigned)k);
        for (uint8_t i = 0; labels && i < k; ++i) printf(" %s", labels[i]);
        printf("\n");
    }
    uint64_t nr = 0;
    char **rids = slow5_get_rids(sp, &nr);
    printf("rids before idx_load: %s\n", rids ? "set" : "NULL");
    if (slow5_idx_load(sp) == 0) {
        rids = slow5_get_rids(sp, &nr);
        printf("rids %" PRIu64 ":", nr);
        for (uint64_t i = 0; i < nr; ++i) printf(" %s", rids[i]);
        printf("\n");
        slow5_idx_unload(sp);
        slow5_rec_t *rec = NULL;
        printf("get after unload -> %d\n", slow5_get("x", &rec, sp));
    }
    slow5_close(sp);
    return 0;
}
"""


@have_tree
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
def test_header_and_index_introspection_matches_the_reference(tmp_path):
    """slow5_get_hdr_keys / _get_aux_names / _get_aux_types / _get_aux_enum_labels / _get_rids (slow5.h:633-654) on the
    reference's example file and on its enum fixture: one program against both libraries prints the same"""
    cli = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
    src = tmp_path / "intro.c"
    src.write_text(INTRO_PROG)
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "theirs")
    _cc(["-O1", "-I", os.path.join(ROOT, "include", "compat"), str(src), "-o", ours, "-L", LIBDIR, "-lslow5b200", "-Wl,-rpath," + LIBDIR])
    refdir = os.path.dirname(REF_SO)
    _cc(["-O1", "-I", os.path.join(REFTREE, "slow5lib", "include"), str(src), "-o", theirs, "-L", refdir, "-l:libslow5_ref.so",
         "-Wl,-rpath," + refdir, "-lm", "-lz"])
    cases = [(os.path.join(REFTREE, "slow5lib", "examples", "example.slow5"), [])]
    enum_file = os.path.join(REFTREE, "test", "data", "raw", "merge", "aux_enum.slow5")
    if os.path.exists(enum_file):
        cases.append((enum_file, ["end_reason"]))
    for text, extra in cases:
        blow = str(tmp_path / (os.path.basename(text) + ".blow5"))
        subprocess.check_call([cli, "view", text, "-o", blow, "-c", "none", "-s", "none"], stderr=subprocess.DEVNULL)
        subprocess.check_call([cli, "index", blow], stderr=subprocess.DEVNULL)
        a = subprocess.run([ours, blow] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        b = subprocess.run([theirs, blow] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr.decode(), b.stderr.decode())
        assert a.stdout == b.stdout, (a.stdout.decode(), b.stdout.decode())
        assert b"rids before idx_load: NULL" in a.stdout and b"rids " in a.stdout


@have_tree
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
@pytest.mark.parametrize("example", ["sequential_read", "random_read", "header_attribute", "auxiliary_field"])
def test_reference_examples_build_unchanged(tmp_path, example):
    """slow5lib/examples/*.c (the programs the reference's documentation walks through: slow5_get_next, slow5_idx_load +
    slow5_get with the index created on the fly, slow5_hdr_get, slow5_aux_get_*), compiled UNCHANGED against include/compat and
    linked with libslow5b200.so, print what the reference build prints and leave the same index file behind.  Their inputs are
    SLOW5 text files: no codec, no device."""
    src = os.path.join(REFTREE, "slow5lib", "examples", example + ".c")
    refdir = os.path.dirname(REF_SO)
    results = []
    for who in ("ours", "theirs"):
        d = tmp_path / who
        os.makedirs(d / "examples")
        for f in ("example.slow5", "example2.slow5"):
            shutil.copy(os.path.join(REFTREE, "slow5lib", "examples", f), d / "examples" / f)
        exe = str(d / "prog")
        if who == "ours":
            _cc(["-O1", "-w", "-I", os.path.join(ROOT, "include", "compat"), src, "-o", exe, "-L", LIBDIR, "-lslow5b200",
                 "-Wl,-rpath," + LIBDIR])
        else:
            _cc(["-O1", "-w", "-I", os.path.join(REFTREE, "slow5lib", "include"), src, "-o", exe, "-L", refdir, "-l:libslow5_ref.so",
                 "-Wl,-rpath," + refdir, "-lm", "-lz"])
        r = subprocess.run([exe], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
        idx = {f: open(d / "examples" / f, "rb").read() for f in sorted(os.listdir(d / "examples")) if f.endswith(".idx")}
        results.append((r.returncode, r.stdout, r.stderr, idx))
    assert results[0][0] == 0 and results[1][0] == 0, (results[0][2].decode(), results[1][2].decode())
    assert results[0][1] == results[1][1] and len(results[0][1]) > 0
    assert results[0][2] == results[1][2]
    assert results[0][3] == results[1][3]
    if example == "random_read":
        assert list(results[0][3]) == ["example.slow5.idx"]


WRITE_PROG = r"""
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <slow5/slow5.h>
/* slow5lib/examples/write.c with two read groups, three records, values left unset, and no compression */
int main(int argc, char **argv) {
    slow5_file_t *sp = slow5_open(argv[1], "w");
    if (!sp) return 2;
    int is_text = strlen(argv[1]) > 6 && !strcmp(argv[1] + strlen(argv[1]) - 6, ".slow5");
    int pr = slow5_set_press(sp, SLOW5_COMPRESS_NONE, SLOW5_COMPRESS_NONE);   /* refused for SLOW5 text (slow5.c:585-589) */
    printf("set_press -> %s\n", pr < 0 ? "refused" : "ok");
    if (pr < 0 && !is_text) return 3;
    slow5_hdr_t *h = sp->header;
    int rc = 0;
    rc |= slow5_hdr_add("run_id", h) | slow5_hdr_add("asic_id", h) | slow5_hdr_add("zeta", h);
    printf("add again -> %d\n", slow5_hdr_add("run_id", h));
    printf("add_rg -> %d\n", (int)slow5_hdr_add_rg(h));
    rc |= slow5_hdr_set("run_id", "run_0", 0, h) | slow5_hdr_set("asic_id", "asic_id_0", 0, h) | slow5_hdr_set("run_id", "run_1", 1, h);
    printf("set unknown -> %d, set rg 7 -> %d\n", slow5_hdr_set("nope", "x", 0, h), slow5_hdr_set("run_id", "x", 7, h));
    rc |= slow5_aux_add("channel_number", SLOW5_STRING, h) | slow5_aux_add("median_before", SLOW5_DOUBLE, h);
    rc |= slow5_aux_add("read_number", SLOW5_INT32_T, h) | slow5_aux_add("start_mux", SLOW5_UINT8_T, h);
    rc |= slow5_aux_add("start_time", SLOW5_UINT64_T, h);
    const char *labels[] = {"unknown", "partial", "mux_change", "signal_positive"};
    const char *bad[] = {"ok", "9lives"};
    printf("add_enum bad label -> %d, add enum through aux_add -> %d\n", slow5_aux_add_enum("oops", bad, 2, h),
           slow5_aux_add("oops2", SLOW5_ENUM, h));
    rc |= slow5_aux_add_enum("end_reason", labels, 4, h);
    printf("aux_add again -> %d\n", slow5_aux_add("start_mux", SLOW5_UINT8_T, h));
    if (rc || slow5_hdr_write(sp) < 0) return 4;
    for (int r = 0; r < 3; ++r) {
        slow5_rec_t *rec = slow5_rec_init();
        char id[32];
        snprintf(id, sizeof id, "read_%d", r);
        rec->read_id = strdup(id);
        rec->read_id_len = strlen(id);
        rec->read_group = r & 1;
        rec->digitisation = 4096.0; rec->offset = 3.0 + r; rec->range = 10.0; rec->sampling_rate = 4000.0;
        rec->len_raw_signal = 10 + r;
        rec->raw_signal = (int16_t *)malloc(sizeof(int16_t) * rec->len_raw_signal);
        for (uint64_t i = 0; i < rec->len_raw_signal; ++i) rec->raw_signal[i] = (int16_t)(i * 7 - 20 * r);
        double mb = 0.1 + r; int32_t rn = 10 + r; uint8_t mux = 1; uint64_t st = 100 + r;
        if (r != 1 && slow5_aux_set_string(rec, "channel_number", r ? "512" : "0", h) < 0) return 5;   /* record 1: left unset */
        if (slow5_aux_set(rec, "median_before", &mb, h) < 0 || slow5_aux_set(rec, "read_number", &rn, h) < 0) return 6;
        if (r != 2 && slow5_aux_set(rec, "start_mux", &mux, h) < 0) return 7;                             /* record 2: left unset */
        if (slow5_aux_set(rec, "start_time", &st, h) < 0) return 8;
        uint8_t why = (uint8_t)(r + 1), beyond = 4;
        if (r != 1 && slow5_aux_set(rec, "end_reason", &why, h) < 0) return 9;
        if (r == 1) printf("enum value beyond the labels -> %d\n", slow5_aux_set(rec, "end_reason", &beyond, h));
        if (r == 0) printf("set unknown field -> %d, set string as primitive -> %d, primitive as string -> %d\n",
                           slow5_aux_set(rec, "nope", &mb, h), slow5_aux_set(rec, "channel_number", &mb, h),
                           slow5_aux_set_string(rec, "start_mux", "1", h));
        if (slow5_write(rec, sp) < 0) return 10;
        slow5_rec_free(rec);
    }
    slow5_close(sp);
    return 0;
}
"""


@have_tree
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
def test_writing_a_file_from_scratch_matches_the_reference(tmp_path):
    """slow5_hdr_add / _hdr_set / _hdr_add_rg / _aux_add / slow5_rec_init / slow5_aux_set[_array,_string] / slow5_write (the calls
    of slow5lib/examples/write.c): one program against both libraries writes the same uncompressed BLOW5 file, byte for byte,
    and prints the same return codes"""
    src = tmp_path / "write.c"
    src.write_text(WRITE_PROG)
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "theirs")
    _cc(["-O1", "-w", "-I", os.path.join(ROOT, "include", "compat"), str(src), "-o", ours, "-L", LIBDIR, "-lslow5b200", "-Wl,-rpath," + LIBDIR])
    refdir = os.path.dirname(REF_SO)
    _cc(["-O1", "-w", "-I", os.path.join(REFTREE, "slow5lib", "include"), str(src), "-o", theirs, "-L", refdir, "-l:libslow5_ref.so",
         "-Wl,-rpath," + refdir, "-lm", "-lz"])
    fa, fb = str(tmp_path / "ours.blow5"), str(tmp_path / "theirs.blow5")
    a = subprocess.run([ours, fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    b = subprocess.run([theirs, fb], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert a.returncode == 0 and b.returncode == 0, (a.returncode, a.stderr.decode(), b.returncode, b.stderr.decode())
    assert a.stdout == b.stdout, (a.stdout.decode(), b.stdout.decode())
    assert open(fa, "rb").read() == open(fb, "rb").read()
    # the same program writing SLOW5 text
    ta, tb = str(tmp_path / "ours.slow5"), str(tmp_path / "theirs.slow5")
    a = subprocess.run([ours, ta], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    b = subprocess.run([theirs, tb], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert a.returncode == 0 and b.returncode == 0, (a.returncode, a.stderr.decode(), b.returncode, b.stderr.decode())
    assert a.stdout == b.stdout and b"set_press -> refused" in a.stdout
    assert open(ta, "rb").read() == open(tb, "rb").read()
    # slow5lib/examples/append.c, UNCHANGED (it appends read_1 to "test.blow5" in the working directory): run on copies of the file
    # just written -- the columns it sets (channel_number ... start_time) exist there -- both builds leave the same file
    app = os.path.join(REFTREE, "slow5lib", "examples", "append.c")
    if os.path.exists(app):
        got = []
        for who, inc_lib in (("ours", ["-I", os.path.join(ROOT, "include", "compat"), app, "-L", LIBDIR, "-lslow5b200", "-Wl,-rpath," + LIBDIR]),
                             ("theirs", ["-I", os.path.join(REFTREE, "slow5lib", "include"), app, "-L", refdir, "-l:libslow5_ref.so",
                                         "-Wl,-rpath," + refdir, "-lm", "-lz"])):
            d = tmp_path / ("append_" + who)
            os.makedirs(d)
            shutil.copy(fa, d / "test.blow5")
            exe = str(d / "append")
            _cc(["-O1", "-w"] + inc_lib[:3] + ["-o", exe] + inc_lib[3:])
            r = subprocess.run([exe], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
            assert r.returncode == 0, (who, r.stderr.decode())
            got.append(open(d / "test.blow5", "rb").read())
        assert got[0] == got[1] and len(got[0]) > os.path.getsize(fa) and got[0].endswith(b"5WOLB")
    # and the file reads back as SLOW5 text through both CLIs identically
    cli = os.path.join(LIBDIR, "bin", "slow5tools-b200")
    mine = subprocess.check_output([cli, "view", fa], stderr=subprocess.DEVNULL)
    ref_cli = os.path.join(refdir, "slow5tools_ref")
    if os.path.exists(ref_cli):
        assert mine == subprocess.check_output([ref_cli, "view", fb], stderr=subprocess.DEVNULL)
    assert b"read_2" in mine and b"\t512\t" in mine


@have_tree
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
def test_reference_example_get_all_read_ids_unchanged(tmp_path):
    """slow5lib/examples/adv/get_all_read_ids.c (slow5_idx_load + slow5_get_rids), unchanged, on an uncompressed BLOW5 without an
    index file: both builds create the same index and print the same ids"""
    cli = os.path.join(LIBDIR, "bin", "slow5tools-b200")
    src = os.path.join(REFTREE, "slow5lib", "examples", "adv", "get_all_read_ids.c")
    refdir = os.path.dirname(REF_SO)
    out = []
    for who in ("ours", "theirs"):
        d = tmp_path / who
        os.makedirs(d)
        blow = str(d / "in.blow5")
        subprocess.check_call([cli, "view", os.path.join(REFTREE, "slow5lib", "examples", "example.slow5"), "-o", blow, "-c", "none",
                               "-s", "none"], stderr=subprocess.DEVNULL)
        exe = str(d / "prog")
        if who == "ours":
            _cc(["-O1", "-w", "-I", os.path.join(ROOT, "include", "compat"), src, "-o", exe, "-L", LIBDIR, "-lslow5b200",
                 "-Wl,-rpath," + LIBDIR])
        else:
            _cc(["-O1", "-w", "-I", os.path.join(REFTREE, "slow5lib", "include"), src, "-o", exe, "-L", refdir, "-l:libslow5_ref.so",
                 "-Wl,-rpath," + refdir, "-lm", "-lz"])
        r = subprocess.run([exe, "in.blow5"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        assert r.returncode == 0, r.stderr.decode()
        out.append((r.stdout, open(blow + ".idx", "rb").read()))
    assert out[0] == out[1] and out[0][0].count(b"\n") >= 5


@have_tree
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
@pytest.mark.parametrize("example", ["mt", "lazymt"])
def test_reference_mt_examples(tmp_path, example):
    """slow5lib/examples/mt/{mt,lazymt}.c -- the batch API as pyslow5 drives it: 4000 records written with slow5_write_batch[_lazy],
    read back with slow5_get_next_batch[_lazy] and by id with slow5_get_batch[_lazy].  The sources are used as they are except for
    ONE inserted line that switches compression off (the examples write zlib + svb-zd, which needs a device here; the unchanged
    sources are only compiled).  Both builds print the same and leave the same test.blow5 and index behind."""
    src = open(os.path.join(REFTREE, "slow5lib", "examples", "mt", example + ".c")).read()
    anchor = "    //set zstd record compression, svb-zd signal compression\n"
    assert src.count(anchor) == 1
    patched = src.replace(anchor, "    if (slow5_set_press(sf, SLOW5_COMPRESS_NONE, SLOW5_COMPRESS_NONE) < 0) exit(EXIT_FAILURE);\n" + anchor)
    refdir = os.path.dirname(REF_SO)
    got = []
    for who in ("ours", "theirs"):
        d = tmp_path / who
        os.makedirs(d)
        (d / "prog.c").write_text(patched)
        exe = str(d / "prog")
        if who == "ours":
            inc = ["-I", os.path.join(ROOT, "include", "compat")]
            lib = ["-L", LIBDIR, "-lslow5b200", "-lpthread", "-Wl,-rpath," + LIBDIR]
            # the unchanged source builds as well
            _cc(["-O1", "-w"] + inc + [os.path.join(REFTREE, "slow5lib", "examples", "mt", example + ".c"), "-o", str(d / "unchanged")] + lib)
        else:
            inc = ["-I", os.path.join(REFTREE, "slow5lib", "include")]
            lib = ["-L", refdir, "-l:libslow5_ref.so", "-Wl,-rpath," + refdir, "-lm", "-lz", "-lpthread"]
        _cc(["-O1", "-w"] + inc + [str(d / "prog.c"), "-o", exe] + lib)
        r = subprocess.run([exe], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, (who, r.stderr.decode()[-2000:])
        files = {f: open(d / f, "rb").read() for f in sorted(os.listdir(d)) if f.startswith("test.blow5")}
        got.append((r.stdout, files))
    assert got[0][0] == got[1][0] and got[0][0].count(b"\n") > 4000
    assert got[0][1] == got[1][1] and "test.blow5" in got[0][1]


@have_tree
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
@pytest.mark.parametrize("example", ["random_read_pthreads", "random_read_openmp", "adv/sequential_read_openmp"])
def test_reference_threaded_examples_build_unchanged(tmp_path, example):
    """the reference's user-threaded examples (slow5_get from several threads after slow5_idx_load; slow5_get_next_bytes + slow5_decode
    under OpenMP), UNCHANGED: the same lines from both builds (compared sorted: the thread order is not fixed)"""
    src = os.path.join(REFTREE, "slow5lib", "examples", example + ".c")
    refdir = os.path.dirname(REF_SO)
    out = []
    for who in ("ours", "theirs"):
        d = tmp_path / who
        os.makedirs(d / "examples")
        for f in ("example.slow5", "example2.slow5"):
            shutil.copy(os.path.join(REFTREE, "slow5lib", "examples", f), d / "examples" / f)
        exe = str(d / "prog")
        if who == "ours":
            _cc(["-O1", "-w", "-fopenmp", "-I", os.path.join(ROOT, "include", "compat"), src, "-o", exe, "-L", LIBDIR, "-lslow5b200",
                 "-lpthread", "-Wl,-rpath," + LIBDIR])
        else:
            _cc(["-O1", "-w", "-fopenmp", "-I", os.path.join(REFTREE, "slow5lib", "include"), src, "-o", exe, "-L", refdir,
                 "-l:libslow5_ref.so", "-Wl,-rpath," + refdir, "-lm", "-lz", "-lpthread"])
        r = subprocess.run([exe], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
        assert r.returncode == 0, r.stdout.decode()
        out.append(sorted(r.stdout.splitlines()))
    assert out[0] == out[1] and len(out[0]) >= 5
