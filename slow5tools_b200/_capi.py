"""ctypes binding of include/slow5b200.h.  Fails loudly when the in-tree library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path():
    # S5B_LIBRARY: development override used for A/B runs of kernel variants (csrc/Makefile `variant`)
    return os.environ.get("S5B_LIBRARY") or os.path.join(_HERE, "libslow5b200.so")


class S5BError(RuntimeError):
    def __init__(self, code, where="", detail=""):
        self.code = code
        msg = f"{where}: {strerror(code)} ({code})"
        if detail:
            msg += f" [{detail}]"
        super().__init__(msg)


class ERR:
    OK = 0
    ARG = -2
    MEM = -10
    PRESS = -13
    NOSPACE = -40
    DEVICE = -41
    DATASET = -42


class METHOD:  # enum slow5_press_method, slow5_press.h:61-67
    NONE = 0
    ZLIB = 1
    SVB_ZD = 2
    ZSTD = 3
    EX_ZD = 4


def _load():
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C slow5tools_b200/csrc` (there is no CPU fallback)")
    return C.CDLL(path)


lib = _load()

_vp, _u64, _u32, _i32, _sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_size_t
_P = C.POINTER

_SIGS = {
    "s5b_version": (C.c_char_p, []),
    "s5b_strerror": (C.c_char_p, [C.c_int]),
    "s5b_device_count": (C.c_int, []),
    "s5b_ctx_create": (C.c_int, [C.c_int, _P(_vp)]),
    "s5b_ctx_destroy": (None, [_vp]),
    "s5b_ctx_last_cuda_error": (C.c_char_p, [_vp]),
    "s5b_ctx_launch_count": (_u64, [_vp]),
    "s5b_svbzd_bound": (_u64, [_u32]),
    "s5b_svbzd_slot": (_u64, [_u32]),
    "s5b_svbzd_encode_dev": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "s5b_svbzd_decode_dev": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "s5b_svbzd_peek_dev": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _vp, _vp]),
    "s5b_exzd_bound": (_u64, [_u32]),
    "s5b_exzd_slot": (_u64, [_u32]),
    "s5b_exzd_encode_dev": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "s5b_exzd_decode_dev": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "s5b_compact_dev": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _u32, _vp, _vp, _vp]),
    "s5b_svbzd_encode_host": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp]),
    "s5b_svbzd_decode_host": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp]),
    "s5b_compress_batch_host": (C.c_int, [_vp, C.c_int, _P(_vp), _P(_sz), _sz, _P(_vp), _P(_sz)]),
    "s5b_depress_batch_host": (C.c_int, [_vp, C.c_int, _P(_vp), _P(_sz), _sz, _P(_vp), _P(_sz)]),
    "s5b_blow5_recode_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _u64, _vp, _vp, _u64, _vp, _u64,
                                        _P(_u64)]),
    "s5b_blow5_recode_batch_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _u64, _vp, _vp, _u64, _vp, _u64,
                                              _P(_u64), _vp]),
    "s5b_blow5_recode_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _u64, _vp, _vp, _u64, _vp, _u64, _vp, _vp]),
    "s5b_ctx_sync": (C.c_int, [_vp]),
    "s5b_ctx_recode_stream": (_vp, [_vp]),
    "s5b_ctx_stage_timing": (C.c_int, [_vp, C.c_int]),
    "s5b_ctx_stage_report": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "s5b_stage_count": (C.c_int, []),
    "s5b_stage_name": (C.c_char_p, [C.c_int]),
    "s5b_ctx_set_aux_layout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "s5b_ctx_set_rg_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "s5b_ctx_set_recode_workspace": (C.c_int, [C.c_void_p, C.c_uint64]),
    "s5b_ctx_set_degrade": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]),
    "s5b_qts_round_dev": (C.c_int, [_vp, _vp, _u64, C.c_int, _vp]),
    "s5b_qts_round_batch_host": (C.c_int, [_vp, C.c_int, _P(_vp), _P(_sz), _sz, _P(_vp), _P(_sz)]),
    "s5b_ptr_compress_solo": (_vp, [C.c_int, _vp, _sz, _P(_sz)]),
    "s5b_ptr_depress_solo": (_vp, [C.c_int, _vp, _sz, _P(_sz)]),
    "s5b_last_error": (C.c_int, []),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

_libc = C.CDLL(None)
_libc.free.argtypes = [_vp]
_libc.free.restype = None


def free(ptr):
    _libc.free(ptr)


def strerror(code):
    return lib.s5b_strerror(int(code)).decode()

lib.s5b_zlib_inflate_dev.restype = C.c_int
lib.s5b_zlib_inflate_dev.argtypes = [_vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp]
lib.s5b_zlib_bound.restype = _u64
lib.s5b_zlib_bound.argtypes = [_u64]
lib.s5b_zlib_deflate_dev.restype = C.c_int
lib.s5b_zlib_deflate_dev.argtypes = [_vp, _vp, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _vp, _vp]
lib.s5b_zstd_decode_dev.restype = C.c_int
lib.s5b_zstd_decode_dev.argtypes = [_vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp]
lib.s5b_zstd_content_size.restype = C.c_int
lib.s5b_zstd_content_size.argtypes = [_vp, _sz, _P(_u64)]
lib.s5b_zstd_bound.restype = _u64
lib.s5b_zstd_bound.argtypes = [_u64]
lib.s5b_zstd_encode_dev.restype = C.c_int
lib.s5b_zstd_encode_dev.argtypes = [_vp, _vp, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _vp, _vp]
