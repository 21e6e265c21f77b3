"""Host mirror of the reference's batch codec slot (src/view.c:292 work_db; slow5_mt.c:336-359)
over the C-ABI: torch tensors carry the slabs, the library does all the work."""
import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import lib, S5BError, METHOD


def sig_layout(n_samples):
    """Signal slab layout: read r starts at a multiple of 8 samples (16-byte TMA granule).
    Returns uint64 offsets with len(n_samples)+1 entries."""
    n = np.asarray(n_samples, dtype=np.uint64)
    pad = (n + np.uint64(7)) // np.uint64(8) * np.uint64(8)
    off = np.zeros(len(n) + 1, dtype=np.uint64)
    np.cumsum(pad, out=off[1:])
    return off


def svb_slot_layout(n_samples):
    """Slot layout for encode output: slot r = worst-case svb-zd size of read r rounded to 16 bytes
    (s5b_svbzd_slot)."""
    n = np.asarray(n_samples, dtype=np.uint64)
    bound = np.uint64(4) + (n + np.uint64(3)) // np.uint64(4) + np.uint64(3) * n
    slot = (bound + np.uint64(15)) // np.uint64(16) * np.uint64(16)
    off = np.zeros(len(n) + 1, dtype=np.uint64)
    np.cumsum(slot, out=off[1:])
    return off


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


class Codec:
    """One per (process, GPU).  Wraps s5b_ctx_t."""

    def __init__(self, device=None):
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = int(device)
        h = C.c_void_p()
        rc = lib.s5b_ctx_create(self.device, C.byref(h))
        if rc != 0:
            raise S5BError(rc, "s5b_ctx_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib.s5b_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, where):
        if rc != 0:
            raise S5BError(rc, where, lib.s5b_ctx_last_cuda_error(self._h).decode())

    @property
    def launches(self):
        return int(lib.s5b_ctx_launch_count(self._h))

    @staticmethod
    def _stream():
        # torch's default stream has handle 0, which the C-ABI reads as "use the context's own stream";
        # name the legacy default stream explicitly (cudaStreamLegacy == 0x1) so launches are ordered with
        # torch's work and visible to torch.cuda.Event
        h = torch.cuda.current_stream().cuda_stream
        return C.c_void_p(h if h else 1)

    # ---- device-resident ---------------------------------------------------------------------
    def svbzd_encode_dev(self, sig, sig_off, n_samples, svb, svb_off, svb_len, status):
        self._check(lib.s5b_svbzd_encode_dev(self._h, _ptr(sig), _ptr(sig_off), _ptr(n_samples), n_samples.numel(),
                                             _ptr(svb), _ptr(svb_off), _ptr(svb_len), _ptr(status), self._stream()),
                    "s5b_svbzd_encode_dev")

    def svbzd_decode_dev(self, svb, svb_off, svb_len, sig, sig_off, n_samples, status):
        self._check(lib.s5b_svbzd_decode_dev(self._h, _ptr(svb), _ptr(svb_off), _ptr(svb_len), svb.numel(),
                                             svb_len.numel(), _ptr(sig), _ptr(sig_off), _ptr(n_samples), _ptr(status),
                                             self._stream()),
                    "s5b_svbzd_decode_dev")

    def svbzd_peek_dev(self, svb, svb_off, svb_len, n_samples):
        self._check(lib.s5b_svbzd_peek_dev(self._h, _ptr(svb), _ptr(svb_off), _ptr(svb_len), svb_len.numel(),
                                           _ptr(n_samples), self._stream()), "s5b_svbzd_peek_dev")

    def exzd_encode_dev(self, sig, sig_off, n_samples, out, out_off, out_len, status):
        self._check(lib.s5b_exzd_encode_dev(self._h, _ptr(sig), _ptr(sig_off), _ptr(n_samples), n_samples.numel(),
                                            _ptr(out), _ptr(out_off), _ptr(out_len), _ptr(status), self._stream()),
                    "s5b_exzd_encode_dev")

    def exzd_decode_dev(self, din, in_off, in_len, sig, sig_off, n_samples, status):
        self._check(lib.s5b_exzd_decode_dev(self._h, _ptr(din), _ptr(in_off), _ptr(in_len), din.numel(),
                                            in_len.numel(), _ptr(sig), _ptr(sig_off), _ptr(n_samples), _ptr(status),
                                            self._stream()),
                    "s5b_exzd_decode_dev")

    def zlib_inflate_dev(self, zin, in_off, in_len, out, out_off, out_len, status):
        self._check(lib.s5b_zlib_inflate_dev(self._h, _ptr(zin), _ptr(in_off), _ptr(in_len), zin.numel(),
                                             in_len.numel(), _ptr(out), _ptr(out_off), _ptr(out_len), _ptr(status),
                                             self._stream()), "s5b_zlib_inflate_dev")

    def zstd_decode_dev(self, zin, in_off, in_len, out, out_off, out_len, status):
        self._check(lib.s5b_zstd_decode_dev(self._h, _ptr(zin), _ptr(in_off), _ptr(in_len), zin.numel(),
                                            in_len.numel(), _ptr(out), _ptr(out_off), _ptr(out_len), _ptr(status),
                                            self._stream()), "s5b_zstd_decode_dev")

    def zlib_deflate_dev(self, din, in_off, in_len, out, out_off, out_len, status, split=None):
        self._check(lib.s5b_zlib_deflate_dev(self._h, _ptr(din), _ptr(in_off), _ptr(in_len), din.numel(), _ptr(split),
                                             in_len.numel(), _ptr(out), _ptr(out_off), _ptr(out_len), _ptr(status),
                                             self._stream()), "s5b_zlib_deflate_dev")

    def zstd_encode_dev(self, din, in_off, in_len, out, out_off, out_len, status, split=None):
        self._check(lib.s5b_zstd_encode_dev(self._h, _ptr(din), _ptr(in_off), _ptr(in_len), din.numel(), _ptr(split),
                                            in_len.numel(), _ptr(out), _ptr(out_off), _ptr(out_len), _ptr(status),
                                            self._stream()), "s5b_zstd_encode_dev")

    def compact_dev(self, src, src_off, length, dst, dst_off, align=16):
        self._check(lib.s5b_compact_dev(self._h, _ptr(src), _ptr(src_off), _ptr(length), length.numel(), align,
                                        _ptr(dst), _ptr(dst_off), self._stream()), "s5b_compact_dev")

    # ---- host slabs ---------------------------------------------------------------------------
    def svbzd_encode_host(self, sig, sig_off, n_samples, svb, svb_off, svb_len, status, check=True):
        """sig/svb: host tensors or arrays (pinned for speed); returns the call's return code."""
        n = len(n_samples)
        rc = lib.s5b_svbzd_encode_host(self._h, _ptr(sig), _ptr(sig_off), _ptr(n_samples), n, _ptr(svb),
                                       svb.numel() if isinstance(svb, torch.Tensor) else svb.size,
                                       _ptr(svb_off), _ptr(svb_len), _ptr(status))
        if check:
            self._check(rc, "s5b_svbzd_encode_host")
        return rc

    def svbzd_decode_host(self, svb, svb_off, svb_len, sig, sig_off, n_samples, status, check=True):
        n = len(svb_len)
        rc = lib.s5b_svbzd_decode_host(self._h, _ptr(svb), _ptr(svb_off), _ptr(svb_len), n, _ptr(sig),
                                       sig.numel() if isinstance(sig, torch.Tensor) else sig.size,
                                       _ptr(sig_off), _ptr(n_samples), _ptr(status))
        if check:
            self._check(rc, "s5b_svbzd_decode_host")
        return rc

    # ---- pointer arrays (db_t shape): list of bytes-like in, list of bytes out -----------------
    def _batch(self, fn, method, bufs):
        n = len(bufs)
        keep = [np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8) for b in bufs]
        ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
        counts = (C.c_size_t * n)(*[len(b) for b in bufs])
        outp = (C.c_void_p * n)()
        outn = (C.c_size_t * n)()
        rc = fn(self._h, method, ptrs, counts, n, outp, outn)
        res = []
        for i in range(n):
            if outp[i]:
                res.append(C.string_at(outp[i], outn[i]))
                _capi.free(outp[i])
            else:
                res.append(None)
        return rc, res

    def blow5_recode(self, in_rec, in_sig, out_rec, out_sig, records):
        """s5b_blow5_recode_host on a list of packed records (bytes, as stored in the file without their size prefix).
        Returns (rc, file image bytes: [u64 size][record] per record)."""
        n = len(records)
        rec_len = np.array([len(r) for r in records], np.uint32)
        rec_off = np.zeros(n + 1, np.uint64)
        np.cumsum((rec_len.astype(np.uint64) + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=rec_off[1:])
        h_in = np.zeros(int(rec_off[-1]) + 16, np.uint8)
        for r, o in zip(records, rec_off):
            h_in[int(o):int(o) + len(r)] = np.frombuffer(r, np.uint8)
        cap = int(rec_len.sum()) * 4 + 64 * n + 4096
        for _ in range(2):
            h_out = np.zeros(cap, np.uint8)
            nb = C.c_uint64()
            rc = lib.s5b_blow5_recode_host(self._h, in_rec, in_sig, out_rec, out_sig, h_in.ctypes.data, h_in.size,
                                           rec_off.ctypes.data, rec_len.ctypes.data, n, h_out.ctypes.data, cap, C.byref(nb))
            if rc == _capi.ERR.NOSPACE and nb.value > cap:
                cap = int(nb.value) + 16
                continue
            break
        return rc, h_out[:nb.value].tobytes() if rc == 0 else None

    # ---- whole-batch record transcoding (the `view` worker for a batch; src/view.c:254-301) ----------------
    def blow5_recode_batch_host(self, in_rec, in_sig, out_rec, out_sig, h_in, in_bytes, rec_off, rec_len, h_out,
                                img_off=None, check=True):
        """s5b_blow5_recode_batch_host: h_in / h_out host tensors or arrays (pinned for speed), rec_off / rec_len numpy
        (uint64 / uint32).  Returns (rc, image bytes); img_off (uint64[n+1]) receives the image offsets."""
        n = len(rec_len)
        nb = C.c_uint64()
        cap = h_out.numel() if isinstance(h_out, torch.Tensor) else h_out.size
        rc = lib.s5b_blow5_recode_batch_host(self._h, in_rec, in_sig, out_rec, out_sig, _ptr(h_in), int(in_bytes),
                                             _ptr(rec_off), _ptr(rec_len), n, _ptr(h_out), cap, C.byref(nb), _ptr(img_off))
        if check:
            self._check(rc, "s5b_blow5_recode_batch_host")
        return rc, int(nb.value)

    def blow5_recode_dev(self, in_rec, in_sig, out_rec, out_sig, d_in, in_bytes, rec_off, rec_len, d_out, d_result,
                         d_img_off=None):
        """s5b_blow5_recode_dev: payload / image in HBM (torch uint8), record table numpy on the host; asynchronous --
        call sync() before reading d_result (uint64[2]: image bytes, first error)."""
        self._check(lib.s5b_blow5_recode_dev(self._h, in_rec, in_sig, out_rec, out_sig, _ptr(d_in), int(in_bytes),
                                             _ptr(rec_off), _ptr(rec_len), len(rec_len), _ptr(d_out), d_out.numel(),
                                             _ptr(d_result), _ptr(d_img_off)), "s5b_blow5_recode_dev")

    def sync(self):
        self._check(lib.s5b_ctx_sync(self._h), "s5b_ctx_sync")

    def recode_stream(self):
        """torch view of the stream the device-resident transcoder enqueues on (for CUDA-event timing)."""
        h = lib.s5b_ctx_recode_stream(self._h)
        if not h:
            raise S5BError(_capi.ERR.DEVICE, "s5b_ctx_recode_stream")
        return torch.cuda.ExternalStream(h, device=torch.device("cuda", self.device))

    def stage_timing(self, enable=True):
        self._check(lib.s5b_ctx_stage_timing(self._h, 1 if enable else 0), "s5b_ctx_stage_timing")

    def stage_report(self, reset=True):
        """{stage name: (milliseconds, launch groups)} accumulated since the last reset."""
        k = lib.s5b_stage_count()
        ms = (C.c_double * k)()
        cnt = (C.c_uint64 * k)()
        self._check(lib.s5b_ctx_stage_report(self._h, ms, cnt, 1 if reset else 0), "s5b_ctx_stage_report")
        return {lib.s5b_stage_name(i).decode(): (ms[i], int(cnt[i])) for i in range(k)}

    def set_recode_workspace(self, max_bytes):
        """workspace budget of blow5_recode_dev in bytes (0 = sized from the free device memory)"""
        self._check(lib.s5b_ctx_set_recode_workspace(self._h, int(max_bytes)), "s5b_ctx_set_recode_workspace")

    def set_degrade(self, bits, check_dataset=False, digitisation=0.0, sampling_rate=0.0):
        """src/degrade.c:240-263: while bits > 0 every transcoding pass rounds that many low bits of each sample away"""
        self._check(lib.s5b_ctx_set_degrade(self._h, int(bits), int(bool(check_dataset)), float(digitisation),
                                            float(sampling_rate)), "s5b_ctx_set_degrade")

    def qts_round_dev(self, sig, bits, n_samples=None):
        """slow5_arr_qts_round (slow5_press.c:1991-2005) on an int16 CUDA tensor, in place"""
        n = sig.numel() if n_samples is None else int(n_samples)
        self._check(lib.s5b_qts_round_dev(self._h, _ptr(sig), n, int(bits), self._stream()), "s5b_qts_round_dev")

    def qts_round_batch(self, bits, bufs):
        """host arrays of int16 samples (bytes) -> degraded copies"""
        return self._batch(lambda h, m, *a: lib.s5b_qts_round_batch_host(h, int(bits), *a), 0, bufs)

    def compress_batch(self, method, bufs):
        return self._batch(lib.s5b_compress_batch_host, method, bufs)

    def depress_batch(self, method, bufs):
        return self._batch(lib.s5b_depress_batch_host, method, bufs)


def ptr_compress_solo(method, data):
    """slow5_ptr_compress_solo twin: bytes in, bytes out (None on failure; see s5b_last_error)."""
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    n = C.c_size_t()
    p = lib.s5b_ptr_compress_solo(method, buf.ctypes.data, len(data), C.byref(n))
    if not p:
        return None
    out = C.string_at(p, n.value)
    _capi.free(p)
    return out


def ptr_depress_solo(method, data):
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    n = C.c_size_t()
    p = lib.s5b_ptr_depress_solo(method, buf.ctypes.data, len(data), C.byref(n))
    if not p:
        return None
    out = C.string_at(p, n.value)
    _capi.free(p)
    return out
