#!/usr/bin/env python
"""bench.py -- BASELINE.json's north-star metric: BLOW5 reads/s for the full record path (svb-zd + zlib), encode and
decode, on BASELINE config[2]: synthetic 1M reads x 4096 int16 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R] [--samples S]

One "step" = one encode pass and one decode pass over the batch, the two directions of `slow5tools view`
(src/view.c:241-323 driving src/view.c:35-57 per record):
    encode: uncompressed BLOW5 records -> svb-zd signal -> packed record -> zlib record -> file image
    decode: zlib records -> inflate -> locate -> svb-zd decode -> uncompressed record -> file image
`value`  = reads/s with the batch resident in HBM (s5b_blow5_recode_dev, CUDA events on the transcoding stream).
`e2e`    = the same two passes through the host-buffer C-ABI call (s5b_blow5_recode_batch_host: pinned host slabs,
           every H2D / D2H copy inside the timed region).
`roofline` is on the slowest kernel of the step (deflate or inflate), from per-stage CUDA events recorded inside the
timed region.  `--impl reference` runs the UNMODIFIED reference library (oracle/_ref/libslow5_ref.so:
slow5_decode + slow5_encode per record, the body of view's worker) on all host threads over a bounded sample.
Multi-GPU: one rank per GPU (torchrun), every rank transcodes its own reads: no collective on the data path (weak).
With N > 1 the line also carries `split_compare`: one fixed batch held by rank 0 and split across the GPUs, through
per-GPU PCIe copies versus an NCCL scatter / gather of byte-balanced shards over NVLink.
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BLOW5 svb-zd+zlib reads/sec (encode+decode)"
UNIT = "reads/s"
M_NONE, M_ZLIB, M_SVB_ZD, M_EX_ZD = 0, 1, 2, 4
# dram__bytes_read.sum + dram__bytes_write.sum per 100 000 records from the committed `ncu --set full` capture
# (profiles/r2_ncu_full_v2.md, r2_ncu_full_final2.md: one 250 000-record chunk); bench.py cannot run under a profiler, so the figure is carried and
# scaled to the launch size
NCU_TRAFFIC_PER_100K = {"record_press": ((1338.1e6 + 212.4e6) + (240.4e6 + 59.4e6) + (866.6e6 + 984.5e6) + (1729.4e6 + 867.3e6)) / 2.5,
                        "record_depress": (1423.4e6 + 1642.1e6) / 2.5}    # count + tree + header + emit / inflate_thread
NCU_TRAFFIC_SOURCE = ("profiles/r2_ncu_full_v2.md (deflate kernels) and profiles/r2_ncu_full_final2.md (inflate_thread_kernel): ncu --set full, "
                      "one 250k-record chunk of the same workload, scaled by launch size")
STAGE_KERNELS = {"record_press": "deflate_count_kernel + deflate_tree_kernel + deflate_header_kernel + deflate_emit_kernel "
                                 "(one launch group per chunk)",
                 "record_depress": "inflate_thread_kernel (+ the warp-per-stream inflate_kernel's sweep for long streams)"}


def load_synth():
    """slow5tools_b200/synth.py without importing the package (which loads the product library): the reference arm must
    not map libslow5b200.so."""
    spec = importlib.util.spec_from_file_location("s5b_synth", os.path.join(ROOT, "slow5tools_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def config_dict(args):
    """identical in both arms (the driver compares them)"""
    return {"workload": "full BLOW5 record path, svb-zd + zlib: encode pass + decode pass over %d reads x %d int16 per GPU "
                        "(BASELINE config[2])" % (args.reads, args.samples),
            "reads_per_gpu": args.reads, "samples_per_read": args.samples, "record_press": "zlib", "signal_press": "svb-zd",
            "record_bytes_uncompressed": 82 + 2 * args.samples,
            "signal_model": "SURVEY 8d nanopore-like (levels N(500,70), dwell Geom(10), noise N(0,9)), seed 42+rank",
            "l2": "each pass streams > 8 GB per GPU, far beyond the 126 MB L2; no flush needed"}


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_info():
    info = {"nproc": os.cpu_count()}
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                info["cpu_model"] = ln.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                info["mem_available_gb"] = round(int(ln.split()[1]) / 1e6, 1)
    except Exception:
        pass
    return info


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# the reference on the host cores (oracle/: test + baseline infrastructure, never on the product path)
# ---------------------------------------------------------------------------------------------------------------------
class RefDriver:
    def __init__(self):
        so = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
        self.L = L = C.CDLL(so)
        L.refdrv_record_pass.restype = C.c_int
        L.refdrv_record_pass.argtypes = [C.c_char_p, C.c_char_p] + [C.c_int] * 4 + [
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
            C.c_void_p, C.POINTER(C.c_double)]
        ref = os.path.join(ROOT, "oracle", "_ref", "libslow5_ref.so")
        self.ref = ref.encode() if os.path.exists(ref) else None
        self.kind = "reference" if self.ref else "port"
        self.tmp = os.environ.get("TMPDIR", "/tmp").encode()

    def record_pass(self, methods, src, off, length, threads, out=None, out_off=None):
        """One pass of view's per-record worker over `len(length)` stored records; returns (seconds, image bytes)."""
        nb, sec = C.c_uint64(), C.c_double()
        rc = self.L.refdrv_record_pass(self.ref, self.tmp, methods[0], methods[1], methods[2], methods[3],
                                       src.ctypes.data, off.ctypes.data, length.ctypes.data, len(length), threads,
                                       out.ctypes.data if out is not None else None, out.size if out is not None else 0,
                                       C.byref(nb), out_off.ctypes.data if out_off is not None else None, C.byref(sec))
        if rc != 0:
            raise RuntimeError("reference driver failed: %d" % rc)
        return sec.value, nb.value


class RefWorkload:
    """encode + decode of `reads` records through the reference, with everything allocated once"""

    def __init__(self, drv, rec, threads):
        self.drv, self.threads = drv, threads
        self.rec = np.ascontiguousarray(rec).reshape(-1)
        self.R, self.rl = rec.shape
        self.off = np.arange(self.R, dtype=np.uint64) * np.uint64(self.rl)
        self.len = np.full(self.R, self.rl, np.uint32)
        self.enc = np.zeros(self.R * (self.rl // 2 + 512), np.uint8)
        self.enc_off = np.zeros(self.R + 1, np.uint64)

    def step(self, reads):
        """returns (encode seconds, decode seconds, encoded bytes)"""
        te, nb = self.drv.record_pass((M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), self.rec, self.off[:reads], self.len[:reads],
                                      self.threads, self.enc, self.enc_off)
        zoff = self.enc_off[:reads] + np.uint64(8)
        zlen = (self.enc_off[1:reads + 1] - self.enc_off[:reads] - np.uint64(8)).astype(np.uint32)
        td, _ = self.drv.record_pass((M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), self.enc, zoff, zlen, self.threads)
        return te, td, nb


def cpu_sample_records(synth, args, reads):
    sig = synth.nanopore_signal(reads * args.samples, seed=42)
    return synth.blow5_records(sig, reads, args.samples, seed=42).numpy()


def cpu_baseline(rec, budget_s=14.0):
    """Times the reference CPU path (all host threads) on a bounded sample of the same workload."""
    drv = RefDriver()
    cores = os.cpu_count() or 1
    w = RefWorkload(drv, rec, cores)
    probe = min(w.R, 256 * cores)
    te, td, _ = w.step(probe)
    rate = probe / (te + td)
    reads = int(max(probe, min(w.R, rate * budget_s)))
    te, td, nb = w.step(reads)
    out = {"value": reads / (te + td), "unit": UNIT, "cores": cores, "kind": drv.kind,
           "sample": "%d records of the same synthetic batch: slow5_decode + slow5_encode per record (view's worker, "
                     "src/view.c:35-57) in a %d-thread fork-join pool, encode pass then decode pass" % (reads, cores),
           "encode_reads_per_s": reads / te, "decode_reads_per_s": reads / td, "zlib_bytes_per_read": nb / reads - 8}
    out.update(host_info())
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth = load_synth()
    drv = RefDriver()
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    pool = min(args.reads, 60000)
    rec = cpu_sample_records(synth, args, pool)
    w = RefWorkload(drv, rec, cores)
    probe = min(pool, 256 * cores)
    te, td, _ = w.step(probe)
    rate = probe / (te + td)
    per_step = int(max(256, min(pool, rate * min(12.0, 150.0 / max(1, total_steps)))))
    for _ in range(args.warmup):
        w.step(per_step)
    t_e = t_d = 0.0
    for _ in range(args.steps):
        te, td, _ = w.step(per_step)
        t_e += te
        t_d += td
    dt = t_e + t_d
    value = per_step * args.steps / dt
    sample = ("%d records per step (bounded sample of the %d-read workload): slow5_decode + slow5_encode per record "
              "(view's worker) in a %d-thread pool, encode pass then decode pass" % (per_step, args.reads, cores))
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": drv.kind, "sample": sample}
    cb.update(host_info())
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_dict(args),
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "encode_reads_per_s": per_step * args.steps / t_e, "decode_reads_per_s": per_step * args.steps / t_d}
    print(json.dumps(out), flush=True)


def bind_to_gpu_numa_node(torch, local):
    """Pins this rank (and so its pinned-buffer allocations and copy submissions) to the CPUs next to its GPU."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return {"node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def pinned(torch, nbytes):
    return torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)


def pcie_probe(torch, barrier, seconds=0.25, mb=512):
    """Concurrent pinned H2D + D2H on this rank's link while every other rank does the same: the platform's ceiling for
    the e2e path.  Returns (h2d GB/s, d2h GB/s) of this rank under duplex load."""
    n = mb << 20
    h_a, h_b = pinned(torch, n), pinned(torch, n)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    reps = 4
    for attempt in range(2):
        barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(s_up):
            e[0].record()
            for _ in range(reps):
                d_a.copy_(h_a, non_blocking=True)
            e[1].record()
        with torch.cuda.stream(s_dn):
            e[2].record()
            for _ in range(reps):
                h_b.copy_(d_b, non_blocking=True)
            e[3].record()
        torch.cuda.synchronize()
        up, dn = e[0].elapsed_time(e[1]) * 1e-3, e[2].elapsed_time(e[3]) * 1e-3
        if attempt == 0:
            reps = max(2, int(reps * seconds / max(up, dn, 1e-3)))
    return n * reps / up / 1e9, n * reps / dn / 1e9


def run_ours(args):
    import torch
    import torch.distributed as dist
    import slow5tools_b200 as s5
    from slow5tools_b200 import synth
    from slow5tools_b200.dist import max_over_ranks, sum_over_ranks, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints exactly one JSON line on stdout: NCCL's version banner / debug lines go to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "WARN", "VERSION"):
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    R, N, K, W = args.reads, args.samples, args.steps, args.warmup
    cdc = s5.Codec(local)
    rl = synth.record_bytes(N)
    ENC = (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD)
    DEC = (M_ZLIB, M_SVB_ZD, M_NONE, M_NONE)

    # ---- the batch, resident in HBM: R uncompressed records (what the encode direction of `view` reads)
    sig = synth.nanopore_signal(R * N, seed=42 + rank, device="cuda")
    d_raw = synth.blow5_records(sig, R, N, seed=42 + rank).view(-1)
    del sig
    raw_off = np.arange(R, dtype=np.uint64) * np.uint64(rl)
    raw_len = np.full(R, rl, np.uint32)
    enc_cap = R * (rl // 2 + 512)
    d_enc = torch.zeros(enc_cap, dtype=torch.uint8, device="cuda")
    d_enc_off = torch.zeros(R + 1, dtype=torch.int64, device="cuda")
    d_back = torch.zeros(R * (rl + 8) + 64, dtype=torch.uint8, device="cuda")
    d_res_e = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_res_d = torch.zeros(2, dtype=torch.int64, device="cuda")
    stream = cdc.recode_stream()

    def encode_dev():
        cdc.blow5_recode_dev(*ENC, d_raw, R * rl, raw_off, raw_len, d_enc, d_res_e, d_enc_off)

    # first encode: learn the table of the encoded records (deterministic: the input never changes)
    encode_dev()
    cdc.sync()
    res = d_res_e.cpu().numpy()
    assert int(res[1]) == 0, "encode pass failed: %s" % s5.strerror(int(res[1]))
    enc_bytes = int(res[0])
    enc_img_off = d_enc_off.cpu().numpy().view(np.uint64)
    z_off = enc_img_off[:-1] + np.uint64(8)
    z_len = (enc_img_off[1:] - enc_img_off[:-1] - np.uint64(8)).astype(np.uint32)

    def decode_dev():
        cdc.blow5_recode_dev(*DEC, d_enc, enc_bytes, z_off, z_len, d_back, d_res_d, None)

    def check_results(what):
        re_, rd_ = d_res_e.cpu().numpy(), d_res_d.cpu().numpy()
        assert int(re_[1]) == 0 and int(rd_[1]) == 0, "%s: pass failed (%d, %d)" % (what, int(re_[1]), int(rd_[1]))
        assert int(re_[0]) == enc_bytes and int(rd_[0]) == R * (rl + 8), "%s: image sizes changed" % what

    def check_round_trip(back, what):
        b = back[:R * (rl + 8)].view(R, rl + 8)
        want = torch.tensor(list(np.uint64(rl).tobytes()), dtype=torch.uint8, device=b.device)
        ok = bool((b[:, :8] == want).all()) and torch.equal(b[:, 8:], d_raw.view(R, rl))
        assert ok, "%s: decode(encode(records)) differs from the records" % what

    for _ in range(W):
        encode_dev()
        decode_dev()
    cdc.sync()
    check_results("warm-up")
    check_round_trip(d_back, "warm-up")
    barrier()
    l0 = cdc.launches
    cdc.stage_timing(True)
    cdc.stage_report(reset=True)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    for k in range(K):
        ev[k][0].record(stream)
        encode_dev()
        ev[k][1].record(stream)
        decode_dev()
        ev[k][2].record(stream)
    cdc.sync()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = cdc.launches - l0
    stages = cdc.stage_report(reset=True)
    cdc.stage_timing(False)
    ms_total = ev[0][0].elapsed_time(ev[K - 1][2])
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    check_results("timed region")
    check_round_trip(d_back, "timed region")
    ms_total, enc_ms, dec_ms = max_over_ranks([ms_total, enc_ms, dec_ms])
    stage_ms = {k: v[0] / K for k, v in stages.items()}            # per step
    stage_launches = {k: v[1] / K for k, v in stages.items()}

    # per-record sizes of the intermediate forms (for the algorithmic byte counts): svb-zd stream bytes from the
    # records' own length fields after a signal-only encode would need another pass; the packed record is the
    # inflated size of an encoded record, which the decode pass reports through its image: head + 8 + C_svb
    # -> measure C_packed directly: transcode zlib+svb-zd -> none+svb-zd for a sample
    SMP = min(R, 20000)
    d_tmp = torch.zeros(SMP * (rl + 64), dtype=torch.uint8, device="cuda")
    d_res_t = torch.zeros(2, dtype=torch.int64, device="cuda")
    cdc.blow5_recode_dev(M_ZLIB, M_SVB_ZD, M_NONE, M_SVB_ZD, d_enc, enc_bytes, z_off[:SMP], z_len[:SMP], d_tmp, d_res_t, None)
    cdc.sync()
    packed_bytes_per_read = int(d_res_t[0]) / SMP - 8
    del d_tmp

    # ---- extra (not part of the step): one `degrade` pass over a slice of the batch -- zlib + svb-zd records in, 3 low bits of
    # every sample rounded away (src/degrade.c:255), zlib + ex-zd records out -- with the rounding kernel's own stage time
    degrade = None
    if not args.profile and not args.e2e_only:
        DG, BITS = min(R, 262144), 3
        d_dg = torch.zeros(DG * (rl // 2 + 512), dtype=torch.uint8, device="cuda")
        d_dg_off = torch.zeros(DG + 1, dtype=torch.int64, device="cuda")
        d_dg_back = torch.zeros(DG * (rl + 8) + 64, dtype=torch.uint8, device="cuda")
        d_res_g = torch.zeros(2, dtype=torch.int64, device="cuda")
        cdc.set_degrade(BITS)
        cdc.blow5_recode_dev(M_ZLIB, M_SVB_ZD, M_ZLIB, M_EX_ZD, d_enc, enc_bytes, z_off[:DG], z_len[:DG], d_dg, d_res_g, d_dg_off)
        cdc.sync()
        cdc.stage_timing(True)
        cdc.stage_report(reset=True)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        cdc.blow5_recode_dev(M_ZLIB, M_SVB_ZD, M_ZLIB, M_EX_ZD, d_enc, enc_bytes, z_off[:DG], z_len[:DG], d_dg, d_res_g, d_dg_off)
        g1.record(stream)
        cdc.sync()
        dg_stages = cdc.stage_report(reset=True)
        cdc.stage_timing(False)
        cdc.set_degrade(0)
        assert int(d_res_g[1]) == 0, "degrade pass failed"
        dg_bytes = int(d_res_g[0])
        go = d_dg_off.cpu().numpy().view(np.uint64)
        cdc.blow5_recode_dev(M_ZLIB, M_EX_ZD, M_NONE, M_NONE, d_dg, dg_bytes, go[:-1] + np.uint64(8),
                             (go[1:] - go[:-1] - np.uint64(8)).astype(np.uint32), d_dg_back, d_res_g, None)
        cdc.sync()
        assert int(d_res_g[1]) == 0 and int(d_res_g[0]) == DG * (rl + 8), "reading the degraded records back failed"
        hs = rl - 2 * N   # bytes of a record in front of its samples (no auxiliary fields in this batch)
        got = d_dg_back[:DG * (rl + 8)].view(DG, rl + 8)[:, 8 + hs:].contiguous().view(torch.int16)
        src = d_raw[:DG * rl].view(DG, rl)[:, hs:].contiguous().view(torch.int16)
        want = ((src.to(torch.int32) + (1 << (BITS - 1))) & ~((1 << BITS) - 1)).to(torch.int16)
        assert torch.equal(got, want), "degrade: samples are not round(signal, 3 bits)"
        qms = dg_stages["signal_degrade"][0]
        degrade = {"what": "zlib+svb-zd records -> qts round (3 bits) -> zlib+ex-zd records, device resident, checked against "
                           "(x + 4) & ~7 on the input samples", "reads": DG, "pass_ms": g0.elapsed_time(g1),
                   "reads_per_s": DG / (g0.elapsed_time(g1) * 1e-3), "bytes_per_read_out": dg_bytes / DG,
                   "size_vs_lossless_zlib_svbzd": (dg_bytes / DG) / (enc_bytes / R),
                   "qts_round_kernel_ms": qms, "qts_algorithmic_bytes": 4 * N * DG,
                   "qts_GBps": 4 * N * DG / (qms * 1e-3) / 1e9 if qms > 0 else None}
        del d_dg, d_dg_back, got, src, want

    if args.profile and not args.e2e_only:
        if rank == 0:
            print(json.dumps({"profile_only": True, "encode_ms": enc_ms, "decode_ms": dec_ms, "ms_per_step": ms_total / K,
                              "gpu_launches": int(launches), "stage_ms": stage_ms}), flush=True)
        cdc.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the host-buffer C-ABI call: pinned host slabs, all copies inside the timed region
    info = host_info()
    lw = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    need_gb = (R * rl + 2 * enc_cap + R * (rl + 8)) / 1e9 * lw
    Re = R
    if info.get("mem_available_gb") and need_gb > 0.5 * info["mem_available_gb"]:
        Re = max(1000, int(R * 0.5 * info["mem_available_gb"] / need_gb))     # pinned slabs must fit the host comfortably
    h_raw = pinned(torch, Re * rl)
    h_raw.copy_(d_raw[:Re * rl])
    h_enc = pinned(torch, Re * (rl // 2 + 512))
    h_back = pinned(torch, Re * (rl + 8) + 64)
    h_enc_off = np.zeros(Re + 1, np.uint64)

    # A step = one encode pass and one decode pass over the batch.  The two are independent calls (the decode pass reads the
    # image a previous encode pass produced -- same batch, same bytes every step), so they are issued together from two
    # host threads on two contexts: the encode pass is H2D-heavy, the decode pass D2H-heavy, and side by side they use both
    # directions of the PCIe link.  h_img holds the encoded image the decode pass reads; h_enc receives this step's image.
    from concurrent.futures import ThreadPoolExecutor
    cdc2 = s5.Codec(local)
    h_img = pinned(torch, Re * (rl // 2 + 512))
    _, nb0 = cdc.blow5_recode_batch_host(*ENC, h_raw, Re * rl, raw_off[:Re], raw_len[:Re], h_img, h_enc_off)
    img_zo = h_enc_off[:-1] + np.uint64(8)
    img_zl = (h_enc_off[1:] - h_enc_off[:-1] - np.uint64(8)).astype(np.uint32)
    pool = ThreadPoolExecutor(2)

    def enc_pass():
        return cdc.blow5_recode_batch_host(*ENC, h_raw, Re * rl, raw_off[:Re], raw_len[:Re], h_enc, None)[1]

    def dec_pass():
        return cdc2.blow5_recode_batch_host(*DEC, h_img, nb0, img_zo, img_zl, h_back, None)[1]

    def e2e_step():
        fe, fd = pool.submit(enc_pass), pool.submit(dec_pass)
        return fe.result(), fd.result()

    KE = max(2, min(K, 3))
    e2e_step()
    barrier()
    if args.e2e_only:   # dev: where the two pipelines spend their time while they run side by side
        cdc.stage_timing(True)
        cdc2.stage_timing(True)
        cdc.stage_report(reset=True)
        cdc2.stage_report(reset=True)
    t0 = time.perf_counter()
    for _ in range(KE):
        nb_e, nb_d = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_stages = None
    if args.e2e_only:
        e2e_stages = {"encode_pass_ms": {k: round(v[0] / KE, 2) for k, v in cdc.stage_report(reset=True).items() if v[0] > 0},
                      "decode_pass_ms": {k: round(v[0] / KE, 2) for k, v in cdc2.stage_report(reset=True).items() if v[0] > 0}}
        cdc.stage_timing(False)
        cdc2.stage_timing(False)
    hb = h_back[:Re * (rl + 8)].view(Re, rl + 8)
    assert nb_d == Re * (rl + 8) and torch.equal(hb[:, 8:], h_raw.view(Re, rl)), "e2e round trip failed"
    assert nb_e == nb0 and torch.equal(h_enc[:nb_e], h_img[:nb_e]), "e2e encode image changed between steps"
    assert torch.equal(h_enc[:nb_e], d_enc[:nb_e].cpu()) if Re == R else True, "e2e image differs from the device-resident one"
    # the same two passes one after the other (what a single `view` conversion sees)
    barrier()
    t0 = time.perf_counter()
    enc_pass()
    t1 = time.perf_counter()
    dec_pass()
    t2 = time.perf_counter()
    e2e_seq = max_over_ranks([t1 - t0, t2 - t1])
    pool.shutdown()
    cdc2.close()
    if args.e2e_only:
        if rank == 0:
            print(json.dumps({"e2e_only": True, "e2e_reads_per_s": Re * world * KE / max_over_ranks(e2e_s), "step_s": e2e_s / KE,
                              "seq_encode_s": e2e_seq[0], "seq_decode_s": e2e_seq[1], "stages": e2e_stages}), flush=True)
        cdc.close()
        if world > 1:
            dist.destroy_process_group()
        return
    e2e_s = max_over_ranks(e2e_s)
    h2d = Re * rl + nb_e               # records up (encode) + compressed records up (decode)
    d2h = nb_e + Re * (rl + 8)         # compressed image down + uncompressed image down
    up_gbs, dn_gbs = pcie_probe(torch, barrier)
    up_sum, dn_sum = sum_over_ranks(up_gbs), sum_over_ranks(dn_gbs)

    # ---- N > 1: one batch held by rank 0, split across the GPUs -- per-GPU PCIe copies vs NCCL scatter / gather
    split = None
    if world > 1 and not args.no_split:
        del h_back, h_img
        split = split_compare(torch, dist, cdc, synth, args, rank, world, min(Re, 262144), rl, barrier, max_over_ranks, shard_bounds)

    if rank == 0:
        peak, peak_src = load_peaks()
        # ---- roofline of the slowest kernel of the step
        # algorithmic bytes: the deflate kernel reads a packed record (M + C_svb) and writes C_rec; the
        # inflate kernel reads C_rec and writes the packed record.  The svb-zd intermediate is materialised in HBM
        # (separate kernels), so the whole pass moves 2N + M + C_rec + 2 (M + C_svb) per read: reported as pass_*.
        c_rec = enc_bytes / R - 8
        kern_bytes = (packed_bytes_per_read + c_rec) * R
        dom_stage = "record_press" if stage_ms["record_press"] >= stage_ms["record_depress"] else "record_depress"
        dom = STAGE_KERNELS[dom_stage]
        dom_ms = stage_ms[dom_stage]
        n_launch = max(1.0, stage_launches[dom_stage])
        achieved = kern_bytes / (dom_ms * 1e-3) / 1e9
        fused_bytes = (2 * N + 82 + c_rec + 8) * R
        cpu = cpu_baseline(h_raw[:min(Re, 40000) * rl].view(-1, rl).numpy()) if world == 1 else None
        out = {
            "metric": METRIC, "value": R * world * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config_dict(args),
            "raw_signal_GBps": 2 * N * R * world * K / (ms_total * 1e-3) / 1e9,
            "encode_ms": enc_ms, "decode_ms": dec_ms,
            "encode_reads_per_s": R * world / (enc_ms * 1e-3), "decode_reads_per_s": R * world / (dec_ms * 1e-3),
            "zlib_bytes_per_read": c_rec, "packed_bytes_per_read": packed_bytes_per_read,
            "zlib_ratio_on_record": c_rec / packed_bytes_per_read,
            "stage_ms_per_step": stage_ms,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_PER_100K[dom_stage] * (R / n_launch) / 1e5, "traffic_source": NCU_TRAFFIC_SOURCE,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kern_bytes / n_launch, "launches_per_step": n_launch,
                         "avg_launch_ms": dom_ms / n_launch,
                         "formula": "(M + C_svb + C_rec) per record: packed record in (out) + zlib stream out (in); "
                                    "the kernel is issue/latency bound (serial bit streams), the fraction is reported as required",
                         "deflate_frac": kern_bytes / (stage_ms["record_press"] * 1e-3) / 1e9 / peak,
                         "inflate_frac": kern_bytes / (stage_ms["record_depress"] * 1e-3) / 1e9 / peak,
                         "svbzd_encode_frac": (2 * N + packed_bytes_per_read - 82) * R / (stage_ms["signal_press"] * 1e-3) / 1e9 / peak,
                         "svbzd_decode_frac": (2 * N + packed_bytes_per_read - 82) * R / (stage_ms["signal_depress"] * 1e-3) / 1e9 / peak,
                         "encode_pass_frac_fused_bytes": fused_bytes / (enc_ms * 1e-3) / 1e9 / peak,
                         "decode_pass_frac_fused_bytes": fused_bytes / (dec_ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": Re * world * KE / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": KE, "reads_per_gpu": Re,
                    "api": "s5b_blow5_recode_batch_host x2 (encode pass and decode pass of a step issued together from two host "
                           "threads / contexts; pinned host slabs, 3-lane pipeline each)",
                    "sequential_encode_reads_per_s": Re * world / e2e_seq[0], "sequential_decode_reads_per_s": Re * world / e2e_seq[1],
                    "pcie_GBps_each_way": max(h2d, d2h) / (e2e_s / KE) / 1e9,
                    "platform_probe": {"what": "concurrent pinned H2D + D2H on every rank's link at once (512 MiB copies)",
                                       "h2d_GBps_sum": up_sum, "d2h_GBps_sum": dn_sum, "rank0_h2d_GBps": up_gbs,
                                       "rank0_d2h_GBps": dn_gbs},
                    "frac_of_platform_copy_ceiling": (max(h2d, d2h) * world / (e2e_s / KE) / 1e9) / max(1e-9, min(up_sum, dn_sum))},
            "gpu_launches": int(launches),
            "sharding": "reads split across ranks, no collective on the data path", "rank0_numa_binding": numa,
            "clocks": clocks,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        else:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                   "sample": "timed at N=1 only (see the --impl reference arm)"}
        if split is not None:
            out["split_compare"] = split
        if degrade is not None:
            if degrade["qts_GBps"]:
                degrade["qts_frac_of_hbm_peak"] = degrade["qts_GBps"] / peak
            out["degrade"] = degrade
        print(json.dumps(out), flush=True)
    cdc.close()
    if world > 1:
        dist.destroy_process_group()


def split_compare(torch, dist, cdc, synth, args, rank, world, T, rl, barrier, max_over_ranks, shard_bounds):
    """Strong-scaling view of SURVEY 8e: ONE batch of T records held by rank 0 (pinned host memory) is encoded by all
    GPUs.  (a) pcie: every rank copies its own byte-balanced shard over its own PCIe link (the shards are in host memory
    every rank can reach; here each rank uses the same bytes from its own pinned slab) and runs the host-form transcoder;
    (b) nccl: rank 0 moves the whole batch over ITS link, scatters variable-size shards with NCCL send/recv (sizes first,
    by all_gather), every rank transcodes device-resident, the images are gathered to rank 0 the same way and copied
    down.  Times are device/wall max over ranks; both produce the identical image."""
    # the ONE batch: every rank holds the same bytes in its own pinned slab (standing in for host memory all ranks of the
    # node can reach); rank 0's copy is "the" batch of the NCCL path
    sig = synth.nanopore_signal(T * args.samples, seed=4242, device="cuda")
    d_batch = synth.blow5_records(sig, T, args.samples, seed=4242).view(-1)
    del sig
    h_raw = pinned(torch, T * rl)
    h_raw.copy_(d_batch)
    del d_batch
    torch.cuda.synchronize()
    tab_off = np.arange(T, dtype=np.uint64) * np.uint64(rl)
    tab_len = np.full(T, rl, np.uint32)
    bounds = shard_bounds(tab_len, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    m = r1 - r0
    ENC = (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD)
    cap = m * (rl // 2 + 512) + 4096
    # (a) per-GPU PCIe
    h_out = pinned(torch, cap)
    cdc.blow5_recode_batch_host(*ENC, h_raw[r0 * rl:], m * rl, tab_off[:m], tab_len[r0:r1], h_out, None)
    barrier()
    t0 = time.perf_counter()
    _, nb_a = cdc.blow5_recode_batch_host(*ENC, h_raw[r0 * rl:], m * rl, tab_off[:m], tab_len[r0:r1], h_out, None)
    torch.cuda.synchronize()
    t_pcie = max_over_ranks(time.perf_counter() - t0)
    # (b) rank 0 link + NCCL scatter / gather
    d_shard = torch.empty(m * rl + 64, dtype=torch.uint8, device="cuda")
    d_img = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_all = torch.empty(T * rl, dtype=torch.uint8, device="cuda") if rank == 0 else None
    d_gather = torch.empty(T * (rl // 2 + 512) + 4096, dtype=torch.uint8, device="cuda") if rank == 0 else None
    h_final = pinned(torch, T * (rl // 2 + 512) + 4096) if rank == 0 else None
    sizes = torch.zeros(world, dtype=torch.int64, device="cuda")

    def nccl_pass():
        nccl_bytes = 0
        if rank == 0:
            d_all.copy_(h_raw[:T * rl], non_blocking=True)
        ops = []
        if rank == 0:
            d_shard[:m * rl].copy_(d_all[:m * rl])
            for p in range(1, world):
                a, b = bounds[p] * rl, bounds[p + 1] * rl
                ops.append(dist.P2POp(dist.isend, d_all[a:b], p))
                nccl_bytes += b - a
        else:
            ops.append(dist.P2POp(dist.irecv, d_shard[:m * rl], 0))
        for w_ in dist.batch_isend_irecv(ops) if ops else []:
            w_.wait()
        # the library works on its own stream: order it after torch's stream, and back
        torch.cuda.current_stream().synchronize()
        cdc.blow5_recode_dev(*ENC, d_shard, m * rl, tab_off[:m], tab_len[r0:r1], d_img, d_res, None)
        cdc.sync()
        mine = d_res[0:1].clone()
        got = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(got, mine)
        sz = [int(g) for g in got]
        ops = []
        if rank == 0:
            at = sz[0]
            d_gather[:at].copy_(d_img[:at])
            for p in range(1, world):
                ops.append(dist.P2POp(dist.irecv, d_gather[at:at + sz[p]], p))
                at += sz[p]
                nccl_bytes += sz[p]
        else:
            ops.append(dist.P2POp(dist.isend, d_img[:sz[rank]], 0))
        for w_ in dist.batch_isend_irecv(ops) if ops else []:
            w_.wait()
        total = sum(sz)
        if rank == 0:
            h_final[:total].copy_(d_gather[:total], non_blocking=True)
        torch.cuda.synchronize()
        return total, nccl_bytes, int(d_res[1])

    nccl_pass()
    barrier()
    t0 = time.perf_counter()
    total, nccl_bytes, err = nccl_pass()
    t_nccl = max_over_ranks(time.perf_counter() - t0)
    assert err == 0, "nccl split: transcoding failed"
    # both paths must give the same bytes: compare this rank's shard image
    assert torch.equal(h_out[:nb_a], d_img[:nb_a].cpu()), "nccl split: shard image differs from the pcie path"
    return {"total_reads": T, "direction": "encode (uncompressed records -> zlib+svb-zd image)", "shards": "byte-balanced, contiguous",
            "pcie_per_gpu_reads_per_s": T / t_pcie, "nccl_scatter_gather_reads_per_s": T / t_nccl,
            "nccl_data_plane_bytes": int(nccl_bytes), "pcie_s": t_pcie, "nccl_s": t_nccl}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1000000, help="reads per GPU")
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--profile", action="store_true", help="device-resident loop only (for runs under ncu)")
    ap.add_argument("--no-split", action="store_true", help="skip the N>1 pcie-vs-nccl split comparison")
    ap.add_argument("--e2e-only", action="store_true", help="dev: one device-resident step, then only the e2e timing (prints a short line)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
