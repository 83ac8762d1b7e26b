#!/usr/bin/env python
"""Benchmark of the SLA block encode/decode hot path on B200 (contract: see DESIGN.md section 7).

One "step" = one whole-file encode of the workload by libsla_b200.so.  The default workload is
BASELINE.json config 2: synthetic 16-bit stereo 44.1 kHz, 1 hour, preset 2 (PARCOR 16, long-term 1,
LMS 8, 12288-sample blocks, mid/side).  With N GPUs every rank encodes its own 1-hour file (files
shard across GPUs, no collective on the math path; sizes are all-gathered for the stitch table).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--seconds S] [--preset P] [--impl reference]

Rank 0 prints ONE JSON line.  `value` = whole-job M channel-samples/s with the PCM already resident
in HBM (CUDA events, max over ranks); `e2e` = the same through SLAEncoder_EncodeWhole with pinned HOST
buffers (H2D + kernels + D2H inside the timed region); `decode` reports the mirror path.
`--impl reference` times the unmodified reference (oracle/_ref/libsla_ref.so) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from sla_b200 import capi, synth  # noqa: E402

PRODUCT_SO = os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsla_ref.so")
METRIC = "encode_throughput"
UNIT = "M channel-samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seconds", type=int, default=3600, help="length of the synthetic file per GPU")
    ap.add_argument("--preset", type=int, default=2)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--bits", type=int, default=16)
    ap.add_argument("--rate", type=int, default=44100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"C2: synthetic {a.bits}-bit {a.channels}ch {a.rate} Hz, {a.seconds} s per GPU, preset {a.preset}"
            if (a.bits, a.channels, a.rate, a.preset) == (16, 2, 44100, 2) else
            f"synthetic {a.bits}-bit {a.channels}ch {a.rate} Hz, {a.seconds} s per GPU, preset {a.preset}")


# ----------------------------------------------------------------------------- CPU reference arm
def _ref_worker(job):
    """Encode+decode one tile with the unmodified reference; returns timings and the stream."""
    file_index, nch, nsamp, bits, rate, preset = job
    lib = capi.SLALibrary(REF_SO)
    pcm = synth.synth_pcm(nch, nsamp, bits, rate, file_index)
    ep = capi.preset_parameter(preset, nch)
    t0 = time.perf_counter()
    rc, data = lib.encode_whole(pcm, bits, rate, ep)
    t1 = time.perf_counter()
    rc2, dec, _ = lib.decode_whole(data)
    t2 = time.perf_counter()
    ok = rc == 0 and rc2 == 0 and np.array_equal(dec, pcm)
    return t1 - t0, t2 - t1, len(data), ok, data


def cpu_reference(a, tiles_per_core=1, tile_seconds=30, keep_streams=False):
    """All host cores, one process per core, each encoding `tiles_per_core` tiles of the workload."""
    if not os.path.exists(REF_SO):
        from oracle import binding
        binding.build("ref")
    cores = os.cpu_count() or 1
    nsamp = tile_seconds * a.rate
    jobs = [(1000 + i, a.channels, nsamp, a.bits, a.rate, a.preset) for i in range(cores * tiles_per_core)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    chsamp = len(jobs) * nsamp * a.channels
    enc_cpu = sum(r[0] for r in res)
    dec_cpu = sum(r[1] for r in res)
    out = dict(
        encode_all_core=chsamp / (enc_cpu / cores) / 1e6,      # cores run concurrently
        decode_all_core=chsamp / (dec_cpu / cores) / 1e6,
        encode_per_core=chsamp / enc_cpu / 1e6, decode_per_core=chsamp / dec_cpu / 1e6,
        cores=cores, wall=wall, ok=all(r[3] for r in res), bytes=sum(r[2] for r in res), chsamp=chsamp,
        sample=f"{len(jobs)} tiles x {tile_seconds} s of the workload signal, one process per core")
    if keep_streams:
        out["streams"] = [(j[0], r[4]) for j, r in zip(jobs, res)]
    return out


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    # each step = one bounded sample on all cores
    for _ in range(max(a.warmup, 0) and 1):
        cpu_reference(a, 1, 10)
    vals, t0 = [], time.perf_counter()
    for _ in range(a.steps):
        r = cpu_reference(a, 1, 30)
        vals.append(r)
    ms = 1e3 * (time.perf_counter() - t0) / max(a.steps, 1)
    v = statistics.mean(x["encode_all_core"] for x in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "input_flush": "n/a (CPU)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": vals[-1]["cores"], "kind": "reference",
                         "sample": vals[-1]["sample"]},
        "decode": {"value": statistics.mean(x["decode_all_core"] for x in vals), "unit": UNIT},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every few
    milliseconds (the timed region is only a few hundred ms long, too short for `nvidia-smi -lms`)."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40))

    def __init__(self, index, uuid=None):
        self.index, self.uuid = index, uuid
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self.thread, self.stop_flag, self.err = None, False, None
        self.recording = False

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + self.uuid).encode() if not self.uuid.startswith("GPU-") else self.uuid.encode())
                except Exception:
                    h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop_flag:
                if self.recording:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    try:
                        self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                time.sleep(0.003)
        except Exception as e:      # noqa: BLE001
            self.err = repr(e)

    def start(self):
        """spawn the poller (NVML initialisation takes longer than the timed region: do it early)"""
        import threading
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def begin(self):
        self.recording = True

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable: %s" % self.err]}
        reasons = sorted(name for name, bit in self.REASONS if self.reason_bits & bit)
        return {"sm_mhz": statistics.median(self.samples), "sm_min_mhz": min(self.samples), "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": reasons}


def bind_extras(L):
    L.SLAB200_Encoder_EncodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.SLAB200_Decoder_DecodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.SLAB200_Encoder_LastTiming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.SLAB200_Decoder_LastTiming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.SLAB200_Encoder_EnableProfile.argtypes = [C.c_void_p, C.c_int]
    L.SLAB200_Decoder_EnableProfile.argtypes = [C.c_void_p, C.c_int]
    L.SLAB200_Encoder_GetProfile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.SLAB200_Encoder_GetProfile.restype = C.c_uint32
    L.SLAB200_Decoder_GetProfile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.SLAB200_Decoder_GetProfile.restype = C.c_uint32
    L.SLAB200_LastError.restype = C.c_char_p


def get_profile(fn, handle):
    names = (C.c_char_p * 64)()
    ms = (C.c_float * 64)()
    n = fn(handle, names, ms, 64)
    return [(names[i].decode(), float(ms[i])) for i in range(n)]


def run_gpu_arm(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - libsla_b200.so has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if not os.path.exists(PRODUCT_SO):
        raise SystemExit("bench.py: sla_b200/lib/libsla_b200.so missing - run `make product`")
    lib = capi.SLALibrary(PRODUCT_SO)
    L = lib.lib
    bind_extras(L)

    nch, bits, rate = a.channels, a.bits, a.rate
    n = a.seconds * rate
    chsamp = n * nch
    ep = capi.preset_parameter(a.preset, nch)

    # ---- workload: pinned host PCM (for e2e) and a resident device copy (for value) ----
    h_pcm_t = torch.empty((nch, n), dtype=torch.int32, pin_memory=True)
    h_pcm = h_pcm_t.numpy()
    synth.synth_long(nch, n, bits, rate, file_index=rank, out=h_pcm)
    d_pcm = h_pcm_t.to(dev, non_blocking=False)
    cap = 43 + int(chsamp * max(bits // 8, 1) * 1.25) + (1 << 20)
    d_stream = torch.zeros(cap, dtype=torch.uint8, device=dev)
    h_stream_t = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    d_dec = torch.empty((nch, n), dtype=torch.int32, device=dev)
    h_dec_t = torch.empty((nch, n), dtype=torch.int32, pin_memory=True)

    enc_cfg = capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)
    dec_cfg = capi.DecoderConfig(**capi.CLI_CAPACITY, enable_crc_check=1, verpose_flag=0)
    enc = L.SLAEncoder_Create(C.byref(enc_cfg))
    dec = L.SLADecoder_Create(C.byref(dec_cfg))
    if not enc or not dec:
        raise SystemExit("bench.py: handle creation failed: " + (L.SLAB200_LastError() or b"").decode())
    wf = capi.WaveFormat(nch, bits, rate, 0)
    assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0
    assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == 0

    def ptr_array(t):
        arr = (C.c_void_p * nch)()
        for c in range(nch):
            arr[c] = t[c].data_ptr()
        return arr
    d_in_ptrs, d_out_ptrs = ptr_array(d_pcm), ptr_array(d_dec)
    h_in_ptrs, h_out_ptrs = ptr_array(h_pcm_t), ptr_array(h_dec_t)
    size = C.c_uint32(0)
    got = C.c_uint32(0)
    ms3 = (C.c_float * 3)()
    nl = C.c_uint32(0)

    def enc_device():
        rc = L.SLAB200_Encoder_EncodeWholeDevice(enc, d_in_ptrs, n, d_stream.data_ptr(), cap, C.byref(size))
        if rc != 0:
            raise SystemExit(f"encode failed rc={rc}: " + (L.SLAB200_LastError() or b"").decode())

    def dec_device():
        rc = L.SLAB200_Decoder_DecodeWholeDevice(dec, d_stream.data_ptr(), size.value, d_out_ptrs, n, C.byref(got))
        if rc != 0:
            raise SystemExit(f"decode failed rc={rc}: " + (L.SLAB200_LastError() or b"").decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    sampler.start()

    # ---- warm-up ----
    for _ in range(a.warmup):
        enc_device()
    dec_device()
    stream_bytes = size.value
    exact = bool(torch.equal(d_dec[:, :got.value], d_pcm)) and got.value == n

    # ---- per-kernel table: one profiled pass per step (CUDA events around every launch; the profiled
    # call runs the file as a single pass so that every kernel appears once) ----
    L.SLAB200_Encoder_EnableProfile(enc, 1)
    kern_ms = {}
    single_pass_ms = 0.0
    for _ in range(a.steps):
        enc_device()
        L.SLAB200_Encoder_LastTiming(enc, ms3, C.byref(nl))
        single_pass_ms += ms3[0] + ms3[1] + ms3[2]
        for name, ms in get_profile(L.SLAB200_Encoder_GetProfile, enc):
            kern_ms[name] = kern_ms.get(name, 0.0) + ms
    L.SLAB200_Encoder_EnableProfile(enc, 0)
    single_pass_ms = max_over_ranks(single_pass_ms / a.steps)
    for _ in range(2):
        enc_device()            # back to the unprofiled path

    # ---- timed region 1: device-resident encode (value) ----
    launches = 0
    barrier()
    sampler.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    lib_ms = 0.0
    for _ in range(a.steps):
        enc_device()
        L.SLAB200_Encoder_LastTiming(enc, ms3, C.byref(nl))
        lib_ms += ms3[0] + ms3[1] + ms3[2]
        launches += nl.value
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    # device time of the library's own stream (CUDA events recorded on that stream inside the call)
    step_ms = max_over_ranks(lib_ms / a.steps)
    wall_step_ms = max_over_ranks(wall_ms / a.steps)
    value = world * chsamp / (step_ms * 1e-3) / 1e6

    # ---- timed region 2: device-resident decode ----
    L.SLAB200_Decoder_EnableProfile(dec, 1)
    dec_kern_ms = {}
    for _ in range(2):
        dec_device()
    barrier()
    dlib_ms = 0.0
    for _ in range(a.steps):
        dec_device()
        L.SLAB200_Decoder_LastTiming(dec, ms3, C.byref(nl))
        dlib_ms += ms3[0] + ms3[1] + ms3[2]
        launches += nl.value
        for name, ms in get_profile(L.SLAB200_Decoder_GetProfile, dec):
            dec_kern_ms[name] = dec_kern_ms.get(name, 0.0) + ms
    barrier()
    L.SLAB200_Decoder_EnableProfile(dec, 0)
    dec_step_ms = max_over_ranks(dlib_ms / a.steps)
    dec_value = world * chsamp / (dec_step_ms * 1e-3) / 1e6

    # ---- timed region 3: end to end through the reference-facing C API with HOST buffers ----
    def enc_host():
        rc = L.SLAEncoder_EncodeWhole(enc, h_in_ptrs, n, h_stream_t.data_ptr(), cap, C.byref(size))
        if rc != 0:
            raise SystemExit(f"host encode failed rc={rc}")

    def dec_host():
        rc = L.SLADecoder_DecodeWhole(dec, h_stream_t.data_ptr(), size.value, h_out_ptrs, n, C.byref(got))
        if rc != 0:
            raise SystemExit(f"host decode failed rc={rc}")
    enc_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        enc_host()
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / a.steps)
    e2e_value = world * chsamp / (e2e_ms * 1e-3) / 1e6
    dec_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        dec_host()
    barrier()
    e2e_dec_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / a.steps)
    host_exact = bool(np.array_equal(h_dec_t.numpy()[:, :got.value], h_pcm)) and got.value == n

    # ---- timed region 4: the raw-PCM entry points (interleaved little-endian PCM in pinned host memory) ----
    pcm_leg = None
    if bits in (8, 16, 24, 32):
        L.SLAB200_Encoder_EncodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.SLAB200_Decoder_DecodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        fb = nch * bits // 8
        h_raw_t = torch.empty(n * fb, dtype=torch.uint8, pin_memory=True)
        h_raw_t.numpy()[:] = np.frombuffer(capi.planar_to_pcm(h_pcm, bits), dtype=np.uint8)
        h_back_t = torch.empty(n * fb, dtype=torch.uint8, pin_memory=True)
        size_pcm = C.c_uint32(0)

        def enc_pcm():
            rc = L.SLAB200_Encoder_EncodePCM(enc, h_raw_t.data_ptr(), n, h_stream_t.data_ptr(), cap, C.byref(size_pcm))
            if rc != 0:
                raise SystemExit(f"PCM encode failed rc={rc}")

        def dec_pcm():
            rc = L.SLAB200_Decoder_DecodePCM(dec, h_stream_t.data_ptr(), size_pcm.value, h_back_t.data_ptr(), n, C.byref(got))
            if rc != 0:
                raise SystemExit(f"PCM decode failed rc={rc}")
        enc_pcm()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            enc_pcm()
        barrier()
        pcm_enc_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / a.steps)
        pcm_same = bool(torch.equal(h_stream_t[:size_pcm.value].to(dev), d_stream[:size_pcm.value])) and size_pcm.value == stream_bytes
        dec_pcm()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            dec_pcm()
        barrier()
        pcm_dec_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / a.steps)
        pcm_exact = bool(torch.equal(h_back_t, h_raw_t)) and got.value == n
        pcm_leg = {"value": world * chsamp / (pcm_enc_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": pcm_enc_ms,
                   "h2d_bytes_per_step": n * fb, "d2h_bytes_per_step": stream_bytes,
                   "decode_value": world * chsamp / (pcm_dec_ms * 1e-3) / 1e6, "decode_ms_per_step": pcm_dec_ms,
                   "stream_equals_device_stream": pcm_same, "roundtrip": pcm_exact,
                   "api": "SLAB200_Encoder_EncodePCM / SLAB200_Decoder_DecodePCM (interleaved little-endian PCM, host)"}
        del h_raw_t, h_back_t
    host_same_as_device = bool(torch.equal(h_stream_t[:size.value].to(dev), d_stream[:size.value])) and size.value == stream_bytes

    # ---- stitch metadata across ranks (the only collective: sizes -> offsets) ----
    sizes = [stream_bytes]
    if world > 1:
        t = torch.zeros(world, dtype=torch.int64, device=dev)
        t[rank] = stream_bytes
        dist.all_reduce(t)
        sizes = [int(x) for x in t.tolist()]
        flags = torch.tensor([int(exact), int(host_exact), int(host_same_as_device)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        exact, host_exact, host_same_as_device = (bool(x) for x in flags.tolist())

    if rank == 0:
        c = stream_bytes / chsamp                      # encoded bytes per channel-sample on this workload
        b_enc = 4.0 + c                                 # int32-planar C-API path
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        top_name, top_ms = max(kern_ms.items(), key=lambda kv: kv[1])
        top_ms /= a.steps
        achieved = b_enc * chsamp / (top_ms * 1e-3) / 1e9
        kernels_ms_total = sum(kern_ms.values()) / a.steps
        # DRAM traffic per launch of each kernel from the committed `ncu --set full` capture of this
        # workload (profiles/traffic_r01.json, written by tools/ncu_traffic.py); None when the capture
        # was taken at another size
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic_r01.json")) as f:
                tj = json.load(f)
            if tj.get("channel_samples") == chsamp:
                traffic = tj.get("dram_bytes_per_launch", {})
        except Exception:
            pass
        # the kernels that only stream: algorithmic bytes per channel-sample each moves
        stream_defs = {"E0 k_enc_scan": 4.0, "E9 k_enc_pack": 4.0 + 2.0 + c, "E10 k_enc_crc": c,
                       "D1a k_dec_crc": c, "D3 k_dec_output": 8.0}
        streaming = {}
        for name, bpcs in stream_defs.items():
            ms = (kern_ms.get(name) or dec_kern_ms.get(name))
            if ms:
                ms /= a.steps
                gbs = bpcs * chsamp / (ms * 1e-3) / 1e9
                streaming[name] = {"ms": ms, "algorithmic_bytes_per_channel_sample": bpcs, "achieved_gbs": gbs,
                                   "frac_of_measured_hbm": gbs / peak,
                                   "traffic": traffic.get(name.split()[-1])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32/f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "channel_samples_per_gpu": chsamp,
                       "blocks": None, "input_flush": "inputs (%.2f GB per GPU) larger than L2" % (chsamp * 4 / 1e9),
                       "timer": "CUDA events on the library stream around the whole call (H2D-less device path), max over ranks",
                       "single_pass_ms_per_step": single_pass_ms,
                       "wall_ms_per_step": wall_step_ms},
            "pcm_mb_per_s": value * bits / 8,
            "compressed_bytes_per_channel_sample": c,
            "compression_ratio": stream_bytes / (chsamp * bits / 8),
            "decode": {"value": dec_value, "unit": UNIT, "ms_per_step": dec_step_ms,
                       "kernels_ms": {k: v / a.steps for k, v in dec_kern_ms.items()}},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": chsamp * 4, "d2h_bytes_per_step": stream_bytes,
                    "decode_value": world * chsamp / (e2e_dec_ms * 1e-3) / 1e6, "decode_ms_per_step": e2e_dec_ms},
            "roofline": {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic.get(top_name.split()[-1]),
                         "algorithmic_bytes_per_channel_sample": b_enc, "kernel_ms": top_ms,
                         "kernel_share_of_step": top_ms / kernels_ms_total,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "whole_step_frac": b_enc * chsamp / (step_ms * 1e-3) / 1e9 / peak},
            "e2e_pcm": pcm_leg,
            "streaming_kernels": streaming,
            "kernels_ms": {k: v / a.steps for k, v in kern_ms.items()},
            "gpu_launches": launches,
            "bit_exact": {"gpu_roundtrip": exact, "host_api_roundtrip": host_exact,
                          "host_api_stream_equals_device_stream": host_same_as_device},
            "stitch": {"sizes": sizes, "offsets": [43 + sum(sizes[:i]) - 43 * i for i in range(len(sizes))]},
            "clocks": clocks,
        }
        if not a.no_cpu_baseline and world == 1:      # reported baseline: rank 0 at N = 1 only
            r = cpu_reference(a, 1, 30, keep_streams=True)
            # byte-identity of GPU streams vs the reference on the same tiles
            same = 0
            for file_index, want in r["streams"]:
                pcm = synth.synth_pcm(nch, 30 * rate, bits, rate, file_index)
                rc, mine = lib.encode_whole(pcm, bits, rate, ep)
                same += int(rc == 0 and mine == want)
            line["cpu_baseline"] = {"value": r["encode_all_core"], "unit": UNIT, "cores": r["cores"],
                                    "kind": "reference", "sample": r["sample"],
                                    "per_core": r["encode_per_core"], "decode_all_core": r["decode_all_core"],
                                    "decode_per_core": r["decode_per_core"], "roundtrip_ok": r["ok"]}
            line["bit_exact"]["byte_identical_to_reference"] = f"{same}/{len(r['streams'])} sample files"
        emit(line)
    L.SLAEncoder_Destroy(enc)
    L.SLADecoder_Destroy(dec)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version
    banner, for one) has been redirected to stderr"""
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
        return
    run_gpu_arm(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
