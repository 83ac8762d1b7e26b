#!/usr/bin/env python
"""Benchmark of the SLA block encode/decode hot path on B200 (contract: see DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--configs C3,C4,C5,strong] [--impl reference]

The headline workload is BASELINE.json config 2 (C2): synthetic 16-bit stereo 44.1 kHz, 1 hour, preset 2.
One "step" = one whole-file encode by libsla_b200.so; with N GPUs every rank encodes its own file
(files shard across GPUs, no collective on the math path).  Rank 0 prints ONE JSON line:

  value     whole-job M channel-samples/s with the PCM resident in HBM (library CUDA events, max over ranks)
  e2e       the same through SLAEncoder_EncodeWhole with page-locked HOST buffers (H2D + kernels + D2H timed);
            e2e_pageable: the same from malloc'ed memory; e2e_pcm: the raw-PCM entry points
  decode    the mirror path
  configs   sub-records for BASELINE configs 3, 4 and 5 (workloads: sla_b200/workloads.py), each with value,
            e2e, its own CPU baseline, byte identity against the reference on a sample, and a full-size
            stream decoded by the REFERENCE decoder
  strong    one file split over all GPUs (chain hand-off + NCCL metadata collectives), stitched stream
            compared with the single-GPU stream

`--impl reference` times the unmodified reference (oracle/_ref/libsla_ref.so) on all host cores on ranges of
the same C2 file.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# before CUDA is initialised (by torch): one hardware work queue per stream of the pipelined calls
# (INTEGRATION.md section 4); libsla_b200.so sets the same default when it is the first CUDA user
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from sla_b200 import capi, workloads  # noqa: E402

PRODUCT_SO = os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsla_ref.so")
METRIC = "encode_throughput"
UNIT = "M channel-samples/s"
SUB_STEPS = 3            # timed steps of the sub-configs (C3, C4, C5, strong); the headline uses --steps


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seconds", type=int, default=None, help="shorten the hour-long files (development only)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--configs", default="C3,C4,C5,strong", help="sub-records to run next to C2 ('' = none)")
    ap.add_argument("--corpus-files", type=int, default=None, help="C5 corpus size (development only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ============================================================================= CPU reference legs
_TILES = None          # list of (pcm planar int32, bits, rate, preset); inherited by the forked workers


def _ref_worker(i):
    """Encode + decode tile i with the unmodified reference."""
    pcm, bits, rate, preset = _TILES[i]
    lib = capi.SLALibrary(REF_SO)
    ep = capi.preset_parameter(preset, pcm.shape[0])
    t0 = time.perf_counter()
    rc, data = lib.encode_whole(pcm, bits, rate, ep)
    t1 = time.perf_counter()
    rc2, dec, _ = lib.decode_whole(data)
    t2 = time.perf_counter()
    ok = rc == 0 and rc2 == 0 and np.array_equal(dec, pcm)
    return t1 - t0, t2 - t1, ok, data


def cpu_reference(tiles, keep_streams=False, what=""):
    """All host cores, one forked process per core, the reference's EncodeWhole + DecodeWhole on each tile."""
    global _TILES
    if not os.path.exists(REF_SO):
        from oracle import binding
        binding.build("ref")
    cores = os.cpu_count() or 1
    _TILES = tiles
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(min(cores, len(tiles))) as pool:
        res = pool.map(_ref_worker, range(len(tiles)), chunksize=1)
    wall = time.perf_counter() - t0
    _TILES = None
    chsamp = sum(t[0].size for t in tiles)
    enc_cpu = sum(r[0] for r in res)
    dec_cpu = sum(r[1] for r in res)
    used = min(cores, len(tiles))
    out = dict(
        encode_all_core=chsamp / (enc_cpu / used) / 1e6 * (cores / used),      # cores run concurrently
        decode_all_core=chsamp / (dec_cpu / used) / 1e6 * (cores / used),
        encode_per_core=chsamp / enc_cpu / 1e6, decode_per_core=chsamp / dec_cpu / 1e6,
        cores=cores, wall=wall, ok=all(r[2] for r in res), chsamp=chsamp, sample=what)
    if keep_streams:
        out["streams"] = [r[3] for r in res]
    return out


def file_tiles(pcm, bits, rate, preset, max_block, count, tile_seconds, shift=0):
    """`count` ranges of the file, each starting on a block boundary, copied out of (possibly page-locked)
    memory so that forked workers can read them"""
    n = pcm.shape[1]
    ranges = workloads.sample_ranges(n, max_block, count, tile_seconds * rate)
    if shift:
        ranges = [((s + shift * max_block) % max(1, n - ln + 1) // max_block * max_block, ln) for s, ln in ranges]
    return [(np.array(pcm[:, s:s + ln], dtype=np.int32, copy=True, order="C"), bits, rate, preset) for s, ln in ranges], ranges


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    c = workloads.CONFIGS["C2"]
    nch, bits, rate, preset = c["channels"], c["bits"], c["rate"], c["preset"]
    maxblk = capi.PRESETS[preset]["max_block"]
    pcm = workloads.long_file("C2", 0, seconds=a.seconds)
    cores = os.cpu_count() or 1
    for _ in range(1 if a.warmup > 0 else 0):
        tiles, _ = file_tiles(pcm, bits, rate, preset, maxblk, cores, 10)
        cpu_reference(tiles)
    vals, t0 = [], time.perf_counter()
    for s in range(a.steps):
        tiles, _ = file_tiles(pcm, bits, rate, preset, maxblk, cores, 30, shift=7 * s)
        vals.append(cpu_reference(tiles, what=f"{cores} ranges x 30 s of the C2 file (block-aligned starts), one process per core"))
    ms = 1e3 * (time.perf_counter() - t0) / max(a.steps, 1)
    v = statistics.mean(x["encode_all_core"] for x in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
        "config": workloads.describe("C2", world, **({"seconds": a.seconds} if a.seconds else {})),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": vals[-1]["cores"], "kind": "reference",
                         "sample": vals[-1]["sample"], "roundtrip_ok": all(x["ok"] for x in vals)},
        "decode": {"value": statistics.mean(x["decode_all_core"] for x in vals), "unit": UNIT},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ============================================================================= GPU arm helpers
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
        import threading
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def begin(self):
        self.recording = True

    def pause(self):
        self.recording = False

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
    u32p = C.POINTER(C.c_uint32)
    L.SLAB200_Encoder_EncodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, u32p]
    L.SLAB200_Decoder_DecodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, u32p]
    L.SLAB200_Encoder_EncodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, u32p]
    L.SLAB200_Decoder_DecodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, u32p]
    L.SLAB200_Encoder_EncodeBatchPCM.argtypes = [C.c_void_p, C.POINTER(capi.EncodeItem), C.c_uint32]
    L.SLAB200_Decoder_DecodeBatchPCM.argtypes = [C.c_void_p, C.POINTER(capi.BatchItem), C.c_uint32]
    L.SLAB200_Decoder_LastBatchTiming.argtypes = [C.c_void_p, C.POINTER(C.c_float), u32p]
    L.SLAB200_Encoder_LastTiming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.SLAB200_Encoder_LastFallbacks.argtypes = [C.c_void_p, C.c_void_p]
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


class Codec:
    """one encoder + one decoder handle of libsla_b200.so for a wave format and a preset"""

    def __init__(self, L, nch, bits, rate, preset):
        self.L, self.nch, self.bits, self.rate, self.preset = L, nch, bits, rate, preset
        self.ep = capi.preset_parameter(preset, nch)
        enc_cfg = capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)
        dec_cfg = capi.DecoderConfig(**capi.CLI_CAPACITY, enable_crc_check=1, verpose_flag=0)
        self.enc = L.SLAEncoder_Create(C.byref(enc_cfg))
        self.dec = L.SLADecoder_Create(C.byref(dec_cfg))
        if not self.enc or not self.dec:
            raise SystemExit("bench.py: handle creation failed: " + (L.SLAB200_LastError() or b"").decode())
        wf = capi.WaveFormat(nch, bits, rate, 0)
        assert L.SLAEncoder_SetWaveFormat(self.enc, C.byref(wf)) == 0
        assert L.SLAEncoder_SetEncodeParameter(self.enc, C.byref(self.ep)) == 0
        self.size, self.got = C.c_uint32(0), C.c_uint32(0)
        self._ms3, self._nl = (C.c_float * 3)(), C.c_uint32(0)
        self.launches = 0

    def close(self):
        self.L.SLAEncoder_Destroy(self.enc)
        self.L.SLADecoder_Destroy(self.dec)

    def _fail(self, what, rc):
        raise RuntimeError(f"{what} failed rc={rc}: " + (self.L.SLAB200_LastError() or b"").decode())

    @staticmethod
    def ptrs(t):
        arr = (C.c_void_p * t.shape[0])()
        for c in range(t.shape[0]):
            arr[c] = t[c].data_ptr()
        return arr

    def enc_device(self, d_ptrs, n, d_stream, cap):
        rc = self.L.SLAB200_Encoder_EncodeWholeDevice(self.enc, d_ptrs, n, d_stream.data_ptr(), cap, C.byref(self.size))
        if rc != 0:
            self._fail("device encode", rc)
        self.L.SLAB200_Encoder_LastTiming(self.enc, self._ms3, C.byref(self._nl))
        self.launches += self._nl.value
        return self._ms3[0] + self._ms3[1] + self._ms3[2]

    def dec_device(self, d_stream, size, d_out_ptrs, n):
        rc = self.L.SLAB200_Decoder_DecodeWholeDevice(self.dec, d_stream.data_ptr(), size, d_out_ptrs, n, C.byref(self.got))
        if rc != 0:
            self._fail("device decode", rc)
        self.L.SLAB200_Decoder_LastTiming(self.dec, self._ms3, C.byref(self._nl))
        self.launches += self._nl.value
        return self._ms3[0] + self._ms3[1] + self._ms3[2]

    def enc_host(self, h_ptrs, n, h_stream, cap):
        rc = self.L.SLAEncoder_EncodeWhole(self.enc, h_ptrs, n, h_stream.data_ptr(), cap, C.byref(self.size))
        if rc != 0:
            self._fail("host encode", rc)

    def dec_host(self, h_stream, size, h_out_ptrs, n):
        rc = self.L.SLADecoder_DecodeWhole(self.dec, h_stream.data_ptr(), size, h_out_ptrs, n, C.byref(self.got))
        if rc != 0:
            self._fail("host decode", rc)

    def enc_pcm(self, h_raw, n, h_stream, cap):
        rc = self.L.SLAB200_Encoder_EncodePCM(self.enc, h_raw.data_ptr(), n, h_stream.data_ptr(), cap, C.byref(self.size))
        if rc != 0:
            self._fail("PCM encode", rc)

    def dec_pcm(self, h_stream, size, h_back, n):
        rc = self.L.SLAB200_Decoder_DecodePCM(self.dec, h_stream.data_ptr(), size, h_back.data_ptr(), n, C.byref(self.got))
        if rc != 0:
            self._fail("PCM decode", rc)


class Dist:
    """rank plumbing: barrier + device synchronise, max over ranks, all-ranks-agree"""

    def __init__(self, rank, world, dev):
        self.rank, self.world, self.dev = rank, world, dev

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return x
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if self.world == 1:
            return x
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t)
        return float(t.item())

    def all_true(self, flag):
        if self.world == 1:
            return bool(flag)
        import torch
        import torch.distributed as dist
        t = torch.tensor([int(bool(flag))], device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(self, fn, steps):
        """wall time per step of fn(), bracketed by barrier + synchronise, max over ranks"""
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return self.max(1e3 * (time.perf_counter() - t0) / steps)


def pinned(shape, dtype):
    import torch
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def spawn_ref_decode(stream_bytes_view, name, rank, index, seconds):
    """background: the reference decoder over a whole GPU-encoded stream (tools/ref_decode_check.py)"""
    if not os.path.exists(REF_SO):
        return None
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, f"sla_b200_bench_{os.getpid()}_{name}.sla")
    np.asarray(stream_bytes_view).tofile(path)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "ref_decode_check.py"), path, name, str(rank), str(index)]
    if seconds:
        cmd.append(str(seconds))
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)


def join_ref_decode(proc, timeout=600):
    if proc is None:
        return {"ok": None, "note": "oracle/_ref/libsla_ref.so not built"}
    try:
        out, _ = proc.communicate(timeout=timeout)
        return json.loads(out.decode().strip().splitlines()[-1])
    except Exception as e:      # noqa: BLE001
        proc.kill()
        return {"ok": None, "note": "reference decode did not finish: %r" % (e,)}


# ============================================================================= one long file per GPU (C2, C3)
def leg_long_file(name, a, D, L, steps, warmup, headline, state):
    """Encode / decode of one hour-long file per GPU.  `headline` adds the per-kernel table, the clock
    sampler, the pageable and raw-PCM end-to-end legs."""
    import torch
    c = workloads.CONFIGS[name]
    nch, bits, rate, preset = c["channels"], c["bits"], c["rate"], c["preset"]
    seconds = a.seconds or c["seconds"]
    n = seconds * rate
    chsamp = n * nch
    dev, rank, world = D.dev, D.rank, D.world
    codec = Codec(L, nch, bits, rate, preset)
    maxblk = codec.ep.max_num_block_samples

    h_pcm_t = pinned((nch, n), torch.int32)
    h_pcm = h_pcm_t.numpy()
    t0 = time.perf_counter()
    workloads.long_file(name, rank, out=h_pcm, seconds=seconds)
    log(f"[{name}] rank {rank}: synthesised {chsamp / 1e6:.0f} M channel-samples in {time.perf_counter() - t0:.1f} s")
    d_pcm = h_pcm_t.to(dev)
    cap = 43 + int(chsamp * max(bits // 8, 1) * 1.25) + (1 << 20)
    d_stream = torch.zeros(cap, dtype=torch.uint8, device=dev)
    h_stream_t = pinned(cap, torch.uint8)
    d_dec = torch.empty((nch, n), dtype=torch.int32, device=dev)
    h_dec_t = pinned((nch, n), torch.int32)
    d_in, d_out = Codec.ptrs(d_pcm), Codec.ptrs(d_dec)
    h_in, h_out = Codec.ptrs(h_pcm_t), Codec.ptrs(h_dec_t)
    torch.cuda.synchronize()              # the library's streams are not ordered against torch's (sla_b200.h)

    # ---- warm-up + device round trip ----
    for _ in range(max(warmup, 1)):
        codec.enc_device(d_in, n, d_stream, cap)
    stream_bytes = codec.size.value
    codec.dec_device(d_stream, stream_bytes, d_out, n)
    exact = bool(torch.equal(d_dec[:, :codec.got.value], d_pcm)) and codec.got.value == n

    # the full-size stream goes to the reference decoder in the background (rank 0)
    bg = None
    if rank == 0 and not a.no_cpu_baseline:
        bg = spawn_ref_decode(d_stream[:stream_bytes].cpu().numpy(), name, rank, 0, a.seconds)

    # ---- per-kernel table (headline only): rank 0 alone, the other ranks idle, so that host contention
    # between ranks does not leak into the per-launch events ----
    kern_ms, dec_kern_ms, single_pass_ms = {}, {}, None
    if headline:
        D.barrier()
        if rank == 0:
            L.SLAB200_Encoder_EnableProfile(codec.enc, 1)
            L.SLAB200_Decoder_EnableProfile(codec.dec, 1)
            tot = 0.0
            for _ in range(steps):
                tot += codec.enc_device(d_in, n, d_stream, cap)
                for kname, ms in get_profile(L.SLAB200_Encoder_GetProfile, codec.enc):
                    kern_ms[kname] = kern_ms.get(kname, 0.0) + ms / steps
                codec.dec_device(d_stream, stream_bytes, d_out, n)
                for kname, ms in get_profile(L.SLAB200_Decoder_GetProfile, codec.dec):
                    dec_kern_ms[kname] = dec_kern_ms.get(kname, 0.0) + ms / steps
            single_pass_ms = tot / steps
            L.SLAB200_Encoder_EnableProfile(codec.enc, 0)
            L.SLAB200_Decoder_EnableProfile(codec.dec, 0)
            for _ in range(2):
                codec.enc_device(d_in, n, d_stream, cap)
        D.barrier()

    # ---- timed region 1: device-resident encode (value) ----
    codec.launches = 0
    sampler = state.get("sampler")
    D.barrier()
    if headline and sampler:
        sampler.begin()
    t0 = time.perf_counter()
    lib_ms = 0.0
    for _ in range(steps):
        lib_ms += codec.enc_device(d_in, n, d_stream, cap)
    D.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    if headline and sampler:
        sampler.pause()
    step_ms = D.max(lib_ms / steps)
    wall_step_ms = D.max(wall_ms / steps)
    enc_launches = codec.launches
    fb = (C.c_uint32 * 3)()
    L.SLAB200_Encoder_LastFallbacks(codec.enc, fb)

    # ---- timed region 2: device-resident decode ----
    for _ in range(2):
        codec.dec_device(d_stream, stream_bytes, d_out, n)
    D.barrier()
    dlib_ms = 0.0
    codec.launches = 0
    for _ in range(steps):
        dlib_ms += codec.dec_device(d_stream, stream_bytes, d_out, n)
    D.barrier()
    dec_step_ms = D.max(dlib_ms / steps)
    dec_launches = codec.launches

    # ---- timed region 3: end to end through the reference-facing C API, page-locked host buffers ----
    codec.enc_host(h_in, n, h_stream_t, cap)
    e2e_ms = D.timed(lambda: codec.enc_host(h_in, n, h_stream_t, cap), steps)
    host_size = codec.size.value
    codec.dec_host(h_stream_t, host_size, h_out, n)
    e2e_dec_ms = D.timed(lambda: codec.dec_host(h_stream_t, host_size, h_out, n), steps)
    host_exact = bool(np.array_equal(h_dec_t.numpy()[:, :codec.got.value], h_pcm)) and codec.got.value == n
    host_same = host_size == stream_bytes and bool(torch.equal(h_stream_t[:host_size].to(dev), d_stream[:host_size]))

    rec = {
        "config": workloads.describe(name, world, **({"seconds": a.seconds} if a.seconds else {})),
        "metric": METRIC, "unit": UNIT, "steps": steps, "warmup": max(warmup, 1),
        "value": world * chsamp / (step_ms * 1e-3) / 1e6, "ms_per_step": step_ms,
        "wall_ms_per_step": wall_step_ms,
        "pcm_mb_per_s": world * chsamp / (step_ms * 1e-3) / 1e6 * bits / 8,
        "compressed_bytes_per_channel_sample": stream_bytes / chsamp,
        "compression_ratio": stream_bytes / (chsamp * bits / 8),
        "decode": {"value": world * chsamp / (dec_step_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": dec_step_ms},
        "e2e": {"value": world * chsamp / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": chsamp * 4, "d2h_bytes_per_step": stream_bytes,
                "decode_value": world * chsamp / (e2e_dec_ms * 1e-3) / 1e6, "decode_ms_per_step": e2e_dec_ms,
                "host_memory": "page-locked", "api": "SLAEncoder_EncodeWhole / SLADecoder_DecodeWhole"},
        "gpu_launches": enc_launches + dec_launches,
        "bit_exact": {"gpu_roundtrip": D.all_true(exact), "host_api_roundtrip": D.all_true(host_exact),
                      "host_api_stream_equals_device_stream": D.all_true(host_same)},
        "exactness_fallbacks": {"block_channels": None, "reference_fft_autocorrelation_listings": int(fb[0]),
                                "reference_order_parcor_lag_sums": int(fb[1]), "scalar_longterm_lag_sums": int(fb[2]),
                                "note": "block x channels of rank 0's file that left the fast path so that ties and "
                                        "ill-conditioned recursions resolve exactly as in the reference (DESIGN.md section 4)"},
    }
    state["enc_launches"] = enc_launches

    # ---- end to end from pageable memory (what a drop-in caller such as the reference CLI passes: malloc) ----
    if headline:
        p_pcm = np.empty((nch, n), dtype=np.int32)
        p_pcm[:] = h_pcm
        p_stream = np.zeros(cap, dtype=np.uint8)
        p_dec = np.zeros((nch, n), dtype=np.int32)
        p_in = capi._planar_pointers(p_pcm)
        p_out = capi._planar_pointers(p_dec)
        sz, got = C.c_uint32(0), C.c_uint32(0)

        def enc_pageable():
            rc = L.SLAEncoder_EncodeWhole(codec.enc, p_in, n, p_stream.ctypes.data, cap, C.byref(sz))
            if rc != 0:
                raise RuntimeError(f"pageable encode rc={rc}")

        def dec_pageable():
            rc = L.SLADecoder_DecodeWhole(codec.dec, p_stream.ctypes.data, sz.value, p_out, n, C.byref(got))
            if rc != 0:
                raise RuntimeError(f"pageable decode rc={rc}")
        enc_pageable()
        pg_ms = D.timed(enc_pageable, steps)
        dec_pageable()
        pg_dec_ms = D.timed(dec_pageable, steps)
        pg_ok = sz.value == stream_bytes and bool(np.array_equal(p_stream[:sz.value], h_stream_t.numpy()[:sz.value])) \
            and bool(np.array_equal(p_dec, h_pcm))
        rec["e2e_pageable"] = {"value": world * chsamp / (pg_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": pg_ms,
                               "decode_value": world * chsamp / (pg_dec_ms * 1e-3) / 1e6, "decode_ms_per_step": pg_dec_ms,
                               "ratio_to_page_locked": pg_ms / e2e_ms, "decode_ratio_to_page_locked": pg_dec_ms / e2e_dec_ms,
                               "stream_identical_and_roundtrip": D.all_true(pg_ok),
                               "host_memory": "pageable (numpy / malloc), staged through pinned pieces by the library"}
        del p_pcm, p_stream, p_dec

    # ---- the raw-PCM entry points (interleaved little-endian PCM in page-locked host memory) ----
    if bits in (8, 16, 24, 32):
        fb = nch * bits // 8
        h_raw_t = pinned(n * fb, torch.uint8)
        if bits == 16:
            h_raw_t.numpy().view("<i2").reshape(n, nch)[:] = (h_pcm >> 16).T
        else:
            h_raw_t.numpy()[:] = np.frombuffer(capi.planar_to_pcm(h_pcm, bits), dtype=np.uint8)
        h_back_t = pinned(n * fb, torch.uint8)
        codec.enc_pcm(h_raw_t, n, h_stream_t, cap)
        pcm_enc_ms = D.timed(lambda: codec.enc_pcm(h_raw_t, n, h_stream_t, cap), steps)
        pcm_size = codec.size.value
        pcm_same = pcm_size == stream_bytes and bool(torch.equal(h_stream_t[:pcm_size].to(dev), d_stream[:pcm_size]))
        codec.dec_pcm(h_stream_t, pcm_size, h_back_t, n)
        pcm_dec_ms = D.timed(lambda: codec.dec_pcm(h_stream_t, pcm_size, h_back_t, n), steps)
        pcm_exact = bool(torch.equal(h_back_t, h_raw_t)) and codec.got.value == n
        rec["e2e_pcm"] = {"value": world * chsamp / (pcm_enc_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": pcm_enc_ms,
                          "h2d_bytes_per_step": n * fb, "d2h_bytes_per_step": stream_bytes,
                          "decode_value": world * chsamp / (pcm_dec_ms * 1e-3) / 1e6, "decode_ms_per_step": pcm_dec_ms,
                          "stream_equals_device_stream": D.all_true(pcm_same), "roundtrip": D.all_true(pcm_exact),
                          "api": "SLAB200_Encoder_EncodePCM / SLAB200_Decoder_DecodePCM (interleaved little-endian PCM, host)"}
        del h_raw_t, h_back_t

    # ---- stitch metadata across ranks (the only collective of the weak-scaling run: sizes -> offsets) ----
    sizes = [stream_bytes]
    if world > 1:
        import torch.distributed as dist
        t = torch.zeros(world, dtype=torch.int64, device=dev)
        t[rank] = stream_bytes
        dist.all_reduce(t)
        sizes = [int(x) for x in t.tolist()]
    rec["stitch"] = {"sizes": sizes, "offsets": [43 + sum(sizes[:i]) - 43 * i for i in range(len(sizes))]}

    if headline:
        rec["kernels_ms"] = kern_ms
        rec["decode"]["kernels_ms"] = dec_kern_ms
        rec["single_pass_profiled_ms_per_step"] = single_pass_ms

    # ---- rank 0: sample tiles for the CPU baseline and byte identity (taken before the buffers go) ----
    if rank == 0 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        tile_s = 30 if name == "C2" else 10
        tiles, ranges = file_tiles(h_pcm, bits, rate, preset, maxblk, cores, tile_s)
        state.setdefault("cpu_jobs", []).append((name, rec, tiles, f"{cores} ranges x {tile_s} s of the {name} file "
                                                 "(block-aligned starts), one process per core", bg))
    state[name + "_single_stream_md5"] = None
    codec.close()
    del d_pcm, d_stream, d_dec, h_pcm_t, h_stream_t, h_dec_t
    torch.cuda.empty_cache()
    return rec


def gpu_encode_tiles(lib, tiles):
    """every tile as a file of its own through SLAEncoder_EncodeWhole of the product"""
    out = []
    for pcm, bits, rate, preset in tiles:
        rc, data = lib.encode_whole(pcm, bits, rate, capi.preset_parameter(preset, pcm.shape[0]))
        out.append(data if rc == 0 else None)
    return out


def finish_cpu_jobs(lib, state):
    """rank 0, after every GPU leg: reference decoders joined, CPU baselines timed with nothing else running"""
    for name, rec, tiles, what, bg in state.get("cpu_jobs", []):
        rec["reference_decoder_ok"] = join_ref_decode(bg)
    for name, rec, tiles, what, bg in state.get("cpu_jobs", []):
        t0 = time.perf_counter()
        r = cpu_reference(tiles, keep_streams=True, what=what)
        mine = gpu_encode_tiles(lib, tiles)
        same = sum(int(m is not None and m == w) for m, w in zip(mine, r["streams"]))
        diff = [i for i, (m, w) in enumerate(zip(mine, r["streams"])) if m != w]
        rec["cpu_baseline"] = {"value": r["encode_all_core"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                               "sample": r["sample"], "per_core": r["encode_per_core"],
                               "decode_all_core": r["decode_all_core"], "decode_per_core": r["decode_per_core"],
                               "roundtrip_ok": r["ok"], "wall_s": round(time.perf_counter() - t0, 1)}
        rec["bit_exact"]["byte_identical_to_reference"] = f"{same}/{len(tiles)} sample files"
        rec["bit_exact"]["mismatching_sample_files"] = [
            {"tile": i, "gpu_bytes": len(mine[i] or b""), "reference_bytes": len(r["streams"][i])} for i in diff]
        log(f"[{name}] CPU baseline {r['encode_all_core']:.1f} M/s on {r['cores']} cores, byte-identical {same}/{len(tiles)}")


# ============================================================================= C4: 8-channel files, sharded by file
def leg_c4(a, D, L, lib, state):
    import torch
    c = workloads.CONFIGS["C4"]
    nch, bits, rate, preset, nfiles = c["channels"], c["bits"], c["rate"], c["preset"], c["files_per_gpu"]
    seconds = min(a.seconds, c["seconds"]) if a.seconds else c["seconds"]
    n = seconds * rate
    fb = nch * bits // 8
    chsamp = n * nch * nfiles
    dev, rank, world = D.dev, D.rank, D.world
    steps = SUB_STEPS
    codec = Codec(L, nch, bits, rate, preset)
    t0 = time.perf_counter()
    base = workloads.c4_base(rank, seconds)
    base24 = np.ascontiguousarray(((base >> 8).T.reshape(n, nch, 1).view(np.uint8))[:, :, :3])      # [n, ch, 3] little-endian
    h_raw = pinned((nfiles, n * fb), torch.uint8)
    for f in range(nfiles):
        h_raw[f].numpy().reshape(n, nch, 3)[:] = np.roll(np.roll(base24, f, axis=1), f * 7919, axis=0)
    d_base = torch.from_numpy(base).to(dev)
    log(f"[C4] rank {rank}: {nfiles} files of {n * nch / 1e6:.0f} M channel-samples built in {time.perf_counter() - t0:.1f} s")
    cap = 43 + int(n * nch * 3 * 1.25) + (1 << 20)
    h_stream = pinned((nfiles, cap), torch.uint8)
    h_back = pinned((nfiles, n * fb), torch.uint8)
    d_stream = torch.zeros(cap, dtype=torch.uint8, device=dev)
    d_dec = torch.empty((nch, n), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    # ---- value: device-resident, one file after the other (library events summed) ----
    def device_pass(check):
        enc_ms = dec_ms = 0.0
        ok = True
        sizes = []
        for f in range(nfiles):
            d_file = torch.roll(torch.roll(d_base, f, 0), f * 7919, 1).contiguous()
            torch.cuda.synchronize()          # the library's streams are not ordered against torch's (sla_b200.h)
            enc_ms += codec.enc_device(Codec.ptrs(d_file), n, d_stream, cap)
            sizes.append(codec.size.value)
            dec_ms += codec.dec_device(d_stream, codec.size.value, Codec.ptrs(d_dec), n)
            if check:
                ok = ok and bool(torch.equal(d_dec, d_file)) and codec.got.value == n
            del d_file
        return enc_ms, dec_ms, ok, sizes
    _, _, exact, dsizes = device_pass(True)
    D.barrier()
    codec.launches = 0
    enc_ms = dec_ms = 0.0
    for _ in range(steps):
        e, d, _, _ = device_pass(False)
        enc_ms += e
        dec_ms += d
    D.barrier()
    enc_ms, dec_ms = D.max(enc_ms / steps), D.max(dec_ms / steps)
    launches = codec.launches

    # ---- e2e: the batch entry points, raw 24-bit PCM in page-locked host memory ----
    eitems = (capi.EncodeItem * nfiles)()
    for f in range(nfiles):
        eitems[f].pcm = h_raw[f].data_ptr(); eitems[f].num_samples = n
        eitems[f].data = h_stream[f].data_ptr(); eitems[f].data_size = cap

    def enc_batch():
        rc = L.SLAB200_Encoder_EncodeBatchPCM(codec.enc, eitems, nfiles)
        if rc != 0 or any(eitems[f].result != 0 for f in range(nfiles)):
            raise RuntimeError("C4 batch encode failed")
    enc_batch()
    e2e_ms = D.timed(enc_batch, steps)
    ditems = (capi.BatchItem * nfiles)()
    for f in range(nfiles):
        ditems[f].data = h_stream[f].data_ptr(); ditems[f].data_size = eitems[f].output_size
        ditems[f].pcm = h_back[f].data_ptr(); ditems[f].capacity_samples = n

    def dec_batch():
        rc = L.SLAB200_Decoder_DecodeBatchPCM(codec.dec, ditems, nfiles)
        if rc != 0 or any(ditems[f].result != 0 for f in range(nfiles)):
            raise RuntimeError("C4 batch decode failed")
    dec_batch()
    e2e_dec_ms = D.timed(dec_batch, steps)
    host_exact = bool(torch.equal(h_back, h_raw))
    same_sizes = [int(eitems[f].output_size) for f in range(nfiles)] == dsizes
    stream_bytes = sum(dsizes)

    rec = {
        "config": workloads.describe("C4", world, **({"seconds": seconds} if a.seconds else {})),
        "metric": METRIC, "unit": UNIT, "steps": steps, "scaling": "weak",
        "value": world * chsamp / (enc_ms * 1e-3) / 1e6, "ms_per_step": enc_ms,
        "value_note": "device-resident planes, the GPU's files encoded one after the other (library CUDA events summed)",
        "compression_ratio": stream_bytes / (chsamp * bits / 8),
        "decode": {"value": world * chsamp / (dec_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": dec_ms},
        "e2e": {"value": world * chsamp / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": nfiles * n * fb, "d2h_bytes_per_step": stream_bytes,
                "decode_value": world * chsamp / (e2e_dec_ms * 1e-3) / 1e6, "decode_ms_per_step": e2e_dec_ms,
                "api": "SLAB200_Encoder_EncodeBatchPCM / SLAB200_Decoder_DecodeBatchPCM (packed 24-bit PCM, page-locked host)"},
        "gpu_launches": launches,
        "bit_exact": {"gpu_roundtrip": D.all_true(exact), "host_api_roundtrip": D.all_true(host_exact),
                      "batch_stream_sizes_equal_device_stream_sizes": D.all_true(same_sizes)},
    }
    if rank == 0 and not a.no_cpu_baseline:
        bg = spawn_ref_decode(h_stream[1].numpy()[:eitems[1].output_size], "C4", rank, 1, seconds if a.seconds else None)
        cores = os.cpu_count() or 1
        f1 = workloads.c4_file(base, 1)
        tiles, _ = file_tiles(f1, bits, rate, preset, codec.ep.max_num_block_samples, cores, 10)
        state.setdefault("cpu_jobs", []).append(("C4", rec, tiles, f"{cores} ranges x 10 s of C4 file 1 (block-aligned "
                                                 "starts), one process per core", bg))
    codec.close()
    del h_raw, h_stream, h_back, d_stream, d_dec, d_base
    torch.cuda.empty_cache()
    return rec


# ============================================================================= C5: corpus decode
def leg_c5(a, D, L, lib, state):
    """10 000 short files, presets mixed, file k on GPU k mod N.  The corpus is encoded on the GPU (batch
    call per preset), proven byte-identical to the reference on a random 1 % sample, then decoded through
    SLAB200_Decoder_DecodeBatchPCM in waves of bounded host memory."""
    import torch
    c = dict(workloads.CONFIGS["C5"])
    if a.corpus_files:
        c["files"] = a.corpus_files
    nch, bits, rate = c["channels"], c["bits"], c["rate"]
    fb = nch * bits // 8
    dev, rank, world = D.dev, D.rank, D.world
    steps = SUB_STEPS
    idx = workloads.corpus_index(c)
    t0 = time.perf_counter()
    pool = workloads.corpus_pool(c)                               # list of int16 [frames, 2]
    pool_frames = pool[0].shape[0]
    h_pool = pinned((len(pool), pool_frames * nch), torch.int16)
    for j, p in enumerate(pool):
        h_pool[j].numpy()[:] = p.reshape(-1)
    mine = np.arange(rank, c["files"], world)
    frames = idx["frames"][mine]
    chsamp_rank = int(frames.sum()) * nch
    chsamp_all = int(idx["frames"].sum()) * nch
    log(f"[C5] rank {rank}: pool of {len(pool)} signals in {time.perf_counter() - t0:.1f} s; {len(mine)} files, "
        f"{chsamp_rank / 1e6:.0f} M channel-samples")

    def src_ptr(k):
        return h_pool[int(idx["pool"][k])].data_ptr() + int(idx["first"][k]) * fb

    def src_view(k):
        f0 = int(idx["first"][k])
        return h_pool[int(idx["pool"][k])].numpy().reshape(pool_frames, nch)[f0:f0 + int(idx["frames"][k])]

    codecs = {p: Codec(L, nch, bits, rate, int(p)) for p in c["presets"]}
    # waves of at most ~1.6 G channel-samples: bounded page-locked memory whatever the corpus size
    wave_limit = 800_000_000
    waves, cur, acc = [], [], 0
    for k in mine:
        fr = int(idx["frames"][k])
        if cur and acc + fr * nch > wave_limit:
            waves.append(cur); cur, acc = [], 0
        cur.append(int(k)); acc += fr * nch
    if cur:
        waves.append(cur)
    max_wave_frames = max(sum(int(idx["frames"][k]) for k in w) for w in waves)
    max_wave_files = max(len(w) for w in waves)
    h_out = pinned(max_wave_frames * fb, torch.uint8)
    stream_cap = int(max_wave_frames * fb * 1.0) + 65536 * max_wave_files
    h_wave_streams = pinned(stream_cap, torch.uint8)

    # ---- corpus preparation: GPU batch encode, one call per preset and wave (timed, reported, not the metric) ----
    streams = {}
    enc_call_s = 0.0
    enc_calls = []
    t_enc0 = time.perf_counter()
    for w in waves:
        for p, codec in codecs.items():
            ks = [k for k in w if int(idx["preset"][k]) == p]
            if not ks:
                continue
            items = (capi.EncodeItem * len(ks))()
            at = 0
            for i, k in enumerate(ks):
                capk = 43 + int(idx["frames"][k]) * fb + 65536
                items[i].pcm = src_ptr(k); items[i].num_samples = int(idx["frames"][k])
                items[i].data = h_wave_streams.data_ptr() + at; items[i].data_size = capk
                at += capk
            assert at <= stream_cap
            torch.cuda.synchronize()
            t_call = time.perf_counter()
            rc = L.SLAB200_Encoder_EncodeBatchPCM(codec.enc, items, len(ks))
            enc_call_s += time.perf_counter() - t_call
            enc_calls.append((time.perf_counter() - t_call, sum(int(idx["frames"][k]) for k in ks) * nch))
            if rc != 0 or any(items[i].result != 0 for i in range(len(ks))):
                raise RuntimeError("C5 corpus encode failed")
            at = 0
            for i, k in enumerate(ks):
                streams[k] = h_wave_streams.numpy()[at:at + items[i].output_size].copy()
                at += 43 + int(idx["frames"][k]) * fb + 65536
    enc_s = D.max(time.perf_counter() - t_enc0)
    enc_call_s = D.max(enc_call_s)
    log(f"[C5] rank {D.rank}: corpus encode {enc_call_s:.2f} s inside SLAB200_Encoder_EncodeBatchPCM, {enc_s:.2f} s with the host-side staging")
    warm = sorted(cs / t / 1e6 for t, cs in enc_calls)
    warm_rate = warm[len(warm) // 2] if warm else 0.0
    log(f"[C5] rank {D.rank}: {len(enc_calls)} calls; M channel-samples/s of the first six: "
        + ", ".join(f"{cs / t / 1e6:.0f}" for t, cs in enc_calls[:6]) + f"; median call {warm_rate:.0f}")
    total_stream_bytes = sum(len(s) for s in streams.values())

    # ---- decode: all waves per step; streams staged into page-locked memory before the timed call ----
    dcodec = codecs[c["presets"][0]]
    kernel_ms = C.c_float(0)
    nl = C.c_uint32(0)

    def decode_pass(check):
        wall = dev_ms = 0.0
        ok = True
        launches = 0
        for w in waves:
            items = (capi.BatchItem * len(w))()
            at = out_at = 0
            for i, k in enumerate(w):
                s = streams[k]
                h_wave_streams.numpy()[at:at + len(s)] = s
                items[i].data = h_wave_streams.data_ptr() + at; items[i].data_size = len(s)
                items[i].pcm = h_out.data_ptr() + out_at; items[i].capacity_samples = int(idx["frames"][k])
                at += (len(s) + 63) & ~63
                out_at += int(idx["frames"][k]) * fb
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = L.SLAB200_Decoder_DecodeBatchPCM(dcodec.dec, items, len(w))
            wall += time.perf_counter() - t0
            L.SLAB200_Decoder_LastBatchTiming(dcodec.dec, C.byref(kernel_ms), C.byref(nl))
            dev_ms += kernel_ms.value
            launches += nl.value
            if rc != 0 or any(items[i].result != 0 for i in range(len(w))):
                raise RuntimeError("C5 batch decode failed")
            if check:
                out_at = 0
                o16 = h_out.numpy().view("<i2")
                for k in w:
                    fr = int(idx["frames"][k])
                    ok = ok and bool(np.array_equal(o16[out_at // 2:out_at // 2 + fr * nch].reshape(fr, nch), src_view(k)))
                    out_at += fr * fb
        return wall, dev_ms, ok, launches
    _, _, exact, _ = decode_pass(True)
    D.barrier()
    wall = dev_ms = 0.0
    launches = 0
    for _ in range(steps):
        w_, d_, _, l_ = decode_pass(False)
        wall += w_; dev_ms += d_; launches += l_
    D.barrier()
    e2e_ms = D.max(1e3 * wall / steps)
    dev_ms = D.max(dev_ms / steps)

    rec = {
        "config": workloads.describe("C5", world, **({"files": c["files"]} if a.corpus_files else {})),
        "metric": "decode_throughput", "unit": UNIT, "steps": steps, "scaling": "strong",
        "value": chsamp_all / (dev_ms * 1e-3) / 1e6, "ms_per_step": dev_ms,
        "value_note": "channel-samples / kernel time summed over the batch groups of the slowest GPU (groups overlap on the "
                      "device, so this understates the device rate)",
        "e2e": {"value": chsamp_all / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": total_stream_bytes, "d2h_bytes_per_step": chsamp_rank * bits // 8,
                "bytes_are": "per GPU", "api": "SLAB200_Decoder_DecodeBatchPCM (streams and 16-bit PCM in page-locked host memory), "
                "waves of <= 800 M channel-samples"},
        "compression_ratio": D.sum(total_stream_bytes) / (chsamp_all * bits / 8),
        "corpus_encode": {"value": chsamp_all / enc_call_s / 1e6, "unit": UNIT, "seconds": enc_call_s,
                          "seconds_with_host_staging": enc_s,
                          "median_call_value_rank0": warm_rate,
                          "value_note": "wall clock inside the calls (PCM and streams in page-locked host memory, first calls grow the arenas); "
                                        "the staging figure adds the bench's own copies of every stream out of the wave buffer",
                          "api": "SLAB200_Encoder_EncodeBatchPCM (files merged into groups of <= 48 M frames per launch sequence), one call per preset and wave"},
        "gpu_launches": launches,
        "bit_exact": {"gpu_decode_equals_source": D.all_true(exact)},
    }

    # ---- rank 0: 1 % random sample: reference encodes (bytes compared) and decodes (the CPU baseline) ----
    if rank == 0 and not a.no_cpu_baseline:
        rng = np.random.default_rng(20261018)
        sample = sorted(int(k) for k in rng.choice(c["files"], max(1, c["files"] // 100), replace=False))
        tiles = [(workloads.planar_of(np.array(src_view(k))), bits, rate, int(idx["preset"][k])) for k in sample]
        state.setdefault("c5_job", (rec, tiles, sample))
    for codec in codecs.values():
        codec.close()
    del h_pool, h_out, h_wave_streams
    torch.cuda.empty_cache()
    return rec


def finish_c5(lib, state):
    if "c5_job" not in state:
        return
    rec, tiles, sample = state["c5_job"]
    t0 = time.perf_counter()
    r = cpu_reference(tiles, keep_streams=True,
                      what=f"a random 1 % of the corpus ({len(tiles)} files, seed 20261018), one process per core")
    mine = gpu_encode_tiles(lib, tiles)
    same = sum(int(m is not None and m == w) for m, w in zip(mine, r["streams"]))
    # the reference decoder on the GPU-encoded streams of the sample
    ref = capi.SLALibrary(REF_SO)
    dec_ok = 0
    for (pcm, _, _, _), m in zip(tiles, mine):
        rc, out, _ = ref.decode_whole(m) if m is not None else (1, None, None)
        dec_ok += int(rc == 0 and np.array_equal(out, pcm))
    rec["cpu_baseline"] = {"value": r["decode_all_core"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                           "sample": r["sample"], "per_core": r["decode_per_core"],
                           "encode_all_core": r["encode_all_core"], "roundtrip_ok": r["ok"],
                           "wall_s": round(time.perf_counter() - t0, 1)}
    rec["bit_exact"]["byte_identical_to_reference"] = f"{same}/{len(tiles)} sample files"
    rec["reference_decoder_ok"] = {"ok": dec_ok == len(tiles), "files": f"{dec_ok}/{len(tiles)}",
                                   "decoder": "reference SLADecoder_DecodeWhole on the GPU-encoded streams of the sample"}
    log(f"[C5] CPU decode baseline {r['decode_all_core']:.1f} M/s, byte-identical {same}/{len(tiles)}")


# ============================================================================= strong scaling: one file over all GPUs
def leg_strong(a, D, L, lib, state):
    from sla_b200 import shard
    return shard.bench_strong(a, D, L, lib, state, SUB_STEPS, log)


# ============================================================================= GPU arm
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
    D = Dist(rank, world, dev)
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    sampler.start()
    state = {"sampler": sampler}
    t_start = time.perf_counter()

    head = leg_long_file("C2", a, D, L, a.steps, a.warmup, True, state)
    clocks = sampler.stop()
    log(f"[C2] done at {time.perf_counter() - t_start:.0f} s")

    want = [x for x in a.configs.split(",") if x]
    subs = {}
    for name in want:
        t0 = time.perf_counter()
        try:
            if name == "C3":
                subs[name] = leg_long_file("C3", a, D, L, SUB_STEPS, 1, False, state)
            elif name == "C4":
                subs[name] = leg_c4(a, D, L, lib, state)
            elif name == "C5":
                subs[name] = leg_c5(a, D, L, lib, state)
            elif name == "strong":
                if world > 1:
                    subs[name] = leg_strong(a, D, L, lib, state)
            else:
                subs[name] = {"error": "unknown config"}
        except Exception as e:      # noqa: BLE001
            if world > 1:
                raise               # ranks must stay in step: a failed leg ends the run
            subs[name] = {"error": repr(e)}
        log(f"[{name}] leg took {time.perf_counter() - t0:.1f} s on rank {rank}")

    if world > 1:
        D.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- rank 0 alone: reference decoders joined, CPU baselines, the line ----
    if not a.no_cpu_baseline:
        finish_cpu_jobs(lib, state)
        finish_c5(lib, state)
    chsamp = head["config"]["channel_samples_per_gpu"]
    c = head["compressed_bytes_per_channel_sample"]
    b_enc = 4.0 + c                                 # int32-planar C-API path
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kern_ms = head.pop("kernels_ms")
    dec_kern_ms = head["decode"].get("kernels_ms", {})
    top_name, top_ms = max(kern_ms.items(), key=lambda kv: kv[1]) if kern_ms else ("?", float("nan"))
    kernels_total = sum(kern_ms.values())
    achieved = b_enc * chsamp / (top_ms * 1e-3) / 1e9
    traffic = {}
    for tname in ("traffic_r02.json", "traffic_r01.json"):
        try:
            with open(os.path.join(ROOT, "profiles", tname)) as f:
                tj = json.load(f)
            if tj.get("channel_samples") == chsamp:
                traffic = tj.get("dram_bytes_per_launch", {})
                break
        except Exception:
            pass
    # what bounds each kernel: only some of them stream
    issue_notes = {}
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_bounds.json")) as f:
            issue_notes = json.load(f)
    except Exception:
        pass
    short = top_name.split()[-1]
    bound = issue_notes.get(short, {}).get("bound", "hbm")
    stream_defs = {"E0 k_enc_scan": 4.0, "E9 k_enc_pack": 4.0 + 2.0 + c, "E10 k_enc_crc": c,
                   "D1a k_dec_crc": c, "D3 k_dec_output": 8.0, "D0 k_dec_findsync": c}
    streaming = {}
    for sname, bpcs in stream_defs.items():
        ms = kern_ms.get(sname) or dec_kern_ms.get(sname)
        if ms:
            gbs = bpcs * chsamp / (ms * 1e-3) / 1e9
            streaming[sname] = {"ms": ms, "algorithmic_bytes_per_channel_sample": bpcs, "achieved_gbs": gbs,
                                "frac_of_measured_hbm": gbs / peak, "traffic": traffic.get(sname.split()[-1]),
                                "note": issue_notes.get(sname.split()[-1], {}).get("note")}
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32/f64", "data": "synthetic",
        "config": head["config"],
        "timing": {"input_flush": "inputs (%.2f GB per GPU) larger than L2" % (chsamp * 4 / 1e9),
                   "timer": "CUDA events on the library stream around the whole call (H2D-less device path), max over ranks",
                   "wall_ms_per_step": head["wall_ms_per_step"],
                   "single_pass_profiled_ms_per_step": head.get("single_pass_profiled_ms_per_step"),
                   "per_kernel_table": "rank 0 alone (the other ranks wait at a barrier)"},
        "pcm_mb_per_s": head["pcm_mb_per_s"],
        "compressed_bytes_per_channel_sample": c,
        "compression_ratio": head["compression_ratio"],
        "decode": head["decode"],
        "e2e": head["e2e"],
        "e2e_pageable": head.get("e2e_pageable"),
        "e2e_pcm": head.get("e2e_pcm"),
        "roofline": {"bound": bound, "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic.get(short),
                     "algorithmic_bytes_per_channel_sample": b_enc, "kernel_ms": top_ms,
                     "kernel_share_of_step": top_ms / kernels_total if kernels_total else None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "whole_step_frac": b_enc * chsamp / (head["ms_per_step"] * 1e-3) / 1e9 / peak,
                     "limiter": issue_notes.get(short, {}).get("note")},
        "streaming_kernels": streaming,
        "kernels_ms": kern_ms,
        "gpu_launches": state.get("enc_launches", 0),
        "bit_exact": head["bit_exact"],
        "exactness_fallbacks": head.get("exactness_fallbacks"),
        "reference_decoder_ok": head.get("reference_decoder_ok"),
        "stitch": head["stitch"],
        "clocks": clocks,
        "configs": subs,
        "bench_wall_s": round(time.perf_counter() - t_start, 1),
    }
    if "cpu_baseline" in head:
        line["cpu_baseline"] = head["cpu_baseline"]
    emit(line)


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
