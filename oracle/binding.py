"""ctypes bindings for the CPU checkers - TEST INFRASTRUCTURE ONLY.

* ``Oracle``      -> oracle/libsla_oracle.so   (our plain-C restatement, sla_oracle.c)
* ``RefWhitebox`` -> oracle/_ref/libsla_ref_wb.so (the unmodified reference + ref_whitebox.c probes)
* ``reference_library()`` -> oracle/_ref/libsla_ref.so bound through the public C API

Imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under sla_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_CH, MAX_ORD, MAX_TAPS = 8, 64, 8


class Block(C.Structure):
    """OraBlock / RefWBBlock."""
    _fields_ = [("sample_offset", C.c_uint32), ("num_samples", C.c_uint32), ("block_type", C.c_uint32),
                ("block_size", C.c_uint32), ("byte_offset", C.c_uint32),
                ("rshift", C.c_uint32 * MAX_CH), ("pitch", C.c_uint32 * MAX_CH),
                ("parcor_code", (C.c_int32 * (MAX_ORD + 1)) * MAX_CH),
                ("lt_q31", (C.c_int32 * MAX_TAPS) * MAX_CH),
                ("rice_init", C.c_uint64 * MAX_CH),
                ("parcor", (C.c_double * (MAX_ORD + 1)) * MAX_CH),
                ("lt", (C.c_double * MAX_TAPS) * MAX_CH)]


class Params(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "num_channels", "bits_per_sample", "sampling_rate", "parcor_order", "longterm_order",
        "lms_order", "ch_process", "window_type", "max_block_samples", "fft_size")]


class Header(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "num_channels", "num_samples", "sampling_rate", "bits_per_sample", "offset_lshift",
        "parcor_order", "longterm_order", "lms_order", "ch_process", "num_blocks",
        "max_block_samples", "max_block_size", "max_bit_per_second")]


def build(target: str = "all") -> None:
    """(Re)build the checkers; `ref` is skipped by the Makefile when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def _ptrs(arr: np.ndarray):
    assert arr.dtype == np.int32 and arr.ndim == 2 and arr.flags.c_contiguous
    p = (C.POINTER(C.c_int32) * arr.shape[0])()
    for ch in range(arr.shape[0]):
        p[ch] = arr[ch].ctypes.data_as(C.POINTER(C.c_int32))
    return p


def roundup_pow2(x: int) -> int:
    return 1 << (x - 1).bit_length()


def make_params(nch, bits, rate, enc_param, handle_max_block=16384) -> Params:
    """enc_param: sla_b200.capi.EncodeParameter."""
    return Params(nch, bits, rate, enc_param.parcor_order, enc_param.longterm_order,
                  enc_param.lms_order_per_filter, int(enc_param.ch_process_method),
                  int(enc_param.window_function_type), enc_param.max_num_block_samples,
                  roundup_pow2(2 * handle_max_block))


class Oracle:
    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "libsla_oracle.so")
        if not os.path.exists(path):
            build("oracle")
        self.lib = L = C.CDLL(path)
        L.ora_crc16.restype = C.c_uint16
        L.ora_crc16.argtypes = [C.c_void_p, C.c_uint64]
        L.ora_code_length.restype = C.c_double
        L.ora_encode_whole.argtypes = [C.POINTER(Params), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                       C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32,
                                       C.POINTER(C.c_uint32), C.c_void_p]
        L.ora_decode_whole.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32,
                                       C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.ora_decode_header.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(Header)]

    def crc16(self, data: bytes) -> int:
        buf = np.frombuffer(data, dtype=np.uint8)
        return int(self.lib.ora_crc16(buf.ctypes.data if len(data) else None, len(data)))

    def encode_whole(self, pcm: np.ndarray, params: Params, want_residual=False, max_blocks=None):
        nch, n = pcm.shape
        cap = 43 + 2 * nch * n * max(params.bits_per_sample // 8, 1) + 65536
        out = np.zeros(cap, dtype=np.uint8)
        max_blocks = max_blocks or (n // 2048 + 16)
        blocks = (Block * max_blocks)()
        size, nblk = C.c_uint32(0), C.c_uint32(0)
        res = np.zeros_like(pcm) if want_residual else None
        rc = self.lib.ora_encode_whole(C.byref(params), _ptrs(pcm), n, out.ctypes.data, cap, C.byref(size),
                                       blocks, max_blocks, C.byref(nblk), _ptrs(res) if want_residual else None)
        return rc, out[:size.value].tobytes(), list(blocks[:nblk.value]), res

    def decode_whole(self, data: bytes, crc=True):
        h = Header()
        buf = np.frombuffer(data, dtype=np.uint8)
        rc = self.lib.ora_decode_header(buf.ctypes.data, len(data), C.byref(h))
        if rc not in (0, 2):
            return rc, None, h, []
        pcm = np.zeros((h.num_channels, max(h.num_samples, 1)), dtype=np.int32)
        got, nblk = C.c_uint32(0), C.c_uint32(0)
        max_blocks = h.num_samples // 1024 + 16
        blocks = (Block * max_blocks)()
        rc = self.lib.ora_decode_whole(buf.ctypes.data, len(data), 1 if crc else 0, _ptrs(pcm), h.num_samples,
                                       C.byref(got), blocks, max_blocks, C.byref(nblk))
        return rc, pcm[:, :got.value], h, list(blocks[:nblk.value])


class RefWhitebox:
    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "_ref", "libsla_ref_wb.so")
        if not os.path.exists(path):
            build("ref")
        self.lib = C.CDLL(path)
        assert self.lib.RefWB_SizeofBlock() == C.sizeof(Block)

    def encode_whole(self, pcm: np.ndarray, bits, rate, enc_param, capacity, want_residual=False):
        from sla_b200.capi import EncoderConfig, WaveFormat
        nch, n = pcm.shape
        cfg = EncoderConfig(**capacity, verpose_flag=0)
        wf = WaveFormat(nch, bits, rate, 0)
        cap = 43 + 2 * nch * n * max(bits // 8, 1) + 65536
        out = np.zeros(cap, dtype=np.uint8)
        max_blocks = n // 2048 + 16
        blocks = (Block * max_blocks)()
        size, nblk = C.c_uint32(0), C.c_uint32(0)
        res = np.zeros_like(pcm) if want_residual else None
        self.lib.RefWB_EncodeWhole.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                               C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p,
                                               C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]
        rc = self.lib.RefWB_EncodeWhole(C.byref(cfg), C.byref(wf), C.byref(enc_param), _ptrs(pcm), n,
                                        out.ctypes.data, cap, C.byref(size), blocks, max_blocks,
                                        C.byref(nblk), _ptrs(res) if want_residual else None)
        return rc, out[:size.value].tobytes(), list(blocks[:nblk.value]), res


def reference_library():
    """The unmodified reference behind its public C API."""
    from sla_b200.capi import SLALibrary
    path = os.path.join(HERE, "_ref", "libsla_ref.so")
    if not os.path.exists(path):
        build("ref")
    return SLALibrary(path)
