"""ctypes view of the SLA public C API (reference: src/include/public/SLA.h:26-86,
SLAEncoder.h:14-53, SLADecoder.h:17-59).

The same structures bind *any* shared library exporting that API, so the parity tests drive
``libsla_b200.so`` (the product) and ``oracle/_ref/libsla_ref.so`` (the unmodified reference)
through one code path.  This module contains no codec logic.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

# SLAApiResult, SLA.h:26-43
OK, NG, INVALID_ARGUMENT, EXCEED_HANDLE_CAPACITY, INSUFFICIENT_BUFFER_SIZE, \
    INVALID_CHPROCESSMETHOD, FAILED_TO_CALCULATE_COEF, FAILED_TO_PREDICT, FAILED_TO_SYNTHESIZE, \
    INSUFFICIENT_DATA_SIZE, INVALID_HEADER_FORMAT, DETECT_DATA_CORRUPTION, \
    FAILED_TO_FIND_SYNC_CODE, INVALID_WINDOWFUNCTION_TYPE, NO_DATA_FRAGMENTS, \
    PARAMETER_NOT_SET = range(16)

CH_NONE, CH_STEREO_MS = 0, 1
WIN_RECT, WIN_SIN, WIN_HANN, WIN_BLACKMAN, WIN_VORBIS = range(5)
HEADER_SIZE = 43


class WaveFormat(C.Structure):
    _fields_ = [("num_channels", C.c_uint32), ("bit_per_sample", C.c_uint32),
                ("sampling_rate", C.c_uint32), ("offset_lshift", C.c_uint8)]


class EncodeParameter(C.Structure):
    _fields_ = [("parcor_order", C.c_uint32), ("longterm_order", C.c_uint32),
                ("lms_order_per_filter", C.c_uint32), ("ch_process_method", C.c_int),
                ("window_function_type", C.c_int), ("max_num_block_samples", C.c_uint32)]


class HeaderInfo(C.Structure):
    _fields_ = [("wave_format", WaveFormat), ("encode_param", EncodeParameter),
                ("num_samples", C.c_uint32), ("num_blocks", C.c_uint32),
                ("max_block_size", C.c_uint32), ("max_bit_per_second", C.c_uint32)]


class EncoderConfig(C.Structure):
    _fields_ = [("max_num_channels", C.c_uint32), ("max_num_block_samples", C.c_uint32),
                ("max_parcor_order", C.c_uint32), ("max_longterm_order", C.c_uint32),
                ("max_lms_order_per_filter", C.c_uint32), ("verpose_flag", C.c_uint8)]


class DecoderConfig(C.Structure):
    _fields_ = [("max_num_channels", C.c_uint32), ("max_num_block_samples", C.c_uint32),
                ("max_parcor_order", C.c_uint32), ("max_longterm_order", C.c_uint32),
                ("max_lms_order_per_filter", C.c_uint32), ("enable_crc_check", C.c_uint8),
                ("verpose_flag", C.c_uint8)]


# the reference CLI's presets (src/main.c:63-70) and handle capacity (src/main.c:94-98)
PRESETS = {
    0: dict(parcor_order=8, longterm_order=1, lms_order_per_filter=4, ms=False, window=WIN_RECT, max_block=4096),
    1: dict(parcor_order=8, longterm_order=1, lms_order_per_filter=8, ms=True, window=WIN_SIN, max_block=12288),
    2: dict(parcor_order=16, longterm_order=1, lms_order_per_filter=8, ms=True, window=WIN_SIN, max_block=12288),
    3: dict(parcor_order=32, longterm_order=3, lms_order_per_filter=8, ms=True, window=WIN_SIN, max_block=12288),
    4: dict(parcor_order=32, longterm_order=3, lms_order_per_filter=8, ms=True, window=WIN_SIN, max_block=16384),
}
CLI_CAPACITY = dict(max_num_channels=8, max_num_block_samples=16384, max_parcor_order=48,
                    max_longterm_order=5, max_lms_order_per_filter=40)


def preset_parameter(preset: int, num_channels: int) -> EncodeParameter:
    """What `sla -e -m <preset>` sets (MS only for stereo, src/main.c:121-133)."""
    p = PRESETS[preset]
    return EncodeParameter(p["parcor_order"], p["longterm_order"], p["lms_order_per_filter"],
                           CH_STEREO_MS if (p["ms"] and num_channels == 2) else CH_NONE,
                           p["window"], p["max_block"])


def _planar_pointers(arr: np.ndarray):
    assert arr.dtype == np.int32 and arr.ndim == 2 and arr.flags.c_contiguous
    ptrs = (C.POINTER(C.c_int32) * arr.shape[0])()
    for ch in range(arr.shape[0]):
        ptrs[ch] = arr[ch].ctypes.data_as(C.POINTER(C.c_int32))
    return ptrs


class SLALibrary:
    """One loaded shared library exporting the SLA C API."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.SLAEncoder_Create.restype = C.c_void_p
        L.SLAEncoder_Create.argtypes = [C.POINTER(EncoderConfig)]
        L.SLAEncoder_Destroy.argtypes = [C.c_void_p]
        L.SLAEncoder_Destroy.restype = None
        L.SLAEncoder_SetWaveFormat.argtypes = [C.c_void_p, C.POINTER(WaveFormat)]
        L.SLAEncoder_SetEncodeParameter.argtypes = [C.c_void_p, C.POINTER(EncodeParameter)]
        L.SLAEncoder_EncodeHeader.argtypes = [C.POINTER(HeaderInfo), C.c_void_p, C.c_uint32]
        for name in ("SLAEncoder_EncodeBlock", "SLAEncoder_EncodeWhole"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                         C.POINTER(C.c_uint32)]
        L.SLADecoder_DecodeHeader.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(HeaderInfo)]
        L.SLADecoder_Create.restype = C.c_void_p
        L.SLADecoder_Create.argtypes = [C.POINTER(DecoderConfig)]
        L.SLADecoder_Destroy.argtypes = [C.c_void_p]
        L.SLADecoder_Destroy.restype = None
        L.SLADecoder_SetWaveFormat.argtypes = [C.c_void_p, C.POINTER(WaveFormat)]
        L.SLADecoder_SetEncodeParameter.argtypes = [C.c_void_p, C.POINTER(EncodeParameter)]
        L.SLADecoder_DecodeWhole.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                             C.c_uint32, C.POINTER(C.c_uint32)]

    # ---- convenience round trips used by tests and bench -------------------------------------
    def encode_whole(self, pcm: np.ndarray, bits: int, rate: int, param: EncodeParameter,
                     capacity: dict | None = None, out_capacity: int | None = None):
        """pcm: int32 [channels, samples], left-justified. Returns (SLAApiResult, bytes)."""
        cfg = EncoderConfig(**(capacity or CLI_CAPACITY), verpose_flag=0)
        enc = self.lib.SLAEncoder_Create(C.byref(cfg))
        if not enc:
            raise RuntimeError("SLAEncoder_Create failed")
        try:
            wf = WaveFormat(pcm.shape[0], bits, rate, 0)
            rc = self.lib.SLAEncoder_SetWaveFormat(enc, C.byref(wf))
            if rc != OK:
                return rc, b""
            rc = self.lib.SLAEncoder_SetEncodeParameter(enc, C.byref(param))
            if rc != OK:
                return rc, b""
            cap = out_capacity if out_capacity is not None else \
                HEADER_SIZE + 2 * pcm.shape[0] * pcm.shape[1] * max(bits // 8, 1) + 65536
            out = np.zeros(cap, dtype=np.uint8)
            size = C.c_uint32(0)
            ptrs = _planar_pointers(pcm)
            rc = self.lib.SLAEncoder_EncodeWhole(enc, ptrs, pcm.shape[1], out.ctypes.data, cap,
                                                 C.byref(size))
            return rc, out[:size.value].tobytes()
        finally:
            self.lib.SLAEncoder_Destroy(enc)

    def decode_header(self, data: bytes):
        h = HeaderInfo()
        buf = np.frombuffer(data, dtype=np.uint8)
        rc = self.lib.SLADecoder_DecodeHeader(buf.ctypes.data, len(data), C.byref(h))
        return rc, h

    def decode_whole(self, data: bytes, capacity: dict | None = None, crc: bool = True,
                     out_samples: int | None = None):
        """Returns (SLAApiResult, int32 [channels, samples] left-justified, header)."""
        rc, h = self.decode_header(data)
        if rc not in (OK, DETECT_DATA_CORRUPTION):
            return rc, None, h
        cfg = DecoderConfig(**(capacity or CLI_CAPACITY), enable_crc_check=1 if crc else 0,
                            verpose_flag=0)
        dec = self.lib.SLADecoder_Create(C.byref(cfg))
        if not dec:
            raise RuntimeError("SLADecoder_Create failed")
        try:
            nch = h.wave_format.num_channels
            n = h.num_samples if out_samples is None else out_samples
            pcm = np.zeros((nch, max(n, 1)), dtype=np.int32)
            got = C.c_uint32(0)
            buf = np.frombuffer(data, dtype=np.uint8)
            ptrs = _planar_pointers(pcm)
            rc = self.lib.SLADecoder_DecodeWhole(dec, buf.ctypes.data, len(data), ptrs, n,
                                                 C.byref(got))
            return rc, pcm[:, :got.value], h
        finally:
            self.lib.SLADecoder_Destroy(dec)

    def decode_whole_device(self, data: bytes, capacity: dict | None = None, crc: bool = True,
                            out_samples: int | None = None, use_torch: bool = False):
        """SLAB200_Decoder_DecodeWholeDevice: stream and output planes in device memory, block chain
        found on the device.  With use_torch the buffers are CUDA tensors (the product library); without,
        plain numpy arrays (only meaningful for the host-simulator build, whose "device" is the host)."""
        rc, h = self.decode_header(data)
        if rc not in (OK, DETECT_DATA_CORRUPTION):
            return rc, None, h
        L = self.lib
        L.SLAB200_Decoder_DecodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                                        C.c_uint32, C.POINTER(C.c_uint32)]
        cfg = DecoderConfig(**(capacity or CLI_CAPACITY), enable_crc_check=1 if crc else 0, verpose_flag=0)
        dec = L.SLADecoder_Create(C.byref(cfg))
        if not dec:
            raise RuntimeError("SLADecoder_Create failed")
        try:
            nch = h.wave_format.num_channels
            n = h.num_samples if out_samples is None else out_samples
            got = C.c_uint32(0)
            if use_torch:
                import torch
                buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
                pcm = torch.zeros((nch, max(n, 1)), dtype=torch.int32, device="cuda")
                ptrs = (C.c_void_p * nch)(*[pcm[c].data_ptr() for c in range(nch)])
                rc = L.SLAB200_Decoder_DecodeWholeDevice(dec, buf.data_ptr(), len(data), ptrs, n, C.byref(got))
                return rc, pcm[:, :got.value].cpu().numpy(), h
            buf = np.frombuffer(data, dtype=np.uint8).copy()
            pcm = np.zeros((nch, max(n, 1)), dtype=np.int32)
            ptrs = _planar_pointers(pcm)
            rc = L.SLAB200_Decoder_DecodeWholeDevice(dec, buf.ctypes.data, len(data), ptrs, n, C.byref(got))
            return rc, pcm[:, :got.value], h
        finally:
            L.SLADecoder_Destroy(dec)


# ---- raw interleaved PCM (WAV data-chunk layout) ------------------------------------------------
def planar_to_pcm(planar: np.ndarray, bits: int) -> bytes:
    """int32 [channels, samples] left-justified -> interleaved little-endian PCM bytes
    (8-bit unsigned, 16/24/32-bit signed; reference src/wav.c:630-668)."""
    nch, n = planar.shape
    v = (planar >> (32 - bits)).T.reshape(-1)                           # frame-major, int32
    if bits == 8:
        return (v + 128).astype(np.uint8).tobytes()
    if bits == 16:
        return v.astype("<i2").tobytes()
    if bits == 32:
        return v.astype("<i4").tobytes()
    b = v.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
    return np.ascontiguousarray(b).tobytes()


def pcm_to_planar(pcm: bytes, bits: int, nch: int) -> np.ndarray:
    raw = np.frombuffer(pcm, dtype=np.uint8)
    if bits == 8:
        v = raw.astype(np.int64) - 128
    elif bits == 16:
        v = raw.view("<i2").astype(np.int64)
    elif bits == 32:
        v = raw.view("<i4").astype(np.int64)
    else:
        t = raw.reshape(-1, 3).astype(np.int64)
        v = t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
    return np.ascontiguousarray((v << (32 - bits)).astype(np.int32).reshape(-1, nch).T)


def encode_pcm(lib: "SLALibrary", pcm: bytes, nch: int, bits: int, rate: int, param: EncodeParameter,
               capacity: dict | None = None, out_capacity: int | None = None):
    """SLAB200_Encoder_EncodePCM. Returns (SLAApiResult, bytes)."""
    L = lib.lib
    L.SLAB200_Encoder_EncodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                            C.POINTER(C.c_uint32)]
    cfg = EncoderConfig(**(capacity or CLI_CAPACITY), verpose_flag=0)
    enc = L.SLAEncoder_Create(C.byref(cfg))
    if not enc:
        raise RuntimeError("SLAEncoder_Create failed")
    try:
        wf = WaveFormat(nch, bits, rate, 0)
        rc = L.SLAEncoder_SetWaveFormat(enc, C.byref(wf))
        if rc == OK:
            rc = L.SLAEncoder_SetEncodeParameter(enc, C.byref(param))
        if rc != OK:
            return rc, b""
        n = len(pcm) // (nch * bits // 8)
        cap = out_capacity if out_capacity is not None else HEADER_SIZE + 2 * len(pcm) + 65536
        out = np.zeros(cap, dtype=np.uint8)
        src = np.frombuffer(pcm, dtype=np.uint8)
        size = C.c_uint32(0)
        rc = L.SLAB200_Encoder_EncodePCM(enc, src.ctypes.data, n, out.ctypes.data, cap, C.byref(size))
        return rc, out[:size.value].tobytes()
    finally:
        L.SLAEncoder_Destroy(enc)


def decode_pcm(lib: "SLALibrary", data: bytes, capacity: dict | None = None, crc: bool = True):
    """SLAB200_Decoder_DecodePCM. Returns (SLAApiResult, pcm bytes, header)."""
    L = lib.lib
    L.SLAB200_Decoder_DecodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                            C.POINTER(C.c_uint32)]
    rc, h = lib.decode_header(data)
    if rc not in (OK, DETECT_DATA_CORRUPTION):
        return rc, b"", h
    cfg = DecoderConfig(**(capacity or CLI_CAPACITY), enable_crc_check=1 if crc else 0, verpose_flag=0)
    dec = L.SLADecoder_Create(C.byref(cfg))
    if not dec:
        raise RuntimeError("SLADecoder_Create failed")
    try:
        fb = h.wave_format.num_channels * (h.wave_format.bit_per_sample // 8)
        out = np.zeros(max(h.num_samples * fb, 1), dtype=np.uint8)
        buf = np.frombuffer(data, dtype=np.uint8)
        got = C.c_uint32(0)
        rc = L.SLAB200_Decoder_DecodePCM(dec, buf.ctypes.data, len(data), out.ctypes.data, h.num_samples, C.byref(got))
        return rc, out[:got.value * fb].tobytes(), h
    finally:
        L.SLADecoder_Destroy(dec)


class BatchItem(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_size", C.c_uint32), ("pcm", C.c_void_p),
                ("capacity_samples", C.c_uint32), ("output_num_samples", C.c_uint32), ("result", C.c_int)]


def decode_batch_pcm(lib: "SLALibrary", streams: list, capacity: dict | None = None, crc: bool = True,
                     capacities: list | None = None):
    """SLAB200_Decoder_DecodeBatchPCM over host buffers.  Returns (rc, [(result, pcm bytes)] per stream)."""
    L = lib.lib
    L.SLAB200_Decoder_DecodeBatchPCM.argtypes = [C.c_void_p, C.POINTER(BatchItem), C.c_uint32]
    cfg = DecoderConfig(**(capacity or CLI_CAPACITY), enable_crc_check=1 if crc else 0, verpose_flag=0)
    dec = L.SLADecoder_Create(C.byref(cfg))
    if not dec:
        raise RuntimeError("SLADecoder_Create failed")
    try:
        items = (BatchItem * len(streams))()
        keep, outs, fbs = [], [], []
        for i, data in enumerate(streams):
            rc, h = lib.decode_header(data)
            ok = rc in (OK, DETECT_DATA_CORRUPTION) and h.wave_format.bit_per_sample in (8, 16, 24, 32)
            fb = h.wave_format.num_channels * (h.wave_format.bit_per_sample // 8) if ok else 1
            n = h.num_samples if ok else 0
            if capacities is not None and capacities[i] is not None:
                n = capacities[i]
            buf = np.frombuffer(data, dtype=np.uint8).copy()
            out = np.zeros(max(n * fb, 1), dtype=np.uint8)
            keep.append(buf); outs.append(out); fbs.append(fb)
            items[i].data = buf.ctypes.data; items[i].data_size = len(data)
            items[i].pcm = out.ctypes.data; items[i].capacity_samples = n
        rc = L.SLAB200_Decoder_DecodeBatchPCM(dec, items, len(streams))
        return rc, [(items[i].result, outs[i][:items[i].output_num_samples * fbs[i]].tobytes()) for i in range(len(streams))]
    finally:
        L.SLADecoder_Destroy(dec)


def encode_whole_device(lib: "SLALibrary", pcm: np.ndarray, bits: int, rate: int, param: EncodeParameter,
                        use_torch: bool = False, capacity: dict | None = None):
    """SLAB200_Encoder_EncodeWholeDevice: planes and stream in device memory (CUDA tensors with use_torch,
    plain numpy arrays for the host-simulator build).  Returns (SLAApiResult, bytes)."""
    L = lib.lib
    L.SLAB200_Encoder_EncodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                                    C.POINTER(C.c_uint32)]
    cfg = EncoderConfig(**(capacity or CLI_CAPACITY), verpose_flag=0)
    enc = L.SLAEncoder_Create(C.byref(cfg))
    if not enc:
        raise RuntimeError("SLAEncoder_Create failed")
    try:
        nch, n = pcm.shape
        wf = WaveFormat(nch, bits, rate, 0)
        rc = L.SLAEncoder_SetWaveFormat(enc, C.byref(wf))
        if rc == OK:
            rc = L.SLAEncoder_SetEncodeParameter(enc, C.byref(param))
        if rc != OK:
            return rc, b""
        cap = HEADER_SIZE + 2 * pcm.size * max(bits // 8, 1) + 65536
        size = C.c_uint32(0)
        if use_torch:
            import torch
            d_in = torch.from_numpy(np.ascontiguousarray(pcm)).cuda()
            d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
            ptrs = (C.c_void_p * nch)(*[d_in[c].data_ptr() for c in range(nch)])
            rc = L.SLAB200_Encoder_EncodeWholeDevice(enc, ptrs, n, d_out.data_ptr(), cap, C.byref(size))
            return rc, d_out[:size.value].cpu().numpy().tobytes()
        src = np.ascontiguousarray(pcm)
        out = np.zeros(cap, dtype=np.uint8)
        rc = L.SLAB200_Encoder_EncodeWholeDevice(enc, _planar_pointers(src), n, out.ctypes.data, cap, C.byref(size))
        return rc, out[:size.value].tobytes()
    finally:
        L.SLAEncoder_Destroy(enc)



# ---- streaming decoder (SLADecoder.h:24-32, 62-101) ------------------------------------------------
class StreamingDecoderConfig(C.Structure):
    _fields_ = [("core_config", DecoderConfig), ("decode_interval_hz", C.c_float),
                ("max_bit_per_sample", C.c_uint32)]


class StreamingDecoder:
    """SLAStreamingDecoder_* of any library exporting the API (product or reference)."""

    def __init__(self, lib: "SLALibrary", interval_hz: float = 120.0, max_bits: int = 24,
                 capacity: dict | None = None, crc: bool = True):
        L = self.L = lib.lib
        L.SLAStreamingDecoder_Create.restype = C.c_void_p
        L.SLAStreamingDecoder_Create.argtypes = [C.POINTER(StreamingDecoderConfig)]
        L.SLAStreamingDecoder_Destroy.argtypes = [C.c_void_p]
        L.SLAStreamingDecoder_Destroy.restype = None
        L.SLAStreamingDecoder_SetWaveFormat.argtypes = [C.c_void_p, C.POINTER(WaveFormat)]
        L.SLAStreamingDecoder_SetEncodeParameter.argtypes = [C.c_void_p, C.POINTER(EncodeParameter)]
        for name in ("EstimateMinimumNessesaryDataSize", "EstimateDecodableNumSamples",
                     "GetOutputNumSamplesPerDecode", "GetRemainDataSize"):
            getattr(L, "SLAStreamingDecoder_" + name).argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.SLAStreamingDecoder_AppendDataFragment.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.SLAStreamingDecoder_CollectDataFragment.argtypes = [C.c_void_p, C.POINTER(C.c_void_p),
                                                              C.POINTER(C.c_uint32)]
        L.SLAStreamingDecoder_Decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        cfg = StreamingDecoderConfig(
            DecoderConfig(**(capacity or CLI_CAPACITY), enable_crc_check=1 if crc else 0, verpose_flag=0),
            interval_hz, max_bits)
        self.handle = L.SLAStreamingDecoder_Create(C.byref(cfg))

    def close(self):
        if self.handle:
            self.L.SLAStreamingDecoder_Destroy(self.handle)
            self.handle = None

    def set_format(self, header: HeaderInfo):
        rc = self.L.SLAStreamingDecoder_SetWaveFormat(self.handle, C.byref(header.wave_format))
        if rc != OK:
            return rc
        return self.L.SLAStreamingDecoder_SetEncodeParameter(self.handle, C.byref(header.encode_param))

    def _u32(self, name):
        v = C.c_uint32(0)
        rc = getattr(self.L, "SLAStreamingDecoder_" + name)(self.handle, C.byref(v))
        return rc, v.value

    def min_data_size(self):
        return self._u32("EstimateMinimumNessesaryDataSize")

    def decodable_samples(self):
        return self._u32("EstimateDecodableNumSamples")

    def samples_per_decode(self):
        return self._u32("GetOutputNumSamplesPerDecode")

    def remain(self):
        return self._u32("GetRemainDataSize")

    def append(self, buf: np.ndarray, start: int, size: int):
        """buf must stay alive until the fragment has been collected (the decoder keeps the pointer)."""
        return self.L.SLAStreamingDecoder_AppendDataFragment(self.handle, buf.ctypes.data + start, size)

    def collect(self):
        p, n = C.c_void_p(0), C.c_uint32(0)
        rc = self.L.SLAStreamingDecoder_CollectDataFragment(self.handle, C.byref(p), C.byref(n))
        return rc, (p.value or 0), n.value

    def decode(self, out: np.ndarray, start: int, capacity: int):
        ptrs = (C.POINTER(C.c_int32) * out.shape[0])()
        for ch in range(out.shape[0]):
            ptrs[ch] = C.cast(out[ch].ctypes.data + 4 * start, C.POINTER(C.c_int32))
        got = C.c_uint32(0)
        rc = self.L.SLAStreamingDecoder_Decode(self.handle, ptrs, capacity, C.byref(got))
        return rc, got.value


def streaming_decode(lib: "SLALibrary", data: bytes, interval_hz: float = 120.0, max_bits: int = 24,
                     fragment=None):
    """The loop of the reference CLI (src/main.c:365-409): first fragment = header.max_block_size, then the
    decoder's own estimate per Decode call; `fragment(estimate)` may override the fragment size.
    Returns (SLAApiResult, int32 [channels, samples])."""
    rc, h = lib.decode_header(data)
    if rc != OK:
        return rc, None
    buf = np.frombuffer(data, dtype=np.uint8)
    sd = StreamingDecoder(lib, interval_hz, max_bits)
    if not sd.handle:
        raise RuntimeError("SLAStreamingDecoder_Create failed")
    try:
        rc = sd.set_format(h)
        if rc != OK:
            return rc, None
        out = np.zeros((h.wave_format.num_channels, max(h.num_samples, 1)), dtype=np.int32)
        done, at = 0, HEADER_SIZE
        while done < h.num_samples:
            est = h.max_block_size if done == 0 and at == HEADER_SIZE else sd.min_data_size()[1]
            give = min(fragment(est) if fragment else est, len(data) - at)
            rc = sd.append(buf, at, give)
            if rc != OK:
                return rc, out[:, :done]
            at += give
            rc, got = sd.decode(out, done, h.num_samples - done)
            if rc != OK:
                return rc, out[:, :done]
            done += got
            while sd.collect()[0] == OK:
                pass
        return OK, out[:, :done]
    finally:
        sd.close()


class EncodeItem(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("num_samples", C.c_uint32), ("data", C.c_void_p),
                ("data_size", C.c_uint32), ("output_size", C.c_uint32), ("result", C.c_int)]


def encode_batch_pcm(lib: "SLALibrary", pcms: list, nch: int, bits: int, rate: int, param: EncodeParameter,
                     capacity: dict | None = None, out_capacities: list | None = None):
    """SLAB200_Encoder_EncodeBatchPCM over host buffers.  Returns (rc, [(result, stream bytes)] per file)."""
    L = lib.lib
    L.SLAB200_Encoder_EncodeBatchPCM.argtypes = [C.c_void_p, C.POINTER(EncodeItem), C.c_uint32]
    cfg = EncoderConfig(**(capacity or CLI_CAPACITY), verpose_flag=0)
    enc = L.SLAEncoder_Create(C.byref(cfg))
    if not enc:
        raise RuntimeError("SLAEncoder_Create failed")
    try:
        wf = WaveFormat(nch, bits, rate, 0)
        rc = L.SLAEncoder_SetWaveFormat(enc, C.byref(wf))
        if rc == OK:
            rc = L.SLAEncoder_SetEncodeParameter(enc, C.byref(param))
        if rc != OK:
            return rc, []
        items = (EncodeItem * len(pcms))()
        keep, outs = [], []
        fb = nch * bits // 8
        for i, pcm in enumerate(pcms):
            src = np.frombuffer(pcm, dtype=np.uint8) if len(pcm) else np.zeros(1, dtype=np.uint8)
            cap = HEADER_SIZE + 2 * len(pcm) + 65536
            if out_capacities is not None and out_capacities[i] is not None:
                cap = out_capacities[i]
            out = np.zeros(max(cap, 1), dtype=np.uint8)
            keep.append(src); outs.append(out)
            items[i].pcm = src.ctypes.data; items[i].num_samples = len(pcm) // fb
            items[i].data = out.ctypes.data; items[i].data_size = cap
        rc = L.SLAB200_Encoder_EncodeBatchPCM(enc, items, len(pcms))
        return rc, [(items[i].result, outs[i][:items[i].output_size].tobytes()) for i in range(len(pcms))]
    finally:
        L.SLAEncoder_Destroy(enc)
