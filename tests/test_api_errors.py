"""Status codes and small-API behaviour of the drop-in boundary, pinned against the reference itself.

Every expectation below is first asserted on the unmodified reference (oracle/_ref/libsla_ref.so), then
on the host-simulator build and, under -m gpu, on libsla_b200.so.  Cases follow the reference's own
tests: test/test_SLAEncoder.c:24-296, test/test_SLADecoder.c:28-147,244-535.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import signal_set
from sla_b200 import capi


def make_encoder(lib, **over):
    cap = dict(capi.CLI_CAPACITY); cap.update(over)
    cfg = capi.EncoderConfig(**cap, verpose_flag=0)
    h = lib.lib.SLAEncoder_Create(C.byref(cfg))
    assert h
    return h


def check_encoder_setters(lib):
    L = lib.lib
    assert L.SLAEncoder_Create(None) is None
    L.SLAEncoder_Destroy(None)                                   # NULL-safe, SLAEncoder.c:135
    enc = make_encoder(lib, max_num_channels=2, max_num_block_samples=8192, max_parcor_order=16,
                       max_longterm_order=3, max_lms_order_per_filter=8)
    try:
        wf = capi.WaveFormat(2, 16, 44100, 0)
        ep = capi.EncodeParameter(8, 1, 4, capi.CH_STEREO_MS, capi.WIN_SIN, 4096)
        assert L.SLAEncoder_SetWaveFormat(None, C.byref(wf)) == capi.INVALID_ARGUMENT
        assert L.SLAEncoder_SetWaveFormat(enc, None) == capi.INVALID_ARGUMENT
        assert L.SLAEncoder_SetEncodeParameter(None, C.byref(ep)) == capi.INVALID_ARGUMENT
        assert L.SLAEncoder_SetEncodeParameter(enc, None) == capi.INVALID_ARGUMENT
        assert L.SLAEncoder_SetWaveFormat(enc, C.byref(capi.WaveFormat(3, 16, 44100, 0))) == capi.EXCEED_HANDLE_CAPACITY
        assert L.SLAEncoder_SetWaveFormat(enc, C.byref(capi.WaveFormat(2, 33, 44100, 0))) == capi.EXCEED_HANDLE_CAPACITY
        assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == capi.OK
        for bad in (capi.EncodeParameter(17, 1, 4, 0, 1, 4096), capi.EncodeParameter(8, 5, 4, 0, 1, 4096),
                    capi.EncodeParameter(8, 1, 16, 0, 1, 4096), capi.EncodeParameter(8, 1, 4, 0, 1, 8193),
                    capi.EncodeParameter(8, 1, 4, 0, 1, 2047)):
            assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(bad)) == capi.EXCEED_HANDLE_CAPACITY
        assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == capi.OK
    finally:
        L.SLAEncoder_Destroy(enc)


def check_header_roundtrip(lib):
    L = lib.lib
    h = capi.HeaderInfo()
    h.wave_format = capi.WaveFormat(2, 24, 96000, 3)
    h.encode_param = capi.EncodeParameter(32, 3, 8, capi.CH_STEREO_MS, capi.WIN_SIN, 16384)
    h.num_samples, h.num_blocks, h.max_block_size, h.max_bit_per_second = 123456789, 7536, 98765, 4321000
    buf = np.zeros(64, dtype=np.uint8)
    assert L.SLAEncoder_EncodeHeader(None, buf.ctypes.data, 64) == capi.INVALID_ARGUMENT
    assert L.SLAEncoder_EncodeHeader(C.byref(h), None, 64) == capi.INVALID_ARGUMENT
    assert L.SLAEncoder_EncodeHeader(C.byref(h), buf.ctypes.data, 42) == capi.INSUFFICIENT_BUFFER_SIZE
    assert L.SLAEncoder_EncodeHeader(C.byref(h), buf.ctypes.data, 43) == capi.OK
    raw = buf[:43].tobytes()
    assert raw[:4] == b"SL*\x01" and raw[4:8] == (35).to_bytes(4, "big") and raw[10:14] == (1).to_bytes(4, "big")
    rc, g = lib.decode_header(raw)
    assert rc == capi.OK
    assert (g.wave_format.num_channels, g.wave_format.bit_per_sample, g.wave_format.sampling_rate,
            g.wave_format.offset_lshift) == (2, 24, 96000, 3)
    assert (g.encode_param.parcor_order, g.encode_param.longterm_order, g.encode_param.lms_order_per_filter,
            g.encode_param.ch_process_method, g.encode_param.max_num_block_samples) == (32, 3, 8, 1, 16384)
    assert (g.num_samples, g.num_blocks, g.max_block_size, g.max_bit_per_second) == (123456789, 7536, 98765, 4321000)
    out = capi.HeaderInfo()
    assert L.SLADecoder_DecodeHeader(None, 43, C.byref(out)) == capi.INVALID_ARGUMENT
    assert L.SLADecoder_DecodeHeader(buf.ctypes.data, 43, None) == capi.INVALID_ARGUMENT
    assert L.SLADecoder_DecodeHeader(buf.ctypes.data, 42, C.byref(out)) == capi.INSUFFICIENT_DATA_SIZE
    bad = bytearray(raw); bad[13] = 2                              # format version
    assert lib.decode_header(bytes(bad))[0] == capi.INVALID_HEADER_FORMAT
    return raw


def check_encode_whole_errors(lib):
    L = lib.lib
    pcm = np.zeros((1, 4096), dtype=np.int32)
    out = np.zeros(65536, dtype=np.uint8)
    size = C.c_uint32(0)
    enc = make_encoder(lib)
    try:
        assert L.SLAEncoder_SetWaveFormat(enc, C.byref(capi.WaveFormat(1, 16, 44100, 0))) == capi.OK
        ep = capi.EncodeParameter(8, 1, 4, capi.CH_NONE, capi.WIN_SIN, 4096)
        assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == capi.OK
        p = capi._planar_pointers(pcm)
        assert L.SLAEncoder_EncodeWhole(None, p, 4096, out.ctypes.data, 65536, C.byref(size)) == capi.INVALID_ARGUMENT
        assert L.SLAEncoder_EncodeWhole(enc, None, 4096, out.ctypes.data, 65536, C.byref(size)) == capi.INVALID_ARGUMENT
        assert L.SLAEncoder_EncodeWhole(enc, p, 4096, None, 65536, C.byref(size)) == capi.INVALID_ARGUMENT
        assert L.SLAEncoder_EncodeWhole(enc, p, 4096, out.ctypes.data, 65536, None) == capi.INVALID_ARGUMENT
        assert L.SLAEncoder_EncodeWhole(enc, p, 4096, out.ctypes.data, 42, C.byref(size)) == capi.INSUFFICIENT_BUFFER_SIZE
        # mid/side on a mono file, SLAEncoder.c:331-337
        ms = capi.EncodeParameter(8, 1, 4, capi.CH_STEREO_MS, capi.WIN_SIN, 4096)
        assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(ms)) == capi.OK
        assert L.SLAEncoder_EncodeWhole(enc, p, 4096, out.ctypes.data, 65536, C.byref(size)) == capi.INVALID_CHPROCESSMETHOD
        # unknown window, SLAEncoder.c:316-318
        bw = capi.EncodeParameter(8, 1, 4, capi.CH_NONE, 9, 4096)
        assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(bw)) == capi.OK
        assert L.SLAEncoder_EncodeWhole(enc, p, 4096, out.ctypes.data, 65536, C.byref(size)) == capi.INVALID_WINDOWFUNCTION_TYPE
    finally:
        L.SLAEncoder_Destroy(enc)


def encode_block(lib, pcm, bits, rate, ep, lshift=0, cap=None):
    L = lib.lib
    enc = make_encoder(lib)
    try:
        assert L.SLAEncoder_SetWaveFormat(enc, C.byref(capi.WaveFormat(pcm.shape[0], bits, rate, lshift))) == capi.OK
        assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == capi.OK
        cap = cap or (2 * pcm.size * 4 + 4096)
        out = np.zeros(cap, dtype=np.uint8)
        size = C.c_uint32(0)
        rc = L.SLAEncoder_EncodeBlock(enc, capi._planar_pointers(pcm), pcm.shape[1], out.ctypes.data, cap, C.byref(size))
        return rc, out[:size.value].tobytes()
    finally:
        L.SLAEncoder_Destroy(enc)


def check_encode_block(lib, reflib):
    L = lib.lib
    enc = make_encoder(lib)
    pcm = np.zeros((2, 4096), dtype=np.int32)
    out = np.zeros(65536, dtype=np.uint8)
    size = C.c_uint32(0)
    p = capi._planar_pointers(pcm)
    # parameters not set yet, SLAEncoder.c:477-480
    assert L.SLAEncoder_EncodeBlock(enc, p, 4096, out.ctypes.data, 65536, C.byref(size)) == capi.PARAMETER_NOT_SET
    assert L.SLAEncoder_SetWaveFormat(enc, C.byref(capi.WaveFormat(2, 16, 44100, 0))) == capi.OK
    assert L.SLAEncoder_EncodeBlock(enc, p, 4096, out.ctypes.data, 65536, C.byref(size)) == capi.PARAMETER_NOT_SET
    ep = capi.EncodeParameter(8, 1, 4, capi.CH_STEREO_MS, capi.WIN_SIN, 8192)
    assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == capi.OK
    assert L.SLAEncoder_EncodeBlock(enc, p, 16385, out.ctypes.data, 65536, C.byref(size)) == capi.EXCEED_HANDLE_CAPACITY
    assert L.SLAEncoder_EncodeBlock(enc, p, 4096, out.ctypes.data, 10, C.byref(size)) == capi.INSUFFICIENT_DATA_SIZE
    assert L.SLAEncoder_EncodeBlock(None, p, 4096, out.ctypes.data, 65536, C.byref(size)) == capi.INVALID_ARGUMENT
    # silent block layout (test/test_SLAEncoder.c:343-371): 11-byte block, type SILENT
    assert L.SLAEncoder_EncodeBlock(enc, p, 4096, out.ctypes.data, 65536, C.byref(size)) == capi.OK
    blk = out[:size.value].tobytes()
    assert size.value == 11 and blk[:2] == b"\xff\xff" and blk[2:6] == (5).to_bytes(4, "big")
    assert blk[8:10] == (4096).to_bytes(2, "big") and blk[10] >> 6 == 1
    L.SLAEncoder_Destroy(enc)
    # real blocks must equal the reference's SLAEncoder_EncodeBlock byte for byte
    for name, sig, bits, rate in signal_set()[:3]:
        blockpcm = np.ascontiguousarray(sig[:, 3000:3000 + 6000])
        for preset in (0, 2):
            ep = capi.preset_parameter(preset, sig.shape[0])
            rc_r, want = encode_block(reflib, blockpcm, bits, rate, ep)
            rc, got = encode_block(lib, blockpcm, bits, rate, ep)
            assert rc_r == capi.OK and rc == capi.OK
            assert got == want, (name, preset)


def test_reference_behaviour(reflib):
    check_encoder_setters(reflib)
    check_header_roundtrip(reflib)
    check_encode_whole_errors(reflib)
    check_encode_block(reflib, reflib)


def test_hostsim_matches_reference_behaviour(hostsim, reflib):
    check_encoder_setters(hostsim)
    assert check_header_roundtrip(hostsim) == check_header_roundtrip(reflib)
    check_encode_whole_errors(hostsim)
    check_encode_block(hostsim, reflib)


@pytest.mark.gpu
def test_gpu_matches_reference_behaviour(product, reflib):
    check_encoder_setters(product)
    assert check_header_roundtrip(product) == check_header_roundtrip(reflib)
    check_encode_whole_errors(product)
    check_encode_block(product, reflib)


def test_empty_file(hostsim, reflib):
    for lib in (reflib, hostsim):
        rc, data = lib.encode_whole(np.zeros((2, 0), dtype=np.int32), 16, 44100, capi.preset_parameter(2, 2))
        assert rc == capi.OK and len(data) == 43
        rc, h = lib.decode_header(data)
        assert rc == capi.OK and h.num_blocks == 0 and h.num_samples == 0
