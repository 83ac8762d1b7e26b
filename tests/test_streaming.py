"""SLAStreamingDecoder_* (reference src/SLADecoder.c:734-1123): the host layer
over the GPU block decoder must deliver the samples SLADecoder_DecodeWhole delivers when it is fed the
way the reference CLI feeds it (src/main.c:365-409), report the reference's estimates and result codes,
and keep the fragment queue's contract (8 packets by pointer, collected once consumed)."""
import ctypes as C

import numpy as np
import pytest

from sla_b200 import capi

GOLDEN = ["a_wav_m0", "a_wav_m2", "s16_special_m0", "s16_special_m2", "s24_impulsive_m4"]


def _check_stream(lib, data, whole):
    rc, pcm = capi.streaming_decode(lib, data)
    assert rc == capi.OK
    assert pcm.shape == whole.shape and np.array_equal(pcm, whole)


@pytest.mark.parametrize("name", GOLDEN)
def test_streaming_equals_whole_hostsim(name, golden_stream, hostsim, oracle):
    data = golden_stream(name)
    rc, whole, _, _ = oracle.decode_whole(data)
    assert rc == 0
    _check_stream(hostsim, data, whole)


def test_streaming_small_fragments_hostsim(golden_stream, hostsim, oracle):
    """Fragments much smaller than a block: Decode returns no samples (OK) until the block is complete."""
    data = golden_stream("s16_special_m2")
    rc, whole, _, _ = oracle.decode_whole(data)
    sizes = iter(np.random.default_rng(3).integers(1, 3000, 1 << 16))
    rc, pcm = capi.streaming_decode(hostsim, data, fragment=lambda est: int(next(sizes)))
    assert rc == capi.OK and np.array_equal(pcm, whole)


def test_streaming_estimates_match_reference(golden_stream, hostsim, reflib):
    """Initial estimates and the per-call quota are the reference's (SLADecoder.c:844-845, 862-884)."""
    data = golden_stream("s16_special_m2")
    _, h = hostsim.decode_header(data)
    for hz, bits in ((120.0, 24), (60.0, 16), (1000.0, 32)):
        a, b = capi.StreamingDecoder(hostsim, hz, bits), capi.StreamingDecoder(reflib, hz, bits)
        try:
            if bits < h.wave_format.bit_per_sample:
                continue
            assert a.set_format(h) == b.set_format(h) == capi.OK
            assert a.samples_per_decode() == b.samples_per_decode()
            assert a.min_data_size() == b.min_data_size()
            assert a.remain() == b.remain() == (capi.OK, 0)
            assert a.decodable_samples() == b.decodable_samples()
        finally:
            a.close(); b.close()


def test_streaming_follows_reference_progress(golden_stream, hostsim, reflib):
    """Fed the reference CLI's way, both decoders finish with identical samples and the byte-rate estimate
    of ours tracks the block being played, like the reference's."""
    data = golden_stream("s16_special_m2")
    ra, a = capi.streaming_decode(hostsim, data)
    rb, b = capi.streaming_decode(reflib, data)
    assert ra == rb == capi.OK and np.array_equal(a, b)


def test_streaming_api_errors(golden_stream, hostsim, reflib):
    data = golden_stream("a_wav_m2")
    _, h = hostsim.decode_header(data)
    buf = np.frombuffer(data, dtype=np.uint8)
    for lib in (hostsim, reflib):
        assert not capi.StreamingDecoder(lib, 0.0).handle                  # SLADecoder.c:762-764
        assert not capi.StreamingDecoder(lib, -5.0).handle
        L = lib.lib
        v = C.c_uint32(0)
        assert L.SLAStreamingDecoder_GetRemainDataSize(None, C.byref(v)) == capi.INVALID_ARGUMENT
        assert L.SLAStreamingDecoder_AppendDataFragment(None, buf.ctypes.data, 4) == capi.INVALID_ARGUMENT
        sd = capi.StreamingDecoder(lib, 120.0, 8)
        try:
            assert L.SLAStreamingDecoder_AppendDataFragment(sd.handle, None, 4) == capi.INVALID_ARGUMENT
            assert L.SLAStreamingDecoder_EstimateMinimumNessesaryDataSize(sd.handle, None) == capi.INVALID_ARGUMENT
            assert sd.collect()[0] == capi.NO_DATA_FRAGMENTS               # nothing queued
            wide = capi.WaveFormat(1, 16, 48000, 0)                        # wider than max_bit_per_sample = 8
            assert L.SLAStreamingDecoder_SetWaveFormat(sd.handle, C.byref(wide)) == capi.EXCEED_HANDLE_CAPACITY
            assert sd.set_format(h) == capi.OK
        finally:
            sd.close()
        lib.lib.SLAStreamingDecoder_Destroy(None)                          # NULL-safe


def test_streaming_queue_contract(golden_stream, hostsim, reflib):
    """Eight packets by pointer; a ninth is refused; consumed packets come back in order (SLAUtility.c:733-870)."""
    data = golden_stream("s16_special_m2")
    buf = np.frombuffer(data, dtype=np.uint8)
    _, h = hostsim.decode_header(data)
    for lib in (hostsim, reflib):
        sd = capi.StreamingDecoder(lib)
        try:
            assert sd.set_format(h) == capi.OK
            at = capi.HEADER_SIZE
            assert sd.append(buf, at, 0) == capi.OK                        # empty fragment: accepted, not queued
            assert sd.collect()[0] == capi.NO_DATA_FRAGMENTS
            for i in range(8):
                assert sd.append(buf, at + 100 * i, 100) == capi.OK
            assert sd.append(buf, at + 800, 100) == capi.EXCEED_HANDLE_CAPACITY
            assert sd.remain() == (capi.OK, 800)
            for i in range(8):
                rc, p, n = sd.collect()
                assert (rc, p, n) == (capi.OK, buf.ctypes.data + at + 100 * i, 100)
            assert sd.collect()[0] == capi.NO_DATA_FRAGMENTS
            assert sd.append(buf, at + 800, 100) == capi.OK
            assert sd.remain() == (capi.OK, 900)
        finally:
            sd.close()


def test_streaming_starved_hostsim(golden_stream, hostsim):
    """No data at all: the first Decode may return nothing, a second one without new data is an error;
    a stream that does not start with a sync code is reported as such."""
    data = golden_stream("a_wav_m2")
    _, h = hostsim.decode_header(data)
    out = np.zeros((1, 4096), dtype=np.int32)
    sd = capi.StreamingDecoder(hostsim)
    try:
        assert sd.set_format(h) == capi.OK
        assert sd.decode(out, 0, 4096) == (capi.INSUFFICIENT_DATA_SIZE, 0)
        buf = np.frombuffer(data, dtype=np.uint8)
        assert sd.append(buf, capi.HEADER_SIZE, 64) == capi.OK
        assert sd.decode(out, 0, 4096) == (capi.OK, 0)                     # block incomplete, caller is feeding
        assert sd.decode(out, 0, 4096) == (capi.INSUFFICIENT_DATA_SIZE, 0)
    finally:
        sd.close()
    sd = capi.StreamingDecoder(hostsim)
    try:
        assert sd.set_format(h) == capi.OK
        junk = np.zeros(4096, dtype=np.uint8)
        assert sd.append(junk, 0, 4096) == capi.OK
        assert sd.decode(out, 0, 4096)[0] == capi.FAILED_TO_FIND_SYNC_CODE
    finally:
        sd.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN + ["ch8_24bit_m2"])
def test_streaming_equals_whole_gpu(name, golden_stream, product, oracle):
    data = golden_stream(name)
    rc, whole, _, _ = oracle.decode_whole(data)
    assert rc == 0
    _check_stream(product, data, whole)


@pytest.mark.gpu
def test_streaming_small_fragments_gpu(golden_stream, product, oracle):
    data = golden_stream("s24_impulsive_m4")
    rc, whole, _, _ = oracle.decode_whole(data)
    sizes = iter(np.random.default_rng(4).integers(1, 5000, 1 << 16))
    rc, pcm = capi.streaming_decode(product, data, fragment=lambda est: int(next(sizes)))
    assert rc == capi.OK and np.array_equal(pcm, whole)
