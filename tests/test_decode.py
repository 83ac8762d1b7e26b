"""GPU decode of reference-encoded streams must be bit-exact (SLADecoder_DecodeWhole through the C ABI).

The same assertions run twice: on the B200 build (-m gpu) and, without a GPU, on the host-simulator
build of the same kernels (kernel-logic unit test; not a product path).
"""
import numpy as np
import pytest

from conftest import pcm_md5, signal_set
from sla_b200 import capi

GOLDEN_NAMES = ["a_wav_m0", "a_wav_m2", "a_wav_m4", "s16_special_m2", "s16_special_m0",
                "s24_impulsive_m4", "ch8_24bit_m2"]


def _decode_golden(lib, name, manifest, golden_stream):
    m = manifest[name]
    rc, pcm, h = lib.decode_whole(golden_stream(name))
    assert rc == capi.OK
    assert pcm.shape == (m["channels"], m["samples"])
    assert pcm_md5(pcm) == m["pcm_md5"]
    assert h.num_blocks == len(m["blocks"])


def _decode_errors(lib, golden_stream):
    good = golden_stream("a_wav_m2")
    # flipped payload byte -> CRC mismatch on that block (test/test_SLADecoder.c:492-522)
    bad = bytearray(good); bad[2000] ^= 0x40
    rc, _, _ = lib.decode_whole(bytes(bad))
    assert rc == capi.DETECT_DATA_CORRUPTION
    # ... with the CRC check off the damaged block decodes to garbage, consumes a different number of
    # bytes than its size field says, and the reference then loses the next sync code
    rc, _, _ = lib.decode_whole(bytes(bad), crc=False)
    assert rc == capi.FAILED_TO_FIND_SYNC_CODE
    # broken sync code on the first block
    bad = bytearray(good); bad[43] = 0
    rc, _, _ = lib.decode_whole(bytes(bad))
    assert rc == capi.FAILED_TO_FIND_SYNC_CODE
    # truncated stream
    rc, _, _ = lib.decode_whole(good[:len(good) // 2])
    assert rc == capi.INSUFFICIENT_DATA_SIZE
    # output buffer too small
    rc, _, _ = lib.decode_whole(good, out_samples=1000)
    assert rc == capi.INSUFFICIENT_BUFFER_SIZE
    # header problems
    bad = bytearray(good); bad[0] = ord("X")
    rc, _, _ = lib.decode_whole(bytes(bad))
    assert rc == capi.INVALID_HEADER_FORMAT
    bad = bytearray(good); bad[20] ^= 1
    rc, _, _ = lib.decode_whole(bytes(bad))
    assert rc == capi.DETECT_DATA_CORRUPTION


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_hostsim_decodes_golden(name, manifest, golden_stream, hostsim):
    _decode_golden(hostsim, name, manifest, golden_stream)


def test_hostsim_decode_errors(golden_stream, hostsim):
    _decode_errors(hostsim, golden_stream)


def test_reference_agrees_on_decode_errors(golden_stream, reflib):
    """the expectations above are the reference's own behaviour"""
    _decode_errors(reflib, golden_stream)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_gpu_decodes_golden(name, manifest, golden_stream, product):
    _decode_golden(product, name, manifest, golden_stream)


@pytest.mark.gpu
def test_gpu_decode_errors(golden_stream, product):
    _decode_errors(product, golden_stream)


@pytest.mark.gpu
@pytest.mark.parametrize("preset", [0, 1, 2, 3, 4])
def test_gpu_decodes_oracle_streams(preset, product, oracle):
    """streams produced by the oracle encoder (== reference bytes) for every preset and signal"""
    from oracle import binding as ob
    for name, pcm, bits, rate in signal_set():
        ep = capi.preset_parameter(preset, pcm.shape[0])
        rc, data, _, _ = oracle.encode_whole(pcm, ob.make_params(pcm.shape[0], bits, rate, ep))
        assert rc == 0
        rc, dec, _ = product.decode_whole(data)
        assert rc == capi.OK, name
        assert np.array_equal(dec, pcm), name


def _patch_first_pitch(data: bytes, oracle, new_pitch: int):
    """Rewrite the pitch field of channel 0 in the first block that uses the long-term stage, and fix the
    block CRC.  The result is a valid stream (any pitch >= taps/2 + 1 is legal) whose audio is garbage,
    which is fine: decoders only have to agree on it."""
    rc, _, h, blocks = oracle.decode_whole(data)
    assert rc == 0
    P = h.parcor_order
    buf = bytearray(data)
    for blk in blocks:
        if blk.block_type != 0 or blk.pitch[0] < 3:
            continue
        base = blk.byte_offset * 8 + 82 + 4 + min(P, 3) * 16 + max(P - 3, 0) * 8
        bit = lambda p: (buf[p >> 3] >> (7 - (p & 7))) & 1
        assert bit(base) == 1
        old = 0
        for k in range(10):
            old = (old << 1) | bit(base + 1 + k)
        assert old == blk.pitch[0]
        for k in range(10):
            p = base + 1 + k
            v = (new_pitch >> (9 - k)) & 1
            buf[p >> 3] = (buf[p >> 3] & ~(0x80 >> (p & 7))) | (v << (7 - (p & 7)))
        crc = oracle.crc16(bytes(buf[blk.byte_offset + 8:blk.byte_offset + blk.block_size]))
        buf[blk.byte_offset + 6] = crc >> 8
        buf[blk.byte_offset + 7] = crc & 0xFF
        return bytes(buf)
    raise AssertionError("no long-term block found")


def _decode_short_pitch(lib, golden_stream, oracle):
    """pitch lags shorter than the kernel's sample chunk take a different code path"""
    for name, pitches in (("s24_impulsive_m4", (2, 3, 9, 10, 11, 16, 17, 18, 24)),
                          ("s16_special_m2", (3, 5, 8, 14, 15, 16, 17, 255))):
        for pitch in pitches:
            data = _patch_first_pitch(golden_stream(name), oracle, pitch)
            rc_o, want, _, _ = oracle.decode_whole(data)
            rc, got, _ = lib.decode_whole(data)
            assert rc_o == 0 and rc == capi.OK
            assert np.array_equal(got, want), (name, pitch)


def test_hostsim_decode_short_pitch(hostsim, golden_stream, oracle):
    _decode_short_pitch(hostsim, golden_stream, oracle)


def test_reference_decode_short_pitch(reflib, golden_stream, oracle):
    _decode_short_pitch(reflib, golden_stream, oracle)


@pytest.mark.gpu
def test_gpu_decode_short_pitch(product, golden_stream, oracle):
    _decode_short_pitch(product, golden_stream, oracle)


# ---- device-resident decode: the block chain is found on the device (parallel D0) -------------
def _device_walk(lib, manifest, golden_stream, use_torch):
    for name in GOLDEN_NAMES:
        m = manifest[name]
        rc, pcm, _ = lib.decode_whole_device(golden_stream(name), use_torch=use_torch)
        assert rc == capi.OK, name
        assert pcm.shape == (m["channels"], m["samples"]) and pcm_md5(pcm) == m["pcm_md5"], name
    good = golden_stream("s16_special_m2")
    cases = {"flip": bytearray(good), "sync0": bytearray(good), "trunc": bytearray(good[:len(good) // 2])}
    cases["flip"][3000] ^= 0x40
    cases["sync0"][43] = 0
    # a stray sync pattern inside a payload, with a size field that lands on nothing
    stray = bytearray(good); stray[5000:5006] = b"\xff\xff\x00\x00\x01\x00"
    cases["stray"] = stray
    # same results as the host-walk path of SLADecoder_DecodeWhole, with and without the CRC check
    for key, data in cases.items():
        for crc in (True, False):
            rc_h, pcm_h, _ = lib.decode_whole(bytes(data), crc=crc)
            rc_d, pcm_d, _ = lib.decode_whole_device(bytes(data), crc=crc, use_torch=use_torch)
            assert rc_d == rc_h, (key, crc, rc_d, rc_h)
            if rc_h == capi.OK:
                assert np.array_equal(pcm_d, pcm_h), (key, crc)
    rc, _, _ = lib.decode_whole_device(good, out_samples=1000, use_torch=use_torch)
    assert rc == capi.INSUFFICIENT_BUFFER_SIZE


def test_hostsim_device_walk(hostsim, manifest, golden_stream):
    _device_walk(hostsim, manifest, golden_stream, use_torch=False)


@pytest.mark.gpu
def test_gpu_device_walk(product, manifest, golden_stream):
    _device_walk(product, manifest, golden_stream, use_torch=True)


# ---- pipelined whole-file decode (block ranges on several contexts) -----------------------------
def _pipelined_decode(lib, manifest, golden_stream, monkeypatch):
    for chunks in ("2", "5"):
        monkeypatch.setenv("SLAB200_PIPE_DEC_CHUNKS", chunks)
        for name in GOLDEN_NAMES:
            _decode_golden(lib, name, manifest, golden_stream)
        _decode_errors(lib, golden_stream)
    monkeypatch.delenv("SLAB200_PIPE_DEC_CHUNKS", raising=False)


def test_hostsim_pipelined_decode(hostsim, manifest, golden_stream, monkeypatch):
    _pipelined_decode(hostsim, manifest, golden_stream, monkeypatch)


@pytest.mark.gpu
def test_gpu_pipelined_decode(product, manifest, golden_stream, monkeypatch):
    _pipelined_decode(product, manifest, golden_stream, monkeypatch)



# ---- the container header's max_num_block_samples is not binding for a block header --------------
def _relabel_max_block(data: bytes, oracle, new_max: int) -> bytes:
    """same stream with header field max_num_block_samples (bytes 33-34) rewritten and the header CRC fixed"""
    buf = bytearray(data)
    buf[33] = new_max >> 8; buf[34] = new_max & 0xFF
    crc = oracle.crc16(bytes(buf[10:43]))
    buf[8] = crc >> 8; buf[9] = crc & 0xFF
    return bytes(buf)


def _decode_block_above_header_max(lib, golden_stream, oracle, device, use_torch):
    """A block header may announce more samples than the container header's max_num_block_samples: the
    reference decodes such a block as long as it fits the handle (SLADecoder.c:633 only checks the
    caller's buffer), so every walk here must decode all of it - not the first max_num_block_samples."""
    for name in ("s16_special_m2", "s24_impulsive_m4"):
        good = golden_stream(name)
        rc, want, _ = lib.decode_whole(good)
        assert rc == capi.OK
        lied = _relabel_max_block(good, oracle, 2048)
        rc, got, h = lib.decode_whole(lied)
        assert rc == capi.OK and h.encode_param.max_num_block_samples == 2048
        assert np.array_equal(got, want), name
        if device:
            rc, got, _ = lib.decode_whole_device(lied, use_torch=use_torch)
            assert rc == capi.OK and np.array_equal(got, want), name


def test_reference_decodes_block_above_header_max(reflib, golden_stream, oracle):
    _decode_block_above_header_max(reflib, golden_stream, oracle, device=False, use_torch=False)


def test_hostsim_decodes_block_above_header_max(hostsim, golden_stream, oracle):
    _decode_block_above_header_max(hostsim, golden_stream, oracle, device=True, use_torch=False)


@pytest.mark.gpu
def test_gpu_decodes_block_above_header_max(product, golden_stream, oracle):
    _decode_block_above_header_max(product, golden_stream, oracle, device=True, use_torch=True)


def _degenerate_size_field(lib, golden_stream, use_torch):
    """size field 0xFFFFFFFA wraps field + 6 to 0: host walk, parallel device walk and the serial device
    walk must refuse it the same way (the reference reads past its buffer on this input: not compared)"""
    bad = bytearray(golden_stream("a_wav_m2"))
    bad[45:49] = b"\xff\xff\xff\xfa"
    rc_h, _, _ = lib.decode_whole(bytes(bad))
    rc_d, _, _ = lib.decode_whole_device(bytes(bad), use_torch=use_torch)
    assert rc_h == capi.INSUFFICIENT_DATA_SIZE and rc_d == rc_h


def test_hostsim_degenerate_size_field(hostsim, golden_stream):
    _degenerate_size_field(hostsim, golden_stream, use_torch=False)


@pytest.mark.gpu
def test_gpu_degenerate_size_field(product, golden_stream):
    _degenerate_size_field(product, golden_stream, use_torch=True)
