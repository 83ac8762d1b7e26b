"""Pins the oracle: plain-C restatement vs the unmodified reference vs the committed golden vectors."""
import hashlib

import numpy as np
import pytest

from conftest import pcm_md5, signal_set
from oracle import binding as ob
from sla_b200 import capi

GOLDEN_NAMES = ["a_wav_m0", "a_wav_m2", "a_wav_m4", "s16_special_m2", "s16_special_m0",
                "s24_impulsive_m4", "ch8_24bit_m2"]
# md5 of the reference CLI's output for test/a.wav, presets 0/2/4
A_WAV_MD5 = {"a_wav_m0": "48c60a59f94f70303be8207d7ea9dc03", "a_wav_m2": "9739dfd1acd3eeaec7a3f4345ee8c4a4",
             "a_wav_m4": "9ad138cb6ad58ab8b074eae1132b3c28"}


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_decodes_golden(name, manifest, golden_stream, oracle):
    data = golden_stream(name)
    m = manifest[name]
    assert hashlib.md5(data).hexdigest() == m["md5"]
    if name in A_WAV_MD5:
        assert m["md5"] == A_WAV_MD5[name]
    rc, pcm, h, blocks = oracle.decode_whole(data)
    assert rc == 0
    assert pcm.shape == (m["channels"], m["samples"])
    assert pcm_md5(pcm) == m["pcm_md5"]
    assert [[b.num_samples, b.block_type, b.block_size] for b in blocks] == m["blocks"]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_reencodes_golden_byte_exact(name, manifest, golden_stream, oracle):
    """decode the reference stream, re-encode with the oracle: the bytes must come back."""
    data = golden_stream(name)
    m = manifest[name]
    rc, pcm, h, _ = oracle.decode_whole(data)
    assert rc == 0
    ep = capi.preset_parameter(m["preset"], m["channels"])
    rc, again, blocks, _ = oracle.encode_whole(np.ascontiguousarray(pcm), ob.make_params(m["channels"], m["bits"], m["rate"], ep))
    assert rc == 0
    assert again == data


@pytest.mark.parametrize("preset", [0, 2, 4])
def test_oracle_matches_reference_everywhere(preset, oracle, reflib, refwb):
    """streams, per-block doubles/codes/pitch/taps and the coder input, against the live reference"""
    for name, pcm, bits, rate in signal_set():
        ep = capi.preset_parameter(preset, pcm.shape[0])
        rc, ref_bytes = reflib.encode_whole(pcm, bits, rate, ep)
        assert rc == 0
        rc, wb_bytes, wb_blocks, wb_res = refwb.encode_whole(pcm, bits, rate, ep, capi.CLI_CAPACITY, True)
        assert rc == 0 and wb_bytes == ref_bytes
        rc, ora_bytes, ora_blocks, ora_res = oracle.encode_whole(pcm, ob.make_params(pcm.shape[0], bits, rate, ep), True)
        assert rc == 0
        assert ora_bytes == ref_bytes, name
        assert np.array_equal(ora_res, wb_res)
        for a, b in zip(wb_blocks, ora_blocks):
            assert (a.num_samples, a.block_type, a.block_size) == (b.num_samples, b.block_type, b.block_size)
            if a.block_type != 0:
                continue
            for ch in range(pcm.shape[0]):
                n = ep.parcor_order + 1
                assert list(a.parcor[ch])[:n] == list(b.parcor[ch])[:n]          # bit-identical doubles
                assert list(a.parcor_code[ch])[1:n] == list(b.parcor_code[ch])[1:n]   # code[0] is never written by the reference
                assert a.pitch[ch] == b.pitch[ch] and a.rshift[ch] == b.rshift[ch]
                assert a.rice_init[ch] == b.rice_init[ch]
                if a.pitch[ch] >= 3:
                    assert list(a.lt[ch])[:ep.longterm_order] == list(b.lt[ch])[:ep.longterm_order]
        rc, dec, _, _ = oracle.decode_whole(ref_bytes)
        assert rc == 0 and np.array_equal(dec, pcm)
        rc, dec2, _ = reflib.decode_whole(ora_bytes)
        assert rc == 0 and np.array_equal(dec2, pcm)


def test_crc16_known_answers(oracle):
    # CRC-16/IBM (ARC) check value and the reference's own KATs (test/test_SLAUtility.c:40-47)
    assert oracle.crc16(b"123456789") == 0xBB3D
    assert oracle.crc16(b"") == 0
    assert oracle.crc16(b"\x00") == 0
    assert oracle.crc16(b"\x01") == 0xC0C1


def test_oracle_detects_corruption(golden_stream, oracle):
    data = bytearray(golden_stream("a_wav_m2"))
    data[2000] ^= 0x40
    rc, *_ = oracle.decode_whole(bytes(data))
    assert rc == 11
    bad = bytearray(golden_stream("a_wav_m2"))
    bad[43] = 0
    rc, *_ = oracle.decode_whole(bytes(bad))
    assert rc == 10
