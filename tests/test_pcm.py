"""Raw-PCM entry points: interleaved little-endian PCM in, .sla out, and back.
The stream must be byte-identical to SLAEncoder_EncodeWhole on the converted planes (and therefore to
the reference), the decoded PCM byte-identical to the input; conversions follow src/wav.c:392-417,630-668.
"""
import numpy as np
import pytest

from conftest import signal_set
from sla_b200 import capi, synth


def _cases():
    out = list(signal_set())
    out.append(("u8_stereo", synth.synth_pcm(2, 30000, 8, 22050, 31, specials=False), 8, 22050))
    out.append(("s32_mono", synth.synth_pcm(1, 25000, 32, 48000, 32, specials=False), 32, 48000))
    return out


def _roundtrip(lib, monkeypatch, presets, chunks):
    for name, planar, bits, rate in _cases():
        planar = np.ascontiguousarray(planar)
        nch = planar.shape[0]
        pcm = capi.planar_to_pcm(planar, bits)
        assert np.array_equal(capi.pcm_to_planar(pcm, bits, nch), planar)       # the helpers themselves
        for preset in presets:
            ep = capi.preset_parameter(preset, nch)
            monkeypatch.delenv("SLAB200_PIPE_CHUNK_SAMPLES", raising=False)
            rc, want = lib.encode_whole(planar, bits, rate, ep)
            assert rc == capi.OK
            # what the int32 API decodes to: the input itself, except for 32-bit material, which the
            # reference format does not round-trip either (its predictors work in wrapping int32)
            rc, dec, _ = lib.decode_whole(want)
            assert rc == capi.OK and (bits == 32 or np.array_equal(dec, planar))
            want_pcm = capi.planar_to_pcm(np.ascontiguousarray(dec), bits)
            for chunk in chunks:
                if chunk:
                    monkeypatch.setenv("SLAB200_PIPE_CHUNK_SAMPLES", str(chunk))
                    monkeypatch.setenv("SLAB200_PIPE_DEC_CHUNKS", "3")
                else:
                    monkeypatch.delenv("SLAB200_PIPE_CHUNK_SAMPLES", raising=False)
                    monkeypatch.delenv("SLAB200_PIPE_DEC_CHUNKS", raising=False)
                rc, got = capi.encode_pcm(lib, pcm, nch, bits, rate, ep)
                assert rc == capi.OK and got == want, (name, preset, chunk)
                rc, back, h = capi.decode_pcm(lib, want)
                assert rc == capi.OK and back == want_pcm, (name, preset, chunk)
            rc, _ = capi.encode_pcm(lib, pcm, nch, bits, rate, ep, out_capacity=len(want) - 1)
            assert rc == capi.INSUFFICIENT_BUFFER_SIZE
    monkeypatch.delenv("SLAB200_PIPE_CHUNK_SAMPLES", raising=False)
    monkeypatch.delenv("SLAB200_PIPE_DEC_CHUNKS", raising=False)
    # a 12-bit wave format has no PCM layout
    rc, _ = capi.encode_pcm(lib, b"\0" * 64, 2, 12, 44100, capi.preset_parameter(2, 2))
    assert rc == capi.INVALID_ARGUMENT


def test_hostsim_pcm_roundtrip(hostsim, monkeypatch):
    _roundtrip(hostsim, monkeypatch, presets=(2,), chunks=(0, 1))


@pytest.mark.gpu
def test_gpu_pcm_roundtrip(product, monkeypatch):
    _roundtrip(product, monkeypatch, presets=(0, 2, 4), chunks=(0, 1, 30000))
