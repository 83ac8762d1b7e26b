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


def _roundtrip(lib, monkeypatch, presets, chunks, quick=False):
    cases = _cases()
    if quick:       # the host simulator is slow: one case per sample width, plus the lshift > 0 file
        cases = [x for x in cases if x[0] in ("s16_special", "s24_impulsive", "u8_stereo", "s32_mono")]
    for name, planar, bits, rate in cases:
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
    _roundtrip(hostsim, monkeypatch, presets=(2,), chunks=(1,), quick=True)


@pytest.mark.gpu
def test_gpu_pcm_roundtrip(product, monkeypatch):
    _roundtrip(product, monkeypatch, presets=(0, 2, 4), chunks=(0, 1, 30000))


# ---- batch decode (BASELINE config 5: many short files, mixed presets) ---------------------------
GOLDEN_NAMES = ["a_wav_m0", "a_wav_m2", "a_wav_m4", "s16_special_m2", "s16_special_m0", "s24_impulsive_m4", "ch8_24bit_m2"]


def _batch(lib, golden_stream, oracle):
    from oracle import binding as ob
    streams = [golden_stream(n) for n in GOLDEN_NAMES]
    # more files of every preset, several per parameter class so that groups hold more than one file
    for k, preset in enumerate((0, 2, 4, 2, 0, 3, 1)):
        pcm = synth.synth_pcm(2, 20000 + 3000 * k, 16, 44100, 40 + k)
        ep = capi.preset_parameter(preset, 2)
        rc, data, _, _ = oracle.encode_whole(pcm, ob.make_params(2, 16, 44100, ep))
        assert rc == 0
        streams.append(data)
    good = len(streams)
    bad = bytearray(streams[1]); bad[2000] ^= 0x40; streams.append(bytes(bad))           # CRC failure
    streams.append(streams[3][:len(streams[3]) // 2])                                    # truncated
    bad = bytearray(streams[7]); bad[43] = 0; streams.append(bytes(bad))                 # no sync code
    streams.append(b"XLA*" + streams[0][4:])                                             # not an .sla file
    capacities = [None] * len(streams)
    streams.append(streams[8]); capacities.append(1000)                                  # output buffer too small
    rc, res = capi.decode_batch_pcm(lib, streams, capacities=capacities)
    assert rc == capi.OK
    for i, data in enumerate(streams):
        if i < good:
            rc1, want, _ = capi.decode_pcm(lib, data)
            assert rc1 == capi.OK
            assert res[i][0] == capi.OK and res[i][1] == want, i
        else:
            rc1, _, _ = lib.decode_whole(data, out_samples=capacities[i])
            assert res[i][0] == rc1 and rc1 != capi.OK, (i, res[i][0], rc1)


def test_hostsim_batch_decode(hostsim, golden_stream, oracle):
    _batch(hostsim, golden_stream, oracle)


@pytest.mark.gpu
def test_gpu_batch_decode(product, golden_stream, oracle):
    _batch(product, golden_stream, oracle)



def _batch_encode(lib, nfiles, nmax):
    """SLAB200_Encoder_EncodeBatchPCM: each stream equals the single-file call's; per-item result codes."""
    rng = np.random.default_rng(77)
    for nch, bits, rate, preset in ((2, 16, 44100, 2), (1, 24, 48000, 4)):
        ep = capi.preset_parameter(preset, nch)
        pcms = []
        for i in range(nfiles):
            n = int(rng.integers(2000, nmax)) if i != 2 else 0                  # one empty file
            planar = np.ascontiguousarray(synth.synth_pcm(nch, max(n, 1), bits, rate, 500 + i, specials=(i % 2 == 0)))[:, :n]
            if i == 1:
                planar = (planar >> 20) << 20                                   # its own offset_lshift
            pcms.append(capi.planar_to_pcm(np.ascontiguousarray(planar), bits))
        want = []
        for pcm in pcms:
            rc, s = capi.encode_pcm(lib, pcm, nch, bits, rate, ep)
            assert rc == capi.OK
            want.append(s)
        caps = [None] * nfiles
        caps[3] = len(want[3]) - 1                                              # too small for file 3 only
        rc, got = capi.encode_batch_pcm(lib, pcms, nch, bits, rate, ep, out_capacities=caps)
        assert rc == capi.OK
        for i in range(nfiles):
            if i == 3:
                assert got[i][0] == capi.INSUFFICIENT_BUFFER_SIZE
            else:
                assert got[i][0] == capi.OK and got[i][1] == want[i], (nch, bits, i)
    rc, got = capi.encode_batch_pcm(lib, [], 2, 16, 44100, capi.preset_parameter(2, 2))
    assert rc == capi.OK and got == []
    rc, _ = capi.encode_batch_pcm(lib, [b"\0" * 64], 2, 12, 44100, capi.preset_parameter(2, 2))
    assert rc == capi.INVALID_ARGUMENT


def test_hostsim_batch_encode(hostsim):
    _batch_encode(hostsim, 5, 9000)


def test_hostsim_batch_encode_groups(hostsim, monkeypatch):
    """the same corpus cut into several merged groups (one, two and three files per launch sequence)"""
    monkeypatch.setenv("SLAB200_BATCH_ENC_FRAMES", "50000")
    _batch_encode(hostsim, 6, 30000)          # files that end in a segment below the minimum block, followed by other files


def _batch_encode_formats(lib):
    """merged batch over other formats: 3 channels, 8-bit and 32-bit, full-scale noise (RAW blocks) next to tonal
    files, one noise file with an offset_lshift of its own - every stream equals the single-file call's"""
    rng = np.random.default_rng(5)
    for nch, bits, rate, preset in ((3, 16, 44100, 0), (2, 8, 22050, 1), (2, 32, 48000, 3)):
        ep = capi.preset_parameter(preset, nch)
        pcms = []
        for i in range(4):
            n = int(rng.integers(3000, 12000))
            if i % 2 == 0:
                planar = rng.integers(-(1 << 31), (1 << 31) - 1, size=(nch, n), dtype=np.int64).astype(np.int32)
                planar = (planar >> (32 - bits)) << (32 - bits)
            else:
                planar = np.ascontiguousarray(synth.synth_pcm(nch, n, bits, rate, 40 + i))
            if i == 2 and bits > 8:
                planar = (planar >> (32 - bits + 3)) << (32 - bits + 3)
            pcms.append(capi.planar_to_pcm(np.ascontiguousarray(planar), bits))
        want = [capi.encode_pcm(lib, p, nch, bits, rate, ep) for p in pcms]
        rc, got = capi.encode_batch_pcm(lib, pcms, nch, bits, rate, ep)
        assert rc == capi.OK
        for i in range(len(pcms)):
            assert want[i][0] == capi.OK and got[i][0] == capi.OK and got[i][1] == want[i][1], (nch, bits, i)


def test_hostsim_batch_encode_formats(hostsim):
    _batch_encode_formats(hostsim)


@pytest.mark.gpu
def test_gpu_batch_encode_formats(product):
    _batch_encode_formats(product)


def test_hostsim_batch_encode_long_files(hostsim, monkeypatch):
    """files alone in their group take the plain single-file job instead of the merged one"""
    monkeypatch.setenv("SLAB200_BATCH_ENC_FRAMES", "1024")
    monkeypatch.setenv("SLAB200_BATCH_LONG_FRAMES", "1")
    _batch_encode(hostsim, 5, 9000)


@pytest.mark.gpu
def test_gpu_batch_encode(product, monkeypatch):
    _batch_encode(product, 24, 120000)
    monkeypatch.setenv("SLAB200_BATCH_ENC_FRAMES", "400000")            # several groups on several contexts
    _batch_encode(product, 24, 120000)
    monkeypatch.setenv("SLAB200_BATCH_ENC_FRAMES", "1024")              # every file alone: the plain single-file job
    monkeypatch.setenv("SLAB200_BATCH_LONG_FRAMES", "1")
    _batch_encode(product, 24, 120000)
