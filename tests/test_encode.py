"""GPU encode parity (SLAEncoder_EncodeWhole through the C ABI) against the oracle / reference.

Bar (BASELINE.json north_star): encoded bytes byte-identical wherever the quantised coefficients
match; pre-quantisation PARCOR doubles within 1e-9 relative; any block whose quantised coefficients
differ is listed with its size delta; every stream must decode bit-exactly with the reference decoder.
The same checks run on the host-simulator build of the kernels (no GPU) and on the B200 (-m gpu).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import signal_set
from oracle import binding as ob
from sla_b200 import capi

REL_TOL = 1e-9      # north_star tolerance for pre-quantisation double coefficients


def encode_with_export(lib, pcm, bits, rate, ep):
    """EncodeWhole + the debug export of per-block intermediates."""
    L = lib.lib
    L.SLAB200_Encoder_SetDebugExport.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    cfg = capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)
    enc = L.SLAEncoder_Create(C.byref(cfg))
    assert enc, "SLAEncoder_Create failed"
    try:
        wf = capi.WaveFormat(pcm.shape[0], bits, rate, 0)
        assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == capi.OK
        assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == capi.OK
        n = pcm.shape[1]
        maxb = n // 2048 + 16
        recs = (ob.Block * maxb)()
        res = np.zeros_like(pcm)
        res_ptrs = capi._planar_pointers(res)
        L.SLAB200_Encoder_SetDebugExport(enc, recs, maxb, res_ptrs)
        cap = 43 + 2 * pcm.size * max(bits // 8, 1) + 65536
        out = np.zeros(cap, dtype=np.uint8)
        size = C.c_uint32(0)
        rc = L.SLAEncoder_EncodeWhole(enc, capi._planar_pointers(pcm), n, out.ctypes.data, cap, C.byref(size))
        assert rc == capi.OK
        data = out[:size.value].tobytes()
        rc, h = lib.decode_header(data)
        assert rc == capi.OK
        return data, list(recs[:h.num_blocks]), res
    finally:
        L.SLAEncoder_Destroy(enc)


def compare_with_oracle(lib, oracle, pcm, bits, rate, preset, reflib=None):
    """Returns the list of mismatching blocks [(index, size_delta)]; asserts everything else."""
    ep = capi.preset_parameter(preset, pcm.shape[0])
    rc, want, want_blocks, want_res = oracle.encode_whole(pcm, ob.make_params(pcm.shape[0], bits, rate, ep), True)
    assert rc == 0
    data, blocks, res = encode_with_export(lib, pcm, bits, rate, ep)
    P, T = ep.parcor_order, ep.longterm_order
    assert [(b.sample_offset, b.num_samples) for b in blocks] == \
           [(b.sample_offset, b.num_samples) for b in want_blocks], "block partition differs"
    mismatched = []
    for i, (a, b) in enumerate(zip(blocks, want_blocks)):
        coeffs_match = a.block_type == b.block_type
        if a.block_type == 0 and b.block_type == 0:
            for ch in range(pcm.shape[0]):
                pa, pb = np.array(a.parcor[ch][1:P + 1]), np.array(b.parcor[ch][1:P + 1])
                rel = np.abs(pa - pb) / np.maximum(np.abs(pb), 1e-300)
                assert np.all((rel <= REL_TOL) | (np.abs(pa - pb) <= 1e-15)), \
                    f"block {i} ch {ch}: PARCOR doubles differ by {rel.max():.3e} relative"
                same = (list(a.parcor_code[ch])[1:P + 1] == list(b.parcor_code[ch])[1:P + 1]
                        and a.rshift[ch] == b.rshift[ch] and a.pitch[ch] == b.pitch[ch]
                        and (a.pitch[ch] < 3 or list(a.lt_q31[ch])[:T] == list(b.lt_q31[ch])[:T]))
                coeffs_match = coeffs_match and same
                if same:
                    s, e = a.sample_offset, a.sample_offset + a.num_samples
                    assert np.array_equal(res[ch, s:e], want_res[ch, s:e]), f"block {i} ch {ch}: residual differs"
                    assert a.rice_init[ch] == b.rice_init[ch]
        if coeffs_match:
            assert a.block_size == b.block_size, f"block {i}: size differs with equal coefficients"
            assert data[a.byte_offset:a.byte_offset + a.block_size] == \
                want[b.byte_offset:b.byte_offset + b.block_size], f"block {i}: bytes differ with equal coefficients"
        else:
            mismatched.append((i, int(a.block_size) - int(b.block_size)))
    if not mismatched:
        assert data == want
    assert abs(len(data) - len(want)) <= 0.001 * len(want)          # compression ratio within 0.1 %
    # round trip through the independent decoders
    rc, dec, _, _ = oracle.decode_whole(data)
    assert rc == 0 and np.array_equal(dec, pcm)
    if reflib is not None:
        rc, dec, _ = reflib.decode_whole(data)
        assert rc == capi.OK and np.array_equal(dec, pcm)
    return mismatched


@pytest.mark.parametrize("preset", [0, 2, 4])
def test_hostsim_encode_matches_oracle(preset, hostsim, oracle, reflib):
    for name, pcm, bits, rate in signal_set():
        mismatched = compare_with_oracle(hostsim, oracle, pcm, bits, rate, preset, reflib)
        assert mismatched == [], f"{name}: blocks with different quantised coefficients {mismatched}"


@pytest.mark.gpu
@pytest.mark.parametrize("preset", [0, 1, 2, 3, 4])
def test_gpu_encode_matches_oracle(preset, product, oracle, reflib):
    report = {}
    for name, pcm, bits, rate in signal_set():
        report[name] = compare_with_oracle(product, oracle, pcm, bits, rate, preset, reflib)
    # every deviation must be listed; on these fixtures none is expected
    assert all(v == [] for v in report.values()), report


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a_wav_m0", "a_wav_m2", "a_wav_m4", "s16_special_m2", "s16_special_m0",
                                  "s24_impulsive_m4", "ch8_24bit_m2"])
def test_gpu_reencodes_golden_byte_exact(name, manifest, golden_stream, product, oracle):
    """decode the committed reference stream, re-encode on the GPU: the reference bytes must come back"""
    data = golden_stream(name)
    m = manifest[name]
    rc, pcm, _, _ = oracle.decode_whole(data)
    assert rc == 0
    ep = capi.preset_parameter(m["preset"], m["channels"])
    rc, again = product.encode_whole(np.ascontiguousarray(pcm), m["bits"], m["rate"], ep)
    assert rc == capi.OK
    assert again == data


def _custom_parameter_roundtrip(lib, oracle, reflib):
    """parameter sets outside the CLI presets: they take the generic (non-specialised) kernels"""
    from conftest import multi_silence
    from sla_b200 import synth
    cases = [
        (capi.EncodeParameter(40, 5, 16, capi.CH_STEREO_MS, capi.WIN_HANN, 8192), synth.synth_pcm(2, 40000, 16, 44100, 9), 16, 44100),
        (capi.EncodeParameter(4, 1, 4, capi.CH_NONE, capi.WIN_BLACKMAN, 2048), synth.synth_pcm(1, 20000, 24, 48000, 10), 24, 48000),
        (capi.EncodeParameter(48, 3, 32, capi.CH_NONE, capi.WIN_VORBIS, 16384), synth.synth_pcm(3, 36000, 16, 32000, 12), 16, 32000),
        (capi.EncodeParameter(16, 1, 8, capi.CH_STEREO_MS, capi.WIN_RECT, 10000), multi_silence(), 16, 44100),
    ]
    for ep, pcm, bits, rate in cases:
        rc, want, _, _ = oracle.encode_whole(pcm, ob.make_params(pcm.shape[0], bits, rate, ep))
        assert rc == 0
        rc, got = lib.encode_whole(pcm, bits, rate, ep)
        assert rc == capi.OK
        assert got == want, (ep.parcor_order, ep.longterm_order, ep.lms_order_per_filter)
        rc, dec, _ = lib.decode_whole(want)
        assert rc == capi.OK and np.array_equal(dec, pcm)
        rc, dec, _ = reflib.decode_whole(got)
        assert rc == capi.OK and np.array_equal(dec, pcm)


def test_hostsim_custom_parameters(hostsim, oracle, reflib):
    _custom_parameter_roundtrip(hostsim, oracle, reflib)


@pytest.mark.gpu
def test_gpu_custom_parameters(product, oracle, reflib):
    _custom_parameter_roundtrip(product, oracle, reflib)


# ---- pipelined whole-file encode (chunks on several contexts): same bytes as the single pass ----
def _pipelined_identical(lib, monkeypatch, presets, quick=False):
    from conftest import multi_silence
    from sla_b200 import synth
    signals = signal_set() + [("long_silences", np.concatenate([multi_silence(), multi_silence()[:, ::-1],
                                                                  synth.synth_pcm(2, 70000, 16, 44100, 21)], axis=1), 16, 44100)]
    if quick:       # the host simulator is slow: the silence re-basing cases and one odd-length mono file
        signals = [x for x in signals if x[0] in ("multi_silence", "mono16_odd")] + \
                  [("silences2", np.concatenate([multi_silence(), multi_silence()[:, ::-1]], axis=1), 16, 44100)]
    for preset in presets:
        for name, pcm, bits, rate in signals:
            pcm = np.ascontiguousarray(pcm)
            ep = capi.preset_parameter(preset, pcm.shape[0])
            monkeypatch.delenv("SLAB200_PIPE_CHUNK_SAMPLES", raising=False)
            rc, want = lib.encode_whole(pcm, bits, rate, ep)
            assert rc == capi.OK
            for chunk in (1, 30000):                  # 1 -> one block per chunk; 30000 -> a few blocks
                monkeypatch.setenv("SLAB200_PIPE_CHUNK_SAMPLES", str(chunk))
                rc, got = lib.encode_whole(pcm, bits, rate, ep)
                assert rc == capi.OK and got == want, (name, preset, chunk)
                # and a too-small output buffer is still reported
                rc, _ = lib.encode_whole(pcm, bits, rate, ep, out_capacity=len(want) - 1)
                assert rc == capi.INSUFFICIENT_BUFFER_SIZE, (name, preset, chunk)
    monkeypatch.delenv("SLAB200_PIPE_CHUNK_SAMPLES", raising=False)


def test_hostsim_pipelined_encode_identical(hostsim, monkeypatch):
    _pipelined_identical(hostsim, monkeypatch, presets=(2,), quick=True)


@pytest.mark.gpu
def test_gpu_pipelined_encode_identical(product, monkeypatch):
    _pipelined_identical(product, monkeypatch, presets=(0, 2, 4))


def _device_pipelined_identical(lib, monkeypatch, use_torch, presets, quick=False):
    monkeypatch.setenv("SLAB200_PIPE_DEVICE", "1")          # chunked device-resident encode is opt-in
    for preset in presets:
        for name, pcm, bits, rate in (signal_set()[:1] if quick else signal_set()):
            pcm = np.ascontiguousarray(pcm)
            ep = capi.preset_parameter(preset, pcm.shape[0])
            monkeypatch.delenv("SLAB200_PIPE_CHUNK_SAMPLES", raising=False)
            rc, want = lib.encode_whole(pcm, bits, rate, ep)
            assert rc == capi.OK
            for chunk in (None, 1, 30000):
                if chunk is None:
                    monkeypatch.delenv("SLAB200_PIPE_CHUNK_SAMPLES", raising=False)
                else:
                    monkeypatch.setenv("SLAB200_PIPE_CHUNK_SAMPLES", str(chunk))
                rc, got = capi.encode_whole_device(lib, pcm, bits, rate, ep, use_torch=use_torch)
                assert rc == capi.OK and got == want, (name, preset, chunk)
    monkeypatch.delenv("SLAB200_PIPE_CHUNK_SAMPLES", raising=False)
    monkeypatch.delenv("SLAB200_PIPE_DEVICE", raising=False)


def test_hostsim_device_encode_pipelined(hostsim, monkeypatch):
    _device_pipelined_identical(hostsim, monkeypatch, False, presets=(2,), quick=True)


@pytest.mark.gpu
def test_gpu_device_encode_pipelined(product, monkeypatch):
    _device_pipelined_identical(product, monkeypatch, True, presets=(0, 2, 4))

