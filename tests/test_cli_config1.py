"""BASELINE.json config 1: the reference's own CLI (unmodified main.c / wav.c, built by oracle/Makefile
into oracle/_ref/sla_gpu_cli) linked against libsla_b200.so encodes and decodes test/a.wav on the GPU;
the .sla bytes must equal the reference CLI's (md5 in BASELINE.md) and the WAV must round-trip."""
import hashlib
import os
import subprocess
import wave

import numpy as np
import pytest

from conftest import ROOT

CLI = os.path.join(ROOT, "oracle", "_ref", "sla_gpu_cli")
A_WAV_MD5 = {0: "48c60a59f94f70303be8207d7ea9dc03", 2: "9739dfd1acd3eeaec7a3f4345ee8c4a4",
             4: "9ad138cb6ad58ab8b074eae1132b3c28"}


@pytest.mark.gpu
@pytest.mark.parametrize("preset", [0, 2, 4])
def test_reference_cli_on_gpu_library(preset, tmp_path, golden_stream, oracle, product):
    if not os.path.exists(CLI):
        pytest.skip("oracle/_ref/sla_gpu_cli not built (needs /root/reference at build time)")
    rc, pcm, _, _ = oracle.decode_whole(golden_stream(f"a_wav_m{preset}"))
    assert rc == 0
    wav_in = tmp_path / "a.wav"
    with wave.open(str(wav_in), "wb") as w:                       # 8-bit mono 48 kHz, as test/a.wav
        w.setnchannels(1); w.setsampwidth(1); w.setframerate(48000)
        w.writeframes(((pcm[0] >> 24) + 128).astype(np.uint8).tobytes())
    sla = tmp_path / "a.sla"
    subprocess.run([CLI, "-e", "-m", str(preset), str(wav_in), str(sla)], check=True, stdout=subprocess.DEVNULL)
    data = sla.read_bytes()
    assert hashlib.md5(data).hexdigest() == A_WAV_MD5[preset]
    assert data == golden_stream(f"a_wav_m{preset}")
    wav_out = tmp_path / "b.wav"
    subprocess.run([CLI, "-d", str(sla), str(wav_out)], check=True, stdout=subprocess.DEVNULL)
    with wave.open(str(wav_out)) as w:
        back = np.frombuffer(w.readframes(w.getnframes()), dtype=np.uint8)
    assert np.array_equal(back, ((pcm[0] >> 24) + 128).astype(np.uint8))
    # streaming mode of the same unmodified CLI (src/main.c:275-420) over SLAStreamingDecoder_*
    wav_strm = tmp_path / "c.wav"
    subprocess.run([CLI, "-d", "-s", str(sla), str(wav_strm)], check=True, stdout=subprocess.DEVNULL)
    assert wav_strm.read_bytes() == wav_out.read_bytes()
