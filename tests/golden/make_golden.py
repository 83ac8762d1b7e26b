"""Regenerates tests/golden/*.sla and manifest.json from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference for a.wav and oracle/_ref/libsla_ref.so):
    python tests/golden/make_golden.py
The streams are outputs of the reference encoder; manifest.json records the parameters, sizes,
md5 of each stream and md5 of the PCM the reference decoder returns for it.
"""
import hashlib
import json
import os
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import binding as ob          # noqa: E402
from sla_b200 import capi, synth          # noqa: E402


def main():
    ref = ob.reference_library()
    wb = ob.RefWhitebox()
    cases = []
    w = wave.open("/root/reference/test/a.wav")
    raw = np.frombuffer(w.readframes(w.getnframes()), dtype=np.uint8)
    a = ((raw.astype(np.int32) - 128) << 24).reshape(1, -1).copy()
    for preset in (0, 2, 4):
        cases.append((f"a_wav_m{preset}", a, 8, 48000, preset))
    s16 = synth.synth_pcm(2, 60000, 16, 44100, 0, clear_low_bits=4)
    cases.append(("s16_special_m2", s16, 16, 44100, 2))
    cases.append(("s16_special_m0", s16, 16, 44100, 0))
    s24 = synth.impulsive_24bit(30000)
    cases.append(("s24_impulsive_m4", s24, 24, 96000, 4))
    s8 = synth.synth_pcm(8, 20000, 24, 48000, 3, specials=False)
    cases.append(("ch8_24bit_m2", s8, 24, 48000, 2))
    manifest = {}
    for name, pcm, bits, rate, preset in cases:
        ep = capi.preset_parameter(preset, pcm.shape[0])
        rc, data = ref.encode_whole(pcm, bits, rate, ep)
        assert rc == 0
        rc, dec, _ = ref.decode_whole(data)
        assert rc == 0 and np.array_equal(dec, pcm)
        _, _, blocks, _ = wb.encode_whole(pcm, bits, rate, ep, capi.CLI_CAPACITY)
        with open(os.path.join(HERE, name + ".sla"), "wb") as f:
            f.write(data)
        manifest[name] = dict(
            preset=preset, channels=int(pcm.shape[0]), samples=int(pcm.shape[1]), bits=bits, rate=rate,
            size=len(data), md5=hashlib.md5(data).hexdigest(),
            pcm_md5=hashlib.md5(np.ascontiguousarray(pcm).tobytes()).hexdigest(),
            blocks=[[int(b.num_samples), int(b.block_type), int(b.block_size)] for b in blocks])
        print(name, len(data), manifest[name]["md5"])
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
