"""Small end-to-end exercise of every kernel path for compute-sanitizer (memcheck).
usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from sla_b200 import capi, synth
from conftest import signal_set
lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so"))
streams = []
for name, pcm, bits, rate in signal_set():
    pcm = np.ascontiguousarray(pcm)
    for preset in (0, 2, 4):
        ep = capi.preset_parameter(preset, pcm.shape[0])
        rc, data = lib.encode_whole(pcm, bits, rate, ep)
        assert rc == 0, (name, preset, rc)
        rc, dec, _ = lib.decode_whole(data)
        assert rc == 0 and np.array_equal(dec, pcm), (name, preset)
        rc, dec, _ = lib.decode_whole_device(data, use_torch=True)
        assert rc == 0 and np.array_equal(dec, pcm), (name, preset)
        os.environ["SLAB200_PIPE_CHUNK_SAMPLES"] = "1"; os.environ["SLAB200_PIPE_DEC_CHUNKS"] = "3"
        raw = capi.planar_to_pcm(pcm, bits)
        rc, d2 = capi.encode_pcm(lib, raw, pcm.shape[0], bits, rate, ep)
        assert rc == 0 and d2 == data, (name, preset)
        rc, back, _ = capi.decode_pcm(lib, data)
        assert rc == 0 and back == raw, (name, preset)
        os.environ.pop("SLAB200_PIPE_CHUNK_SAMPLES"); os.environ.pop("SLAB200_PIPE_DEC_CHUNKS")
        streams.append(data)
rc, res = capi.decode_batch_pcm(lib, streams)
assert rc == 0 and all(r == 0 for r, _ in res)
# damaged streams must not touch memory they do not own either
bad = bytearray(streams[4]); bad[len(bad) // 2] ^= 0x55
for crc in (True, False):
    lib.decode_whole(bytes(bad), crc=crc)
    lib.decode_whole_device(bytes(bad), crc=crc, use_torch=True)
lib.decode_whole(streams[4][:len(streams[4]) // 3])
print("sanitize run ok:", len(streams), "streams")
