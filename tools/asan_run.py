"""Every kernel path on the AddressSanitizer build of the host simulator (make hostsim-asan): heap = the simulated device\nmemory, so an out-of-bounds access of any kernel is reported.  compute-sanitizer is closed on the GPU pool.\nusage: LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_run.py tests/hostsim/libsla_hostsim_asan.so"""
import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from sla_b200 import capi, synth
from conftest import signal_set
lib = capi.SLALibrary(sys.argv[1] if len(sys.argv) > 1 else '/tmp/libsla_hostsim_asan.so')
streams = []
for name, pcm, bits, rate in signal_set():
    pcm = np.ascontiguousarray(pcm)
    for preset in (2, 4):
        ep = capi.preset_parameter(preset, pcm.shape[0])
        rc, data = lib.encode_whole(pcm, bits, rate, ep); assert rc == 0
        rc, dec, _ = lib.decode_whole(data); assert rc == 0 and np.array_equal(dec, pcm)
        rc, dec, _ = lib.decode_whole_device(data); assert rc == 0 and np.array_equal(dec, pcm)
        os.environ["SLAB200_PIPE_CHUNK_SAMPLES"] = "1"; os.environ["SLAB200_PIPE_DEC_CHUNKS"] = "3"
        raw = capi.planar_to_pcm(pcm, bits)
        rc, d2 = capi.encode_pcm(lib, raw, pcm.shape[0], bits, rate, ep); assert rc == 0 and d2 == data
        rc, back, _ = capi.decode_pcm(lib, data); assert rc == 0 and back == raw
        os.environ.pop("SLAB200_PIPE_CHUNK_SAMPLES"); os.environ.pop("SLAB200_PIPE_DEC_CHUNKS")
        streams.append(data)
    print(name, "ok", flush=True)
rc, res = capi.decode_batch_pcm(lib, streams); assert rc == 0 and all(r == 0 for r, _ in res)
bad = bytearray(streams[2]); bad[len(bad)//2] ^= 0x55
for crc in (True, False):
    lib.decode_whole(bytes(bad), crc=crc); lib.decode_whole_device(bytes(bad), crc=crc)
lib.decode_whole(streams[2][:len(streams[2])//3])
print("asan run ok")
