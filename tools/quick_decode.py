"""Scratch timing of the GPU decoder on a reference-encoded stream (dev helper, not the bench)."""
import ctypes as C, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sla_b200 import capi, synth
from oracle import binding as ob
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 60
ref = ob.reference_library()
lib = capi.SLALibrary("sla_b200/lib/libsla_b200.so")
pcm = synth.synth_pcm(2, 44100 * secs, 16, 44100, 0)
ep = capi.preset_parameter(2, 2)
t0 = time.time(); rc, data = ref.encode_whole(pcm, 16, 44100, ep); t1 = time.time()
print(f"ref encode rc={rc} {len(data)} B in {t1-t0:.2f}s = {pcm.size/(t1-t0)/1e6:.2f} Msmp/s")
t0 = time.time(); rc, dec, h = ref.decode_whole(data); t1 = time.time()
print(f"ref decode rc={rc} in {t1-t0:.2f}s = {pcm.size/(t1-t0)/1e6:.2f} Msmp/s")
for it in range(3):
    t0 = time.time(); rc, dec, h = lib.decode_whole(data); t1 = time.time()
    print(f"gpu decode rc={rc} exact={np.array_equal(dec, pcm)} wall {t1-t0:.3f}s")
# handle reuse for timing
cfg = capi.DecoderConfig(**capi.CLI_CAPACITY, enable_crc_check=1, verpose_flag=0)
L = lib.lib
dec_h = L.SLADecoder_Create(C.byref(cfg))
out = np.zeros_like(pcm); got = C.c_uint32(0)
buf = np.frombuffer(data, dtype=np.uint8)
ptrs = capi._planar_pointers(out)
ms = (C.c_float * 3)(); nl = C.c_uint32(0)
L.SLAB200_Decoder_LastTiming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
for it in range(4):
    t0 = time.time(); rc = L.SLADecoder_DecodeWhole(dec_h, buf.ctypes.data, len(data), ptrs, pcm.shape[1], C.byref(got)); t1 = time.time()
    L.SLAB200_Decoder_LastTiming(dec_h, ms, C.byref(nl))
    print(f"reuse: rc={rc} wall {1e3*(t1-t0):.2f} ms  h2d {ms[0]:.2f} kernels {ms[1]:.2f} d2h {ms[2]:.2f} ms launches {nl.value} -> kernels-only {pcm.size/ms[1]/1e3:.1f} Msmp/s")
L.SLADecoder_Destroy(dec_h)
