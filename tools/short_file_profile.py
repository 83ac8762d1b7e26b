"""Where the time of one short file goes: per-kernel device times of SLAEncoder_EncodeWhole / SLADecoder_DecodeWhole
for a 16 s stereo file (dev helper).  usage: python tools/short_file_profile.py [seconds] [preset]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sla_b200 import capi, synth
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
preset = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = capi.SLALibrary("sla_b200/lib/libsla_b200.so"); L = lib.lib
pcm = synth.synth_pcm(2, 44100 * secs, 16, 44100, 7)
ep = capi.preset_parameter(preset, 2)
enc = L.SLAEncoder_Create(C.byref(capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)))
wf = capi.WaveFormat(2, 16, 44100, 0)
L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)); L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep))
cap = 43 + pcm.size * 4 + 65536
out = np.zeros(cap, dtype=np.uint8); size = C.c_uint32(0)
ptrs = capi._planar_pointers(pcm)
L.SLAB200_Encoder_EnableProfile.argtypes = [C.c_void_p, C.c_int]
L.SLAB200_Encoder_GetProfile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
L.SLAB200_Encoder_GetProfile.restype = C.c_uint32
for it in range(3):
    t0 = time.perf_counter(); rc = L.SLAEncoder_EncodeWhole(enc, ptrs, pcm.shape[1], out.ctypes.data, cap, C.byref(size)); t1 = time.perf_counter()
    print(f"encode {secs} s file: wall {1e3 * (t1 - t0):.2f} ms rc={rc}")
L.SLAB200_Encoder_EnableProfile(enc, 1)
L.SLAEncoder_EncodeWhole(enc, ptrs, pcm.shape[1], out.ctypes.data, cap, C.byref(size))
names = (C.c_char_p * 256)(); ms = (C.c_float * 256)()
n = L.SLAB200_Encoder_GetProfile(enc, names, ms, 256)
tot = 0.0
for i in range(n):
    print(f"  {names[i].decode():28s} {ms[i]:8.3f} ms"); tot += ms[i]
print(f"  kernels total {tot:.3f} ms over {n} launches")
L.SLAEncoder_Destroy(enc)
