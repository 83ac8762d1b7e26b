"""Scratch timing of the GPU encoder (dev helper, not the bench)."""
import ctypes as C, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sla_b200 import capi, synth
from oracle import binding as ob
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 60
preset = int(sys.argv[2]) if len(sys.argv) > 2 else 2
check = len(sys.argv) > 3
lib = capi.SLALibrary("sla_b200/lib/libsla_b200.so")
L = lib.lib
pcm = synth.synth_pcm(2, 44100 * secs, 16, 44100, 0)
ep = capi.preset_parameter(preset, 2)
if check:
    ref = ob.reference_library()
    t0 = time.time(); rc, want = ref.encode_whole(pcm, 16, 44100, ep); t1 = time.time()
    print(f"ref encode rc={rc} {len(want)} B in {t1-t0:.2f}s = {pcm.size/(t1-t0)/1e6:.2f} Msmp/s")
cfg = capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)
enc = L.SLAEncoder_Create(C.byref(cfg))
wf = capi.WaveFormat(2, 16, 44100, 0)
L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)); L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep))
cap = 43 + pcm.size * 4 + 65536
out = np.zeros(cap, dtype=np.uint8); size = C.c_uint32(0)
ptrs = capi._planar_pointers(pcm)
ms = (C.c_float * 3)(); nl = C.c_uint32(0)
L.SLAB200_Encoder_LastTiming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
for it in range(4):
    t0 = time.time(); rc = L.SLAEncoder_EncodeWhole(enc, ptrs, pcm.shape[1], out.ctypes.data, cap, C.byref(size)); t1 = time.time()
    L.SLAB200_Encoder_LastTiming(enc, ms, C.byref(nl))
    print(f"gpu encode rc={rc} {size.value} B wall {1e3*(t1-t0):.1f} ms  h2d {ms[0]:.2f} kernels {ms[1]:.2f} d2h {ms[2]:.2f} launches {nl.value} -> kernels-only {pcm.size/ms[1]/1e3:.1f} Msmp/s")
if check:
    print("identical to reference:", out[:size.value].tobytes() == want)
L.SLAEncoder_Destroy(enc)
