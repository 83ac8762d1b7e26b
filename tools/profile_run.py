"""One device-resident encode + decode of a synthetic file: the command profiled under ncu.
usage: python tools/profile_run.py [seconds] [preset] [repeats]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sla_b200 import capi, synth
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 600
preset = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so"))
L = lib.lib
L.SLAB200_Encoder_EncodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
L.SLAB200_Decoder_DecodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
nch, bits, rate = 2, 16, 44100
n = secs * rate
pcm = torch.from_numpy(synth.synth_long(nch, n, bits, rate, 0)).cuda()
cap = 43 + n * nch * 3 + (1 << 20)
stream = torch.zeros(cap, dtype=torch.uint8, device="cuda")
out = torch.empty_like(pcm)
enc = L.SLAEncoder_Create(C.byref(capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)))
dec = L.SLADecoder_Create(C.byref(capi.DecoderConfig(**capi.CLI_CAPACITY, enable_crc_check=1, verpose_flag=0)))
wf = capi.WaveFormat(nch, bits, rate, 0); ep = capi.preset_parameter(preset, nch)
assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0 and L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == 0
ip = (C.c_void_p * nch)(*[pcm[c].data_ptr() for c in range(nch)])
op = (C.c_void_p * nch)(*[out[c].data_ptr() for c in range(nch)])
size, got = C.c_uint32(0), C.c_uint32(0)
for _ in range(reps):
    assert L.SLAB200_Encoder_EncodeWholeDevice(enc, ip, n, stream.data_ptr(), cap, C.byref(size)) == 0
    assert L.SLAB200_Decoder_DecodeWholeDevice(dec, stream.data_ptr(), size.value, op, n, C.byref(got)) == 0
torch.cuda.synchronize()
print("ok", size.value, got.value, bool(torch.equal(out, pcm)))
