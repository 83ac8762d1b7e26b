"""Dev helper: device-resident encode of C2 (or C3) as one pass and as K chunks on K contexts
(SLAB200_PIPE_DEVICE=1), wall clock per call; streams compared byte for byte.
usage: python tools/dev_overlap.py [C2|C3] [seconds]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
from sla_b200 import capi, workloads
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
c = workloads.CONFIGS[name]
secs = int(sys.argv[2]) if len(sys.argv) > 2 else c["seconds"]
nch, bits, rate, preset = c["channels"], c["bits"], c["rate"], c["preset"]
n = secs * rate
lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so")); L = lib.lib
L.SLAB200_Encoder_EncodeWholeDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
L.SLAB200_Encoder_LastTiming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
pcm = torch.from_numpy(workloads.long_file(name, 0, seconds=secs)).cuda()
cap = 43 + int(n * nch * max(bits // 8, 1) * 1.25) + (1 << 20)
out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ep = capi.preset_parameter(preset, nch)
ref = None
def run(env, reps=5):
    global ref
    for k in ("SLAB200_PIPE_DEVICE", "SLAB200_PIPE_CHUNKS", "SLAB200_PIPE_ENC_WORKERS", "SLAB200_PIPE_TAPER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    enc = L.SLAEncoder_Create(C.byref(capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)))
    wf = capi.WaveFormat(nch, bits, rate, 0)
    assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0 and L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == 0
    ip = (C.c_void_p * nch)(*[pcm[ch].data_ptr() for ch in range(nch)])
    size = C.c_uint32(0)
    ts = []
    for r in range(reps + 2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = L.SLAB200_Encoder_EncodeWholeDevice(enc, ip, n, out.data_ptr(), cap, C.byref(size))
        t1 = time.perf_counter()
        assert rc == 0, rc
        if r >= 2: ts.append(1e3 * (t1 - t0))
    st = out[:size.value].clone()
    if ref is None: ref = st
    print({k[13:]: v for k, v in env.items()}, "wall ms", [round(x, 2) for x in ts], "same", bool(torch.equal(st, ref)), flush=True)
    L.SLAEncoder_Destroy(enc)
run({})
for k in (2, 3, 4, 6, 8):
    run({"SLAB200_PIPE_DEVICE": "1", "SLAB200_PIPE_CHUNKS": str(k), "SLAB200_PIPE_ENC_WORKERS": str(k), "SLAB200_PIPE_TAPER": "0"})
