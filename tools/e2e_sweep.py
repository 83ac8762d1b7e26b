"""Dev helper: wall time of the host-API whole-file calls under different pipeline settings.
usage: python tools/e2e_sweep.py [seconds]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sla_b200 import capi, synth
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so")); L = lib.lib
L.SLAB200_Encoder_EncodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
L.SLAB200_Decoder_DecodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
nch, bits, rate = 2, 16, 44100
n = secs * rate
h_pcm_t = torch.empty((nch, n), dtype=torch.int32, pin_memory=True)
synth.synth_long(nch, n, bits, rate, 0, out=h_pcm_t.numpy())
h_raw = torch.empty(n * nch * 2, dtype=torch.uint8, pin_memory=True)
h_raw.numpy()[:] = np.frombuffer(capi.planar_to_pcm(h_pcm_t.numpy(), bits), dtype=np.uint8)
h_back = torch.empty(n * nch * 2, dtype=torch.uint8, pin_memory=True)
cap = 43 + n * nch * 3 + (1 << 20)
h_stream = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
h_dec = torch.empty((nch, n), dtype=torch.int32, pin_memory=True)
ep = capi.preset_parameter(2, nch)
def run(env, reps=3):
    for k in ("SLAB200_PIPE_WORKERS", "SLAB200_PIPE_CHUNK_SAMPLES", "SLAB200_PIPE_DEC_CHUNKS", "SLAB200_PIPE_TRACE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    enc = L.SLAEncoder_Create(C.byref(capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)))
    dec = L.SLADecoder_Create(C.byref(capi.DecoderConfig(**capi.CLI_CAPACITY, enable_crc_check=1, verpose_flag=0)))
    wf = capi.WaveFormat(nch, bits, rate, 0)
    assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0 and L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == 0
    ip = (C.c_void_p * nch)(*[h_pcm_t[c].data_ptr() for c in range(nch)])
    op = (C.c_void_p * nch)(*[h_dec[c].data_ptr() for c in range(nch)])
    size, got = C.c_uint32(0), C.c_uint32(0)
    te, td, tpe, tpd = [], [], [], []
    for r in range(reps + 1):
        t0 = time.perf_counter(); rc = L.SLAEncoder_EncodeWhole(enc, ip, n, h_stream.data_ptr(), cap, C.byref(size)); t1 = time.perf_counter()
        assert rc == 0, rc
        rc = L.SLADecoder_DecodeWhole(dec, h_stream.data_ptr(), size.value, op, n, C.byref(got)); t2 = time.perf_counter()
        assert rc == 0, rc
        rc = L.SLAB200_Encoder_EncodePCM(enc, h_raw.data_ptr(), n, h_stream.data_ptr(), cap, C.byref(size)); t3 = time.perf_counter()
        assert rc == 0, rc
        rc = L.SLAB200_Decoder_DecodePCM(dec, h_stream.data_ptr(), size.value, h_back.data_ptr(), n, C.byref(got)); t4 = time.perf_counter()
        assert rc == 0, rc
        if r: te.append(1e3 * (t1 - t0)); td.append(1e3 * (t2 - t1)); tpe.append(1e3 * (t3 - t2)); tpd.append(1e3 * (t4 - t3))
    ok = bool(torch.equal(h_dec, h_pcm_t)) and bool(torch.equal(h_back, h_raw))
    f = lambda v: round(min(v), 1)
    print({k[13:]: v for k, v in env.items()}, "enc", f(te), "dec", f(td), "| pcm enc", f(tpe), "pcm dec", f(tpd), "exact", ok, flush=True)
    L.SLAEncoder_Destroy(enc); L.SLADecoder_Destroy(dec)
for w in (3, 4, 5, 6):
    run({"SLAB200_PIPE_WORKERS": str(w), "SLAB200_PIPE_DEC_CHUNKS": str(w)}, reps=4)
run({"SLAB200_PIPE_WORKERS": "4", "SLAB200_PIPE_CHUNK_SAMPLES": str(n // 4), "SLAB200_PIPE_DEC_CHUNKS": "4"}, reps=4)
