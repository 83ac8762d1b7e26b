"""Dev helper: milestones of the pipelined host-API calls (SLAB200_PIPE_TRACE=1) on C2, the box's host
(cores, memory) and its PCIe rates from pinned and pageable memory.
usage: python tools/e2e_trace.py [seconds]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sla_b200 import capi, synth
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
print("cores", os.cpu_count(), flush=True)
os.system("free -g | head -2; nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv")
# ---- raw copy rates ----
nb = 1 << 30
pin = torch.empty(nb, dtype=torch.uint8, pin_memory=True); pin.fill_(1)
pag = torch.empty(nb, dtype=torch.uint8); pag.fill_(1)
dev = torch.empty(nb, dtype=torch.uint8, device="cuda")
def rate(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return nb / best / 1e9
print("H2D pinned GB/s", round(rate(lambda: dev.copy_(pin, non_blocking=True)), 1))
print("H2D pageable GB/s", round(rate(lambda: dev.copy_(pag)), 1))
print("D2H pinned GB/s", round(rate(lambda: pin.copy_(dev, non_blocking=True)), 1))
print("D2H pageable GB/s", round(rate(lambda: pag.copy_(dev)), 1))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
pin2 = torch.empty(nb, dtype=torch.uint8, pin_memory=True); dev2 = torch.empty(nb, dtype=torch.uint8, device="cuda")
def both():
    with torch.cuda.stream(s1): dev.copy_(pin, non_blocking=True)
    with torch.cuda.stream(s2): pin2.copy_(dev2, non_blocking=True)
print("H2D+D2H concurrent, GB/s each way", round(rate(both), 1))
t0 = time.perf_counter(); torch.cuda.cudart().cudaHostRegister(pag.data_ptr(), nb, 0); t1 = time.perf_counter()
print("cudaHostRegister 1 GiB ms", round(1e3 * (t1 - t0), 1))
print("H2D registered GB/s", round(rate(lambda: dev.copy_(pag, non_blocking=True)), 1))
t0 = time.perf_counter(); torch.cuda.cudart().cudaHostUnregister(pag.data_ptr()); t1 = time.perf_counter()
print("cudaHostUnregister ms", round(1e3 * (t1 - t0), 1))
t0 = time.perf_counter(); pin[:] = pag; t1 = time.perf_counter()
print("host memcpy 1 GiB pageable->pinned, 1 thread GB/s", round(nb / (t1 - t0) / 1e9, 1))
del pin, pag, dev, pin2, dev2

lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so")); L = lib.lib
nch, bits, rate_hz = 2, 16, 44100
n = secs * rate_hz
h_pcm_t = torch.empty((nch, n), dtype=torch.int32, pin_memory=True)
t0 = time.perf_counter(); synth.synth_long(nch, n, bits, rate_hz, 0, out=h_pcm_t.numpy()); print("synth s", round(time.perf_counter() - t0, 1), flush=True)
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
    wf = capi.WaveFormat(nch, bits, rate_hz, 0)
    assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0 and L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == 0
    ip = (C.c_void_p * nch)(*[h_pcm_t[c].data_ptr() for c in range(nch)])
    op = (C.c_void_p * nch)(*[h_dec[c].data_ptr() for c in range(nch)])
    size, got = C.c_uint32(0), C.c_uint32(0)
    te, td = [], []
    for r in range(reps + 1):
        if r == reps: os.environ["SLAB200_PIPE_TRACE"] = "1"
        sys.stderr.flush()
        t0 = time.perf_counter(); rc = L.SLAEncoder_EncodeWhole(enc, ip, n, h_stream.data_ptr(), cap, C.byref(size)); t1 = time.perf_counter()
        assert rc == 0, rc
        rc = L.SLADecoder_DecodeWhole(dec, h_stream.data_ptr(), size.value, op, n, C.byref(got)); t2 = time.perf_counter()
        assert rc == 0, rc
        if r: te.append(1e3 * (t1 - t0)); td.append(1e3 * (t2 - t1))
    print({k[13:]: v for k, v in env.items()}, "enc", [round(x, 1) for x in te], "dec", [round(x, 1) for x in td], flush=True)
    L.SLAEncoder_Destroy(enc); L.SLADecoder_Destroy(dec)
run({}, reps=4)
for w in (4, 8):
    for ch in (8, 16):
        chunk = (n // ch // 12288 + 1) * 12288
        run({"SLAB200_PIPE_WORKERS": str(w), "SLAB200_PIPE_CHUNK_SAMPLES": str(chunk), "SLAB200_PIPE_DEC_CHUNKS": str(ch)}, reps=3)
