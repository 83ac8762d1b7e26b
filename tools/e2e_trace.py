"""Dev helper: wall time and milestones (SLAB200_PIPE_TRACE=1) of the pipelined host-API calls on C2 for
a sweep of chunk / context counts, from page-locked and from pageable caller memory.
usage: python tools/e2e_trace.py [seconds] [quick]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sla_b200 import capi, synth
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
quick = len(sys.argv) > 2
print("cores", os.cpu_count(), flush=True)
lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so")); L = lib.lib
L.SLAB200_Encoder_EncodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
L.SLAB200_Decoder_DecodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
nch, bits, rate_hz = 2, 16, 44100
n = secs * rate_hz
cap = 43 + n * nch * 3 + (1 << 20)
def buffers(pinned):
    mk = (lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)) if pinned else (lambda shape, dt: torch.empty(shape, dtype=dt))
    return dict(pcm=mk((nch, n), torch.int32), stream=mk(cap, torch.uint8), dec=mk((nch, n), torch.int32),
                raw=mk(n * nch * 2, torch.uint8), back=mk(n * nch * 2, torch.uint8))
B = {True: buffers(True), False: buffers(False)}
t0 = time.perf_counter(); synth.synth_long(nch, n, bits, rate_hz, 0, out=B[True]["pcm"].numpy()); print("synth s", round(time.perf_counter() - t0, 1), flush=True)
B[False]["pcm"].copy_(B[True]["pcm"])
raw = np.frombuffer(capi.planar_to_pcm(B[True]["pcm"].numpy(), bits), dtype=np.uint8)
B[True]["raw"].numpy()[:] = raw; B[False]["raw"].numpy()[:] = raw
for b in B.values(): b["stream"].zero_(); b["dec"].zero_(); b["back"].zero_()      # touch every page
ep = capi.preset_parameter(2, nch)
KEYS = ("SLAB200_PIPE_WORKERS", "SLAB200_PIPE_CHUNKS", "SLAB200_PIPE_CHUNK_SAMPLES", "SLAB200_PIPE_DEC_CHUNKS", "SLAB200_PIPE_TRACE",
        "SLAB200_PIPE_ENC_WORKERS", "SLAB200_PIPE_DEC_WORKERS", "SLAB200_PIPE_TAPER", "SLAB200_BOUNCE_THREADS")
ref_stream = None
def run(env, pinned=True, reps=3, trace=False, pcm=False):
    global ref_stream
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update(env)
    b = B[pinned]
    enc = L.SLAEncoder_Create(C.byref(capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)))
    dec = L.SLADecoder_Create(C.byref(capi.DecoderConfig(**capi.CLI_CAPACITY, enable_crc_check=1, verpose_flag=0)))
    wf = capi.WaveFormat(nch, bits, rate_hz, 0)
    assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0 and L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == 0
    ip = (C.c_void_p * nch)(*[b["pcm"][c].data_ptr() for c in range(nch)])
    op = (C.c_void_p * nch)(*[b["dec"][c].data_ptr() for c in range(nch)])
    size, got = C.c_uint32(0), C.c_uint32(0)
    te, td = [], []
    for r in range(reps + 1):
        if r == reps and trace: os.environ["SLAB200_PIPE_TRACE"] = os.environ.get("TRACE_LEVEL", "1")
        sys.stderr.flush()
        t0 = time.perf_counter()
        rc = (L.SLAB200_Encoder_EncodePCM(enc, b["raw"].data_ptr(), n, b["stream"].data_ptr(), cap, C.byref(size)) if pcm else
              L.SLAEncoder_EncodeWhole(enc, ip, n, b["stream"].data_ptr(), cap, C.byref(size)))
        t1 = time.perf_counter()
        assert rc == 0, rc
        rc = (L.SLAB200_Decoder_DecodePCM(dec, b["stream"].data_ptr(), size.value, b["back"].data_ptr(), n, C.byref(got)) if pcm else
              L.SLADecoder_DecodeWhole(dec, b["stream"].data_ptr(), size.value, op, n, C.byref(got)))
        t2 = time.perf_counter()
        assert rc == 0, rc
        if r: te.append(1e3 * (t1 - t0)); td.append(1e3 * (t2 - t1))
    st = b["stream"][:size.value].clone()
    if ref_stream is None: ref_stream = st
    ok = bool(torch.equal(st, ref_stream)) and (bool(torch.equal(b["back"], B[True]["raw"])) if pcm else bool(torch.equal(b["dec"], B[True]["pcm"])))
    print({k[13:]: v for k, v in env.items()}, "pinned" if pinned else "pageable", "pcm" if pcm else "planar",
          "enc", [round(x, 1) for x in te], "dec", [round(x, 1) for x in td], "ok", ok, flush=True)
    L.SLAEncoder_Destroy(enc); L.SLADecoder_Destroy(dec)
mode = os.environ.get("E2E_MODE", "sweep")
def cfg(k, w, **extra):
    d = {"SLAB200_PIPE_CHUNKS": str(k), "SLAB200_PIPE_DEC_CHUNKS": str(k), "SLAB200_PIPE_ENC_WORKERS": str(w), "SLAB200_PIPE_DEC_WORKERS": str(w)}
    d.update(extra)
    return d
if mode == "prof":
    os.environ["TRACE_LEVEL"] = "2"
run({}, reps=3, trace=True)
if mode == "short":
    run(cfg(8, 8))
    run(cfg(12, 8))
    run({}, pinned=False, reps=3, trace=True)
    run({}, pcm=True, reps=3, trace=True)
    run({}, pcm=True, pinned=False, reps=3)
if mode == "sweep":
    for k, w in ((8, 8), (10, 8), (12, 8), (10, 10), (14, 10), (6, 6)):
        run(cfg(k, w))
    run(cfg(8, 8, SLAB200_PIPE_TAPER="0"))
    run({}, pinned=False, reps=3, trace=True)
    for t in (3, 8):
        run({"SLAB200_BOUNCE_THREADS": str(t)}, pinned=False, reps=3)
    run({}, pcm=True, reps=3, trace=True)
    run({}, pcm=True, pinned=False, reps=3)
    run(cfg(16, 8), pcm=True)
