"""Dev helper: per-block intermediates of the GPU encoder against the reference white box with the fallback
paths switched off one by one (SLAB200_DEBUG_OFF, SLAB200_PACK_FAST).
usage: python tools/gpu_bisect.py"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from sla_b200 import capi, synth, parity
from oracle import binding as ob
import test_encode as TE
lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so"))
refwb = ob.RefWhitebox()
cases = [("C2 4 s", synth.synth_pcm(2, 4 * 44100, 16, 44100, 1000), 16, 44100, 2),
         ("C3 2 s", synth.synth_pcm(2, 2 * 96000, 24, 96000, 1001), 24, 96000, 4),
         ("p0", synth.synth_pcm(2, 60000, 16, 44100, 0, clear_low_bits=4), 16, 44100, 0)]
for env in ({}, {"SLAB200_DEBUG_OFF": "3", "SLAB200_PACK_FAST": "0"}, {"SLAB200_DEBUG_OFF": "1"}, {"SLAB200_DEBUG_OFF": "2"}, {"SLAB200_PACK_FAST": "0"}):
    for k in ("SLAB200_DEBUG_OFF", "SLAB200_PACK_FAST"): os.environ.pop(k, None)
    os.environ.update(env)
    for name, pcm, bits, rate, preset in cases:
        ep = capi.preset_parameter(preset, 2)
        rc, want, wblocks, wres = refwb.encode_whole(pcm, bits, rate, ep, capi.CLI_CAPACITY, want_residual=True)
        data, recs, res = TE.encode_with_export(lib, pcm, bits, rate, ep)
        P, T = ep.parcor_order, ep.longterm_order
        bad = {"type": 0, "parcor": 0, "parcor_zero": 0, "code": 0, "rshift": 0, "pitch": 0, "taps": 0, "rice": 0, "size": 0}
        worst = 0.0
        for m, w in zip(recs, wblocks):
            if m.block_type != w.block_type: bad["type"] += 1
            if m.block_size != w.block_size: bad["size"] += 1
            if m.block_type != 0 or w.block_type != 0: continue
            for c in range(2):
                pa, pb = np.array(m.parcor[c][1:P + 1]), np.array(w.parcor[c][1:P + 1])
                rel = float(np.max(np.abs(pa - pb) / np.maximum(np.abs(pb), 1e-300)))
                worst = max(worst, rel)
                if rel > 1e-9: bad["parcor"] += 1
                if not np.any(pa): bad["parcor_zero"] += 1
                if list(m.parcor_code[c][1:P + 1]) != list(w.parcor_code[c][1:P + 1]): bad["code"] += 1
                if m.rshift[c] != w.rshift[c]: bad["rshift"] += 1
                if m.pitch[c] != w.pitch[c]: bad["pitch"] += 1
                if list(m.lt_q31[c][:T]) != list(w.lt_q31[c][:T]): bad["taps"] += 1
                if m.rice_init[c] != w.rice_init[c]: bad["rice"] += 1
        print(env, name, "blocks", len(recs), len(wblocks), "identical", data == want, "worst parcor rel", f"{worst:.2e}", {k: v for k, v in bad.items() if v}, flush=True)
