"""Corpus encode rate of SLAB200_Encoder_EncodeBatchPCM: `files` stereo 16-bit files of 3-30 s cut from a pool of
synthetic signals, PCM and streams in page-locked host memory; the call is repeated so that the second and third
figures are with warm arenas.  usage: python tools/batch_rate.py [files] [preset]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sla_b200 import capi, synth  # noqa: E402


def main():
    nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    preset = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    nch, bits, rate = 2, 16, 44100
    lib = capi.SLALibrary(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sla_b200", "lib", "libsla_b200.so"))
    L = lib.lib
    rng = np.random.default_rng(5)
    pool_frames = 30 * rate
    pool = [torch.frombuffer(bytearray(capi.planar_to_pcm(np.ascontiguousarray(synth.synth_pcm(nch, pool_frames, bits, rate, 900 + i)), bits)),
                             dtype=torch.uint8).pin_memory() for i in range(8)]
    fb = nch * bits // 8
    frames = rng.integers(3 * rate, pool_frames, nfiles)
    first = np.array([rng.integers(0, pool_frames - f + 1) for f in frames])
    which = rng.integers(0, len(pool), nfiles)
    caps = [43 + int(f) * fb + 65536 for f in frames]
    out = torch.empty(int(sum(caps)), dtype=torch.uint8).pin_memory()
    items = (capi.EncodeItem * nfiles)()
    at = 0
    for i in range(nfiles):
        items[i].pcm = pool[int(which[i])].data_ptr() + int(first[i]) * fb
        items[i].num_samples = int(frames[i])
        items[i].data = out.data_ptr() + at
        items[i].data_size = caps[i]
        at += caps[i]
    L.SLAB200_Encoder_EncodeBatchPCM.argtypes = [C.c_void_p, C.POINTER(capi.EncodeItem), C.c_uint32]
    cfg = capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)
    enc = L.SLAEncoder_Create(C.byref(cfg))
    wf = capi.WaveFormat(nch, bits, rate, 0)
    ep = capi.preset_parameter(preset, nch)
    assert enc and L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0 and L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == 0
    chs = int(frames.sum()) * nch
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = L.SLAB200_Encoder_EncodeBatchPCM(enc, items, nfiles)
        dt = time.perf_counter() - t0
        bad = sum(1 for i in range(nfiles) if items[i].result != 0)
        print(f"pass {rep}: rc {rc}, {bad} failed items, {dt * 1e3:.1f} ms, {chs / dt / 1e6:.0f} M channel-samples/s "
              f"({nfiles} files, {chs / 1e6:.0f} M channel-samples)", flush=True)


if __name__ == "__main__":
    main()
