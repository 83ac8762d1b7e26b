"""Dev helper: repeat small pipelined encodes and report any run whose bytes differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from sla_b200 import capi, synth
from conftest import signal_set
lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for workers in ("4", "1"):
    os.environ["SLAB200_PIPE_WORKERS"] = workers
    for name, pcm, bits, rate in signal_set():
        pcm = np.ascontiguousarray(pcm)
        for preset in (2, 0):
            ep = capi.preset_parameter(preset, pcm.shape[0])
            os.environ.pop("SLAB200_PIPE_CHUNK_SAMPLES", None)
            rc, want = lib.encode_whole(pcm, bits, rate, ep)
            assert rc == 0
            bad = 0
            for chunk in ("1", "30000"):
                os.environ["SLAB200_PIPE_CHUNK_SAMPLES"] = chunk
                for r in range(reps):
                    rc, got = lib.encode_whole(pcm, bits, rate, ep)
                    if rc != 0 or got != want:
                        bad += 1
                        first = next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), -1)
                        print("MISMATCH", workers, name, preset, chunk, "rep", r, "rc", rc, len(got), len(want), "first diff at", first, flush=True)
            os.environ.pop("SLAB200_PIPE_CHUNK_SAMPLES", None)
            # single pass repeated as well
            for r in range(reps):
                rc, got = lib.encode_whole(pcm, bits, rate, ep)
                if rc != 0 or got != want:
                    bad += 1
                    print("MISMATCH single", workers, name, preset, "rep", r, rc, len(got), len(want), flush=True)
            print("workers", workers, name, "preset", preset, "bad", bad, flush=True)
