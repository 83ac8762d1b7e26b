"""Full-size check of a GPU-encoded stream with the REFERENCE decoder (oracle/_ref/libsla_ref.so, the
unmodified reference compiled by oracle/Makefile): SLADecoder_DecodeWhole over the whole stream, output
compared with the regenerated synthetic input.  bench.py runs this in the background on one host core.

usage: python tools/ref_decode_check.py <stream file> <C2|C3|C4> <rank> <file index> [seconds]
prints one JSON line: {"ok": bool, "rc": int, "samples": n, "decode_s": t}
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sla_b200 import capi, workloads  # noqa: E402


def main():
    path, name, rank, index = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    seconds = int(sys.argv[5]) if len(sys.argv) > 5 else None
    ref = capi.SLALibrary(os.path.join(ROOT, "oracle", "_ref", "libsla_ref.so"))
    data = np.fromfile(path, dtype=np.uint8).tobytes()
    t0 = time.perf_counter()
    rc, pcm, h = ref.decode_whole(data)
    dt = time.perf_counter() - t0
    if name == "C4":
        want = workloads.c4_file(workloads.c4_base(rank, seconds), index)
    else:
        want = workloads.long_file(name, rank, seconds=seconds)
    ok = rc == 0 and pcm is not None and pcm.shape == want.shape and bool(np.array_equal(pcm, want))
    print(json.dumps({"ok": ok, "rc": int(rc), "samples": int(want.shape[1]), "channels": int(want.shape[0]),
                      "blocks": int(h.num_blocks), "decode_s": round(dt, 2),
                      "decoder": "reference SLADecoder_DecodeWhole (oracle/_ref/libsla_ref.so), whole stream, 1 core"}))
    try:
        os.unlink(path)
    except OSError:
        pass


if __name__ == "__main__":
    main()
