"""N>1 host logic on CPU: gloo ranks shard ONE file by sample ranges, exchange only metadata (trailing-zero
counts, the segment chain hand-off, block counts and sizes), write their spans into a shared file, and the
stitched stream must equal the single-encoder stream byte for byte - also when leading-silence blocks
re-base the segment grid across shard boundaries (SLAEncoder.c:393-408) - and decode bit-exactly.
Kernels run through the host simulator here; the same code drives NCCL ranks (bench.py --configs strong)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import HOSTSIM_SO, ROOT, multi_silence
from sla_b200 import capi, shard, synth


def _signals():
    # (pcm, bits, preset): cleared low bits (offset_lshift 3), and silences that end off the block grid and
    # straddle the shard boundaries of 2 and 3 ranks
    return [(synth.synth_pcm(2, 6 * 12288 + 5000, 16, 44100, 21, specials=False, clear_low_bits=3), 16, 2),
            (multi_silence(), 16, 2),
            (multi_silence(70000, seed=5), 16, 0)]


def _worker(rank, world, port, path, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = capi.SLALibrary(HOSTSIM_SO)
        out = []
        for i, (pcm, bits, preset) in enumerate(_signals()):
            ep = capi.preset_parameter(preset, pcm.shape[0])
            total, offset, lshift, res = shard.encode_sharded(lib, dist, rank, world, pcm, bits, 44100, ep, f"{path}.{i}")
            same = exact = None
            if rank == 0:
                with open(f"{path}.{i}", "rb") as f:
                    stream = f.read()[:total]
                rc, single = lib.encode_whole(pcm, bits, 44100, ep)
                rc2, dec, h = lib.decode_whole(stream)
                same = rc == 0 and stream == single
                exact = rc2 == 0 and bool(np.array_equal(dec, pcm))
            dist.barrier()
            out.append((same, exact, lshift, offset, int(res.num_blocks)))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_plan_ranges():
    assert shard.plan_ranges(100000, 12288, 2) == [(0, 61440), (61440, 100000)]
    assert shard.plan_ranges(5000, 12288, 4) == [(0, 5000), (5000, 5000), (5000, 5000), (5000, 5000)]
    r = shard.plan_ranges(158760000, 12288, 8)
    assert r[0][0] == 0 and r[-1][1] == 158760000 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert all(a % 12288 == 0 for a, _ in r)
    assert shard.upload_range(100000, 12288, 0, 61440) == (0, 73728)
    assert shard.upload_range(100000, 12288, 61440, 100000) == (61440, 100000)
    assert shard.upload_range(100000, 12288, 73728, 98304) == (73728, 100000)
    assert shard.lshift_from_ntz(shard.ntz_of_mask(0xFFF00000), 16) == 4
    assert shard.lshift_from_ntz(shard.ntz_of_mask(0), 16) == 0


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_encode_matches_single(world, hostsim):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "stitched.sla")
        procs = [ctx.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
        for p in procs:
            p.start()
        results = dict(q.get(timeout=600) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    for i, (same, exact, lshift, offset, nblocks) in enumerate(results[0]):
        assert same, f"signal {i}: stitched stream differs from the single-encoder stream"
        assert exact, f"signal {i}: stitched stream does not decode bit-exactly"
        assert offset == 43
    assert results[0][0][2] == 3                     # cleared low bits agreed across ranks
    # every rank contributed blocks on the first signal
    assert all(results[r][0][4] > 0 for r in range(world))
