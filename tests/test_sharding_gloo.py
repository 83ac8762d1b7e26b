"""N>1 host logic on CPU: two gloo ranks shard one file by sample ranges, exchange only metadata
(OR masks, block counts, sizes), and the stitched stream must equal the single-encoder stream and
decode bit-exactly.  Kernels run through the host simulator here; the same code drives NCCL ranks."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import HOSTSIM_SO, ROOT
from sla_b200 import capi, shard, synth


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = capi.SLALibrary(HOSTSIM_SO)
        pcm = synth.synth_pcm(2, 6 * 12288 + 5000, 16, 44100, 21, specials=False, clear_low_bits=3)
        ep = capi.preset_parameter(2, 2)
        stream, offsets, lshift = shard.encode_sharded(lib, dist, rank, world, pcm, 16, 44100, ep)
        rc, single = lib.encode_whole(pcm, 16, 44100, ep)
        rc2, dec, h = lib.decode_whole(stream)
        q.put((rank, rc == 0 and stream == single, rc2 == 0 and bool(np.array_equal(dec, pcm)), lshift, offsets))
    finally:
        dist.destroy_process_group()


def test_plan_ranges():
    assert shard.plan_ranges(100000, 12288, 2) == [(0, 61440), (61440, 100000)]
    assert shard.plan_ranges(5000, 12288, 4) == [(0, 5000), (5000, 5000), (5000, 5000), (5000, 5000)]
    r = shard.plan_ranges(158760000, 12288, 8)
    assert r[0][0] == 0 and r[-1][1] == 158760000 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert all(a % 12288 == 0 for a, _ in r)
    assert shard.lshift_from_mask(0xFFF00000, 16) == 4 and shard.lshift_from_mask(0, 16) == 0


def test_two_rank_sharded_encode_matches_single(hostsim):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, exact, lshift, offsets in results:
        assert same, f"rank {rank}: stitched stream differs from the single-encoder stream"
        assert exact, f"rank {rank}: stitched stream does not decode bit-exactly"
        assert lshift == 3 and offsets[0] == 43
