"""Host-side sharding of one file across ranks (one process per GPU).

Blocks are independent (all codec state resets at a block start, reference src/SLAEncoder.c:594-675),
so a file shards by contiguous sample ranges that start on the encoder's segment grid.  The only
cross-rank data are metadata:

  1. OR mask of the input words  -> offset_lshift   (all-gather of one uint32 per rank; the reference
     computes it over the whole file, src/SLAEncoder.c:425-455)
  2. per-rank {num_blocks, bytes, max_block_size, max_bit_per_second} -> byte offsets + file header

No collective touches samples or bitstream bytes on the math path; each rank writes its span at its
offset.  `dist` is torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Range boundaries are multiples of max_num_block_samples.  That reproduces the single-encoder stream
exactly as long as the leading-silence rule (src/SLAEncoder.c:393-408) has not re-based the segment
grid before the boundary; when it has, the stitched stream is still a valid .sla stream that decodes
bit-exactly, but its block boundaries after that point differ from the single-encoder ones.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class RangeResult(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_blocks", "total_bytes", "max_block_size",
                                           "max_bit_per_second", "input_or_mask")]


def plan_ranges(num_samples: int, max_block: int, world: int):
    """Contiguous [start, stop) per rank, boundaries on the max_block grid, sizes as even as possible."""
    nseg = (num_samples + max_block - 1) // max_block
    out, seg = [], 0
    for r in range(world):
        take = nseg // world + (1 if r < nseg % world else 0)
        start, stop = seg * max_block, min((seg + take) * max_block, num_samples)
        out.append((min(start, num_samples), stop))
        seg += take
    return out


def lshift_from_mask(mask: int, bits: int) -> int:
    if mask == 0:
        return 0
    ntz = (mask & -mask).bit_length() - 1
    return bits - (32 - ntz)


def encode_sharded(lib: capi.SLALibrary, dist, rank: int, world: int, pcm: np.ndarray, bits: int, rate: int,
                   param: capi.EncodeParameter, capacity: dict | None = None):
    """Every rank passes the same `pcm` view of the file (or at least its own range); returns the full
    stitched stream on every rank (gathered through `dist`)."""
    import torch
    L = lib.lib
    L.SLAB200_Encoder_InputOrMask.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.SLAB200_Encoder_EncodeRange.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                              C.c_uint32, C.POINTER(RangeResult)]
    nch, n = pcm.shape
    start, stop = plan_ranges(n, param.max_num_block_samples, world)[rank]
    mine = np.ascontiguousarray(pcm[:, start:stop])
    cfg = capi.EncoderConfig(**(capacity or capi.CLI_CAPACITY), verpose_flag=0)
    enc = L.SLAEncoder_Create(C.byref(cfg))
    if not enc:
        raise RuntimeError("SLAEncoder_Create failed")
    try:
        wf = capi.WaveFormat(nch, bits, rate, 0)
        assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == capi.OK
        assert L.SLAEncoder_SetEncodeParameter(enc, C.byref(param)) == capi.OK
        # (1) offset_lshift from the OR of every rank's mask
        mask = C.c_uint32(0)
        if mine.shape[1]:
            assert L.SLAB200_Encoder_InputOrMask(enc, capi._planar_pointers(mine), mine.shape[1], C.byref(mask)) == capi.OK
        masks = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(masks, torch.tensor([mask.value], dtype=torch.int64))
        total_mask = 0
        for m in masks:
            total_mask |= int(m.item())
        lshift = lshift_from_mask(total_mask, bits)
        # (2) this rank's blocks
        res = RangeResult()
        cap = 2 * mine.size * max(bits // 8, 1) + 65536
        out = np.zeros(cap, dtype=np.uint8)
        if mine.shape[1]:
            rc = L.SLAB200_Encoder_EncodeRange(enc, capi._planar_pointers(mine), mine.shape[1], lshift,
                                               out.ctypes.data, cap, C.byref(res))
            assert rc == capi.OK, rc
        # (3) metadata all-gather -> offsets, header
        meta = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(meta, torch.tensor([res.num_blocks, res.total_bytes, res.max_block_size,
                                            res.max_bit_per_second], dtype=torch.int64))
        sizes = [int(m[1]) for m in meta]
        header = capi.HeaderInfo()
        header.wave_format = capi.WaveFormat(nch, bits, rate, lshift)
        header.encode_param = param
        header.num_samples = n
        header.num_blocks = sum(int(m[0]) for m in meta)
        header.max_block_size = max(int(m[2]) for m in meta)
        header.max_bit_per_second = max(int(m[3]) for m in meta)
        head = np.zeros(capi.HEADER_SIZE, dtype=np.uint8)
        assert L.SLAEncoder_EncodeHeader(C.byref(header), head.ctypes.data, capi.HEADER_SIZE) == capi.OK
        # (4) spans: gathered here so the test can look at the whole stream (a real deployment writes
        #     each span at its offset instead)
        biggest = max(sizes) if sizes else 0
        spans = [torch.zeros(biggest, dtype=torch.uint8) for _ in range(world)]
        padded = torch.zeros(biggest, dtype=torch.uint8)
        padded[:res.total_bytes] = torch.from_numpy(out[:res.total_bytes].copy())
        dist.all_gather(spans, padded)
        stream = head.tobytes() + b"".join(spans[r][:sizes[r]].numpy().tobytes() for r in range(world))
        offsets = [capi.HEADER_SIZE + sum(sizes[:r]) for r in range(world)]
        return stream, offsets, lshift
    finally:
        L.SLAEncoder_Destroy(enc)
