"""One file encoded by several ranks (one process per GPU), chain-exact: the stitched stream is
byte-identical to the single-encoder stream (reference src/SLAEncoder.c:846-900 segment loop,
:393-408 leading-silence re-basing, :425-455 whole-file offset_lshift).

Blocks are independent (all codec state resets at a block start), so the file shards by contiguous
sample ranges.  What crosses ranks is metadata only:

  1. all_reduce(MIN) of each rank's trailing-zero count of its OR mask  -> offset_lshift
     (NCCL has no bitwise OR; ntz(OR of all) = min of the per-rank ntz)
  2. the segment chain: rank r receives, from rank r - 1, the sample where its chain starts (one integer,
     send/recv), computes its own chain - SLAB200_Encoder_EncodeShard calls back as soon as it is known,
     long before the blocks are encoded - and sends the end of it to rank r + 1.  Because every rank's
     chain starts where the previous one really stopped, leading-silence blocks that re-base the segment
     grid are reproduced exactly.
  3. all_gather of {num_blocks, bytes, max_block_size, max_bit_per_second} -> byte offsets + file header

Payload bytes never go through a collective: every rank copies its span from its device straight to its
offset in the destination (a shared host mapping here).  `dist` is torch.distributed (NCCL on GPUs, gloo
in the CPU tests, where the "device" of the host-simulator build is host memory).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import mmap
import os
import time

import numpy as np

from . import capi


class ShardResult(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_blocks", "total_bytes", "max_block_size",
                                           "max_bit_per_second", "next_start")]


CHAIN_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32)


def bind(L):
    u32p = C.POINTER(C.c_uint32)
    L.SLAB200_Encoder_InputOrMask.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, u32p]
    L.SLAB200_Encoder_InputOrMaskDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, u32p]
    L.SLAB200_Encoder_EncodeShard.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_uint32, C.c_void_p, C.c_int, C.c_uint32, CHAIN_CB, C.c_void_p,
                                              C.POINTER(ShardResult)]
    L.SLAB200_Encoder_Download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]


def plan_ranges(num_samples: int, max_block: int, world: int):
    """Nominal [begin, end) per rank, boundaries on the max_block grid, sizes as even as possible."""
    nseg = (num_samples + max_block - 1) // max_block
    out, seg = [], 0
    for r in range(world):
        take = nseg // world + (1 if r < nseg % world else 0)
        start, stop = seg * max_block, min((seg + take) * max_block, num_samples)
        out.append((min(start, num_samples), stop))
        seg += take
    return out


def upload_range(num_samples: int, max_block: int, begin: int, end: int):
    """samples a rank needs: its nominal range plus one block (its last segment may run past `end`)"""
    hi = end + max_block if num_samples - end > max_block else num_samples
    return begin, hi


def lshift_from_ntz(ntz: int, bits: int) -> int:
    return 0 if ntz >= 32 else bits - (32 - ntz)


def ntz_of_mask(mask: int) -> int:
    return 32 if mask == 0 else (mask & -mask).bit_length() - 1


class SharedStream:
    """The destination of the stitched stream: one host mapping every rank of the node writes its span
    into (a file under /dev/shm)."""

    def __init__(self, path: str, size: int, create: bool):
        self.path, self.size = path, size
        if create:
            with open(path, "wb") as f:
                f.truncate(size)
        self.fd = os.open(path, os.O_RDWR)
        self.map = mmap.mmap(self.fd, size)
        self.view = np.frombuffer(self.map, dtype=np.uint8)

    @property
    def address(self) -> int:
        return self.view.ctypes.data

    def close(self, unlink: bool = False):
        self.view = None
        try:
            self.map.close()
        except BufferError:
            pass
        os.close(self.fd)
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


class ShardEncoder:
    """A rank's encoder for one shard of a file; buffers persist over repeated encodes (benchmark steps)."""

    def __init__(self, lib: capi.SLALibrary, dist, rank: int, world: int, nch: int, bits: int, rate: int,
                 param: capi.EncodeParameter, num_samples: int, device=None, capacity: dict | None = None):
        self.lib, self.L, self.dist, self.rank, self.world = lib, lib.lib, dist, rank, world
        self.nch, self.bits, self.rate, self.param, self.n = nch, bits, rate, param, num_samples
        self.device = device                          # torch device of the collectives' tensors (None = CPU / gloo)
        bind(self.L)
        cfg = capi.EncoderConfig(**(capacity or capi.CLI_CAPACITY), verpose_flag=0)
        self.enc = self.L.SLAEncoder_Create(C.byref(cfg))
        if not self.enc:
            raise RuntimeError("SLAEncoder_Create failed")
        wf = capi.WaveFormat(nch, bits, rate, 0)
        assert self.L.SLAEncoder_SetWaveFormat(self.enc, C.byref(wf)) == capi.OK
        assert self.L.SLAEncoder_SetEncodeParameter(self.enc, C.byref(param)) == capi.OK
        maxblk = param.max_num_block_samples
        self.begin, self.end = plan_ranges(num_samples, maxblk, world)[rank]
        self.up_lo, self.up_hi = upload_range(num_samples, maxblk, self.begin, self.end)
        self.cap = 2 * nch * (self.up_hi - self.up_lo) * max(bits // 8, 1) + 65536

    def close(self):
        self.L.SLAEncoder_Destroy(self.enc)

    def _tensor(self, values):
        import torch
        return torch.tensor(values, dtype=torch.int64, device=self.device)

    def encode(self, plane_ptrs, planes_on_device: bool, out_ptr: int, out_on_device: bool):
        """plane_ptrs: ctypes array of nch pointers to sample `up_lo` of each plane.  Returns
        (ShardResult, offset_lshift, meta of all ranks [world][4], chain milliseconds)."""
        import torch
        dist, L = self.dist, self.L
        # (1) offset_lshift: min over ranks of ntz(OR mask) of the nominal ranges
        mask = C.c_uint32(0)
        nominal = self.end - self.begin
        if nominal > 0:
            fn = L.SLAB200_Encoder_InputOrMaskDevice if planes_on_device else L.SLAB200_Encoder_InputOrMask
            assert fn(self.enc, plane_ptrs, nominal, C.byref(mask)) == capi.OK
        t = self._tensor([ntz_of_mask(mask.value)])
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        lshift = lshift_from_ntz(int(t.item()), self.bits)
        # (2) chain start from the previous rank; my chain end to the next one from inside the call
        t0 = time.perf_counter()
        start_abs = self.begin
        if self.rank > 0:
            s = self._tensor([0])
            dist.recv(s, src=self.rank - 1)
            start_abs = int(s.item())
        chain_ms = [1e3 * (time.perf_counter() - t0)]
        sent = [False]

        def on_chain(_user, next_rel):
            if self.rank + 1 < self.world and not sent[0]:
                dist.send(self._tensor([self.up_lo + int(next_rel)]), dst=self.rank + 1)
                sent[0] = True
        cb = CHAIN_CB(on_chain)
        res = ShardResult()
        last = self.rank + 1 == self.world or self.end >= self.n
        soft_end = 0 if last else self.end - self.up_lo
        nsmp = self.up_hi - self.up_lo
        if nsmp > 0:
            rc = L.SLAB200_Encoder_EncodeShard(self.enc, plane_ptrs, 1 if planes_on_device else 0, nsmp,
                                               max(start_abs, self.up_lo) - self.up_lo, soft_end, lshift, out_ptr,
                                               1 if out_on_device else 0, min(self.cap, 0xFFFFFFFF), cb, None,
                                               C.byref(res))
            if rc != capi.OK:
                raise RuntimeError(f"SLAB200_Encoder_EncodeShard rc={rc}")
        if not sent[0]:
            on_chain(None, max(start_abs, self.up_lo) - self.up_lo)           # empty shard: pass the start on
        # (3) metadata of every rank
        mine = self._tensor([res.num_blocks, res.total_bytes, res.max_block_size, res.max_bit_per_second])
        if self.world > 1:
            metas = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(metas, mine)
            meta = [[int(x) for x in m.tolist()] for m in metas]
        else:
            meta = [[int(x) for x in mine.tolist()]]
        return res, lshift, meta, chain_ms[0]

    def header_bytes(self, lshift: int, meta) -> bytes:
        header = capi.HeaderInfo()
        header.wave_format = capi.WaveFormat(self.nch, self.bits, self.rate, lshift)
        header.encode_param = self.param
        header.num_samples = self.n
        header.num_blocks = sum(m[0] for m in meta)
        header.max_block_size = max(m[2] for m in meta)
        header.max_bit_per_second = max(m[3] for m in meta)
        head = np.zeros(capi.HEADER_SIZE, dtype=np.uint8)
        assert self.L.SLAEncoder_EncodeHeader(C.byref(header), head.ctypes.data, capi.HEADER_SIZE) == capi.OK
        return head.tobytes()

    def place(self, shared: SharedStream, out_ptr: int, out_on_device: bool, res: ShardResult, lshift: int, meta):
        """my span to its offset in the shared destination (device -> host directly); rank 0 adds the header"""
        offset = capi.HEADER_SIZE + sum(m[1] for m in meta[:self.rank])
        if res.total_bytes:
            if out_on_device:
                rc = self.L.SLAB200_Encoder_Download(self.enc, shared.address + offset, out_ptr, res.total_bytes)
                if rc != capi.OK:
                    raise RuntimeError(f"SLAB200_Encoder_Download rc={rc}")
            else:
                C.memmove(shared.address + offset, out_ptr, res.total_bytes)
        if self.rank == 0:
            shared.view[:capi.HEADER_SIZE] = np.frombuffer(self.header_bytes(lshift, meta), dtype=np.uint8)
        return offset


def encode_sharded(lib: capi.SLALibrary, dist, rank: int, world: int, pcm: np.ndarray, bits: int, rate: int,
                   param: capi.EncodeParameter, shared_path: str, capacity: dict | None = None):
    """Host planes in, stitched stream in the shared file `shared_path` (every rank passes the same path).
    Returns (stream size, byte offset of this rank's span, offset_lshift, ShardResult)."""
    nch, n = pcm.shape
    se = ShardEncoder(lib, dist, rank, world, nch, bits, rate, param, n, capacity=capacity)
    try:
        mine = np.ascontiguousarray(pcm[:, se.up_lo:se.up_hi]) if se.up_hi > se.up_lo else np.zeros((nch, 1), np.int32)
        out = np.zeros(max(se.cap, 1), dtype=np.uint8)
        res, lshift, meta, _ = se.encode(capi._planar_pointers(mine), False, out.ctypes.data, False)
        total = capi.HEADER_SIZE + sum(m[1] for m in meta)
        if rank == 0:
            SharedStream(shared_path, max(total, 1), create=True).close()
        if world > 1:
            dist.barrier()
        shared = SharedStream(shared_path, max(total, 1), create=False)
        offset = se.place(shared, out.ctypes.data, False, res, lshift, meta)
        shared.map.flush()
        if world > 1:
            dist.barrier()
        shared.close()
        return total, offset, lshift, res
    finally:
        se.close()


# ----------------------------------------------------------------------------- bench leg (bench.py --configs strong)
def bench_strong(a, D, L, lib, state, steps, log):
    """C2 and C3 as ONE file over all GPUs of the node: every rank holds only its range (+ one block),
    chain hand-off by send/recv, metadata collectives, spans written device -> shared host mapping.
    Checked: the stitched stream equals, byte for byte, the stream one GPU writes for the whole file."""
    import torch
    import torch.distributed as dist
    from . import workloads
    out = {}
    for name in ("C2", "C3"):
        c = workloads.CONFIGS[name]
        nch, bits, rate, preset = c["channels"], c["bits"], c["rate"], c["preset"]
        seconds = a.seconds or c["seconds"]
        n = seconds * rate
        param = capi.preset_parameter(preset, nch)
        se = ShardEncoder(lib, dist, D.rank, D.world, nch, bits, rate, param, n, device=D.dev)
        # every rank synthesises the same file (deterministic) and keeps its own range only
        t0 = time.perf_counter()
        full = workloads.long_file(name, 0, seconds=seconds)
        nsmp = se.up_hi - se.up_lo
        h_mine = torch.empty((nch, max(nsmp, 1)), dtype=torch.int32, pin_memory=True)
        h_mine.numpy()[:, :nsmp] = full[:, se.up_lo:se.up_hi]
        log(f"[strong {name}] rank {D.rank}: range [{se.begin}, {se.end}) of {n}, synthesised in {time.perf_counter() - t0:.1f} s")
        d_mine = h_mine.to(D.dev)
        d_out = torch.zeros(se.cap, dtype=torch.uint8, device=D.dev)
        d_ptrs = (C.c_void_p * nch)(*[d_mine[ch].data_ptr() for ch in range(nch)])
        torch.cuda.synchronize()          # the library's streams are not ordered against torch's
        # single-GPU stream of the whole file on rank 0: the expected bytes
        want_md5, want_size, single_ms = None, None, None
        if D.rank == 0:
            d_full = torch.from_numpy(full).to(D.dev)
            cap = 43 + int(n * nch * max(bits // 8, 1) * 1.25) + (1 << 20)
            d_single = torch.zeros(cap, dtype=torch.uint8, device=D.dev)
            enc_cfg = capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)
            enc = L.SLAEncoder_Create(C.byref(enc_cfg))
            wf = capi.WaveFormat(nch, bits, rate, 0)
            assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0 and L.SLAEncoder_SetEncodeParameter(enc, C.byref(param)) == 0
            size = C.c_uint32(0)
            fp = (C.c_void_p * nch)(*[d_full[ch].data_ptr() for ch in range(nch)])
            torch.cuda.synchronize()
            ms3, nl = (C.c_float * 3)(), C.c_uint32(0)
            for _ in range(2):
                assert L.SLAB200_Encoder_EncodeWholeDevice(enc, fp, n, d_single.data_ptr(), cap, C.byref(size)) == 0
            L.SLAB200_Encoder_LastTiming(enc, ms3, C.byref(nl))
            single_ms = ms3[0] + ms3[1] + ms3[2]
            want = d_single[:size.value].cpu().numpy()
            want_md5, want_size = hashlib.md5(want.tobytes()).hexdigest(), int(size.value)
            L.SLAEncoder_Destroy(enc)
            del d_full, d_single, want
            torch.cuda.empty_cache()
        del full
        # destination: one shared host mapping, page-locked in every rank
        t = torch.tensor([want_size or 0], dtype=torch.int64, device=D.dev)
        dist.broadcast(t, src=0)
        total_expected = int(t.item())
        path = f"/dev/shm/sla_b200_strong_{name}_{os.environ.get('MASTER_PORT', '0')}.sla"
        if D.rank == 0:
            SharedStream(path, total_expected + (1 << 20), create=True).close()
        D.barrier()
        shared = SharedStream(path, total_expected + (1 << 20), create=False)
        reg = torch.cuda.cudart().cudaHostRegister(shared.address, shared.size, 0)

        def step(device_resident):
            if device_resident:
                res, lshift, meta, chain_ms = se.encode(d_ptrs, True, d_out.data_ptr(), True)
                return res, lshift, meta, chain_ms
            d_mine.copy_(h_mine, non_blocking=True)                         # my range only: host -> device
            torch.cuda.current_stream().synchronize()
            res, lshift, meta, chain_ms = se.encode(d_ptrs, True, d_out.data_ptr(), True)
            se.place(shared, d_out.data_ptr(), True, res, lshift, meta)
            return res, lshift, meta, chain_ms
        res, lshift, meta, _ = step(False)
        D.barrier()
        total = capi.HEADER_SIZE + sum(m[1] for m in meta)
        same = None
        if D.rank == 0:
            got_md5 = hashlib.md5(shared.view[:total].tobytes()).hexdigest()
            same = (total == want_size and got_md5 == want_md5)
        # device-resident (ranges in HBM, spans left in HBM) and end to end (host range -> shared host stream)
        D.barrier()
        t0 = time.perf_counter()
        chain = 0.0
        for _ in range(steps):
            _, _, _, cm = step(True)
            chain += cm
        D.barrier()
        dev_ms = D.max(1e3 * (time.perf_counter() - t0) / steps)
        chain_wait_ms = D.max(chain / steps)
        e2e_ms = D.timed(lambda: step(False), steps)
        chsamp = n * nch
        rec = {
            "config": {"workload": f"{name} as ONE file over {D.world} GPUs: rank r holds the r-th range of whole "
                                   f"max_num_block_samples segments (+ one block)", "channel_samples": chsamp},
            "scaling": "strong", "unit": "M channel-samples/s", "steps": steps,
            "value": chsamp / (dev_ms * 1e-3) / 1e6, "ms_per_step": dev_ms,
            "value_note": "ranges resident in HBM, spans left in HBM; wall clock around lshift all-reduce + chain "
                          "send/recv + encode + metadata all-gather, max over ranks",
            "e2e": {"value": chsamp / (e2e_ms * 1e-3) / 1e6, "ms_per_step": e2e_ms, "unit": "M channel-samples/s",
                    "h2d_bytes_per_step": (se.up_hi - se.up_lo) * nch * 4, "d2h_bytes_per_step": int(res.total_bytes),
                    "bytes_are": "of rank 0", "destination": "one shared host mapping (/dev/shm), page-locked in every rank"},
            "chain_wait_ms_last_rank": chain_wait_ms,
            "collectives": ["all_reduce(MIN) 1 x int64 (trailing zeros -> offset_lshift)",
                            "send/recv 1 x int64 between consecutive ranks (segment chain)",
                            "all_gather 4 x int64 (blocks, bytes, max block, max bit rate)"],
            "payload_through_collectives": False,
        }
        if D.rank == 0:
            rec["single_gpu_ms"] = single_ms
            rec["speedup_vs_single_gpu_device_resident"] = single_ms / dev_ms if single_ms else None
            rec["stitched_stream_equals_single_gpu_stream"] = same
            rec["stream_md5"] = want_md5
            rec["stream_bytes"] = want_size
        if reg == 0 or getattr(reg, "value", 1) == 0:
            torch.cuda.cudart().cudaHostUnregister(shared.address)
        shared.close(unlink=(D.rank == 0))
        se.close()
        del d_mine, d_out, h_mine
        torch.cuda.empty_cache()
        out[name] = rec
        if a.seconds is None and name == "C2" and not getattr(a, "strong_c3", True):
            break
    return out
