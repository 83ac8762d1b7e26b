"""The named workloads of BASELINE.json (configs 2-5) as deterministic synthetic signals, shared by
bench.py (both arms), the tools and the GPU tests so that every leg works on the same samples.

  C2  1 file per GPU, stereo 16-bit 44.1 kHz, 1 h, preset 2
  C3  1 file per GPU, stereo 24-bit 96 kHz, 1 h, preset 4 (PARCOR 32, long-term 3, LMS 8, 16384-sample blocks)
  C4  8-channel 24-bit 48 kHz files of 10 min, preset 2 (mid/side off), file i on GPU i mod N, 8 files per GPU
      (64 files on 8 GPUs)
  C5  a corpus of 10 000 stereo 16-bit 44.1 kHz files of 3-30 s, presets {0, 2, 4} mixed, file i on GPU i mod N

Host-side Python only (numpy); no codec logic.
"""
from __future__ import annotations

import numpy as np

from . import capi, synth

CONFIGS = {
    "C2": dict(channels=2, bits=16, rate=44100, seconds=3600, preset=2),
    "C3": dict(channels=2, bits=24, rate=96000, seconds=3600, preset=4),
    "C4": dict(channels=8, bits=24, rate=48000, seconds=600, preset=2, files_per_gpu=8),
    "C5": dict(channels=2, bits=16, rate=44100, files=10000, presets=(0, 2, 4), min_seconds=3, max_seconds=30,
               pool=16),
}


def describe(name: str, world: int = 1, **over) -> dict:
    """The `config` object both bench arms print for a workload (identical dicts -> same_config)."""
    c = dict(CONFIGS[name], **over)
    if name in ("C2", "C3"):
        n = c["seconds"] * c["rate"]
        return {"workload": f"{name}: synthetic {c['bits']}-bit {c['channels']}ch {c['rate']} Hz, {c['seconds']} s per GPU, "
                            f"preset {c['preset']}",
                "files_per_gpu": 1, "channel_samples_per_gpu": n * c["channels"],
                "signal": "sla_b200.synth.synth_long(file_index = rank): 60 s tiles of tones + AR(2) noise with silence, "
                          "full-scale noise and near-silent passages"}
    if name == "C4":
        n = c["seconds"] * c["rate"]
        return {"workload": f"C4: synthetic {c['bits']}-bit {c['channels']}ch {c['rate']} Hz files of {c['seconds']} s, preset "
                            f"{c['preset']} (mid/side off), file i on GPU i mod N",
                "files_per_gpu": c["files_per_gpu"], "files": c["files_per_gpu"] * world,
                "channel_samples_per_gpu": n * c["channels"] * c["files_per_gpu"],
                "signal": "one seeded 10-minute signal per GPU; its files are channel rotations + time shifts of it"}
    c5 = corpus_index(c)
    return {"workload": f"C5: decode of a corpus of {c['files']} stereo 16-bit 44.1 kHz files of {c['min_seconds']}-"
                        f"{c['max_seconds']} s, presets {list(c['presets'])} mixed, file i on GPU i mod N",
            "files": c["files"], "channel_samples": int(c5["frames"].sum()) * c["channels"],
            "signal": f"file k = a slice of one of {c['pool']} seeded 30 s signals (offset, length, preset from rng(k))"}


# ----------------------------------------------------------------------------- C2 / C3
def long_file(name: str, rank: int = 0, out: np.ndarray | None = None, seconds: int | None = None) -> np.ndarray:
    c = CONFIGS[name]
    n = (seconds or c["seconds"]) * c["rate"]
    return synth.synth_long(c["channels"], n, c["bits"], c["rate"], file_index=rank, out=out)


def sample_ranges(num_samples: int, max_block: int, count: int, length: int) -> list:
    """`count` ranges of `length` samples (rounded to whole blocks) spread over the file, each starting on a
    multiple of max_num_block_samples (how a single file is cut for the CPU baseline)."""
    length = max(max_block, (length // max_block) * max_block)
    length = min(length, (num_samples // max_block) * max_block)
    slots = max(1, (num_samples - length) // max_block + 1)
    out = []
    for k in range(count):
        start = ((k * slots) // count) * max_block
        out.append((start, min(length, num_samples - start)))
    return out


# ----------------------------------------------------------------------------- C4
def c4_base(rank: int, seconds: int | None = None) -> np.ndarray:
    c = CONFIGS["C4"]
    n = (seconds or c["seconds"]) * c["rate"]
    return synth.synth_long(c["channels"], n, c["bits"], c["rate"], file_index=4000 + rank)


def c4_file(base: np.ndarray, index: int) -> np.ndarray:
    """file `index` of a GPU: channels rotated by index, time rotated by index * 7919 samples"""
    return np.ascontiguousarray(np.roll(np.roll(base, index, axis=0), index * 7919, axis=1))


# ----------------------------------------------------------------------------- C5
def corpus_index(c: dict | None = None) -> dict:
    """per-file (pool signal, first frame, frames, preset) of the whole corpus; deterministic"""
    c = c or CONFIGS["C5"]
    nf, rate = c["files"], c["rate"]
    pool_frames = c["max_seconds"] * rate
    rng = np.random.default_rng(0xC5C5)
    frames = rng.integers(c["min_seconds"] * rate, c["max_seconds"] * rate + 1, nf)
    first = (rng.random(nf) * (pool_frames - frames + 1)).astype(np.int64)
    preset = np.asarray(c["presets"])[rng.integers(0, len(c["presets"]), nf)]
    return {"pool": np.arange(nf) % c["pool"], "first": first, "frames": frames, "preset": preset}


def corpus_pool(c: dict | None = None) -> list:
    """the pool signals as interleaved int16 frames [frames, 2] (WAV data-chunk layout)"""
    c = c or CONFIGS["C5"]
    out = []
    for j in range(c["pool"]):
        pcm = synth.synth_pcm(c["channels"], c["max_seconds"] * c["rate"], c["bits"], c["rate"], 5000 + j)
        out.append(np.ascontiguousarray((pcm >> 16).astype(np.int16).T))
    return out


def planar_of(frames16: np.ndarray) -> np.ndarray:
    """interleaved int16 frames -> the planar left-justified int32 planes the C API takes"""
    return np.ascontiguousarray(frames16.T.astype(np.int32) << 16)


def pcm24_bytes(planar: np.ndarray) -> np.ndarray:
    """planar left-justified int32 -> interleaved packed 24-bit little-endian bytes (uint8 array)"""
    return np.frombuffer(capi.planar_to_pcm(planar, 24), dtype=np.uint8)
