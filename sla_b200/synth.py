"""Deterministic synthetic PCM for the benchmark and parity tests.

Per channel: a few harmonic tones with slow envelopes plus AR(2)-coloured Gaussian noise, with
inter-channel correlation so that mid/side matters; peak about -1 dBFS; rounded to ``bits`` and
left-justified to int32 (the SLA API convention, reference src/wav.c:392-417).  Optional special
passages exercise every block type: an all-zero run (SILENT blocks and the leading-silence
partition rule), a full-scale white-noise stretch (RAW blocks), a +-few-LSB stretch (fixed-Golomb
mode) and cleared low bits (offset_lshift > 0).
"""
from __future__ import annotations

import numpy as np
from scipy.signal import lfilter

SEED_BASE = 0x51A0000


def synth_pcm(num_channels: int, num_samples: int, bits: int, rate: int, file_index: int = 0,
              specials: bool = True, clear_low_bits: int = 0) -> np.ndarray:
    """Returns int32 [num_channels, num_samples], left-justified to 32 bits."""
    rng = np.random.default_rng(SEED_BASE + file_index)
    n = num_samples
    t = np.arange(n, dtype=np.float64) / rate
    full = float(2 ** (bits - 1) - 1)

    def colour():
        r1, r2 = rng.uniform(0.8, 0.98, 2)
        th = rng.uniform(0.02, 0.6)
        a = [1.0, -(r1 + r2) * np.cos(th), r1 * r2]
        y = lfilter([1.0], a, rng.standard_normal(n))
        return y / (np.std(y) + 1e-12)

    common = np.zeros(n)
    for _ in range(int(rng.integers(3, 7))):
        f0 = rng.uniform(55.0, 880.0)
        env = 0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.05, 0.5) * t + rng.uniform(0, 6.28))
        for h in range(1, 4):
            common += env * (0.6 ** h) * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28))
    common /= (np.max(np.abs(common)) + 1e-12)
    common = 0.7 * common + 0.12 * colour()

    out = np.empty((num_channels, n), dtype=np.float64)
    for ch in range(num_channels):
        rho = rng.uniform(0.6, 0.95)
        own = 0.7 * np.sin(2 * np.pi * rng.uniform(110.0, 1760.0) * t) * 0.3 + 0.12 * colour()
        out[ch] = rho * common + (1.0 - rho) * own
    out *= (10 ** (-1.0 / 20)) / (np.max(np.abs(out)) + 1e-12)
    pcm = np.rint(out * full).astype(np.int64)

    if specials and n >= 8 * 4096:
        u = n // 16
        # (1) zero run (about 3 s at 44.1 kHz when the file is long enough), deliberately off-grid
        z0, z1 = 3 * u + 137, 3 * u + 137 + min(max(3 * rate, 5000), 3 * u)
        pcm[:, z0:z1] = 0
        # (2) full-scale white noise -> RAW blocks
        w0, w1 = 8 * u, 8 * u + min(rate, 2 * u)
        pcm[:, w0:w1] = rng.integers(-int(full), int(full) + 1, size=(num_channels, w1 - w0))
        # (3) very quiet passage -> fixed-Golomb mode
        q0, q1 = 12 * u, 12 * u + min(rate, 2 * u)
        pcm[:, q0:q1] = rng.integers(-4, 5, size=(num_channels, q1 - q0))
    if clear_low_bits:
        pcm = (pcm >> clear_low_bits) << clear_low_bits
    pcm = np.clip(pcm, -(2 ** (bits - 1)), 2 ** (bits - 1) - 1)
    return (pcm << (32 - bits)).astype(np.int32)


def impulsive_24bit(num_samples: int = 40000, rate: int = 96000, seed: int = 7) -> np.ndarray:
    """24-bit stereo pulse train through a resonator plus sparse huge impulses: drives the 3-tap
    long-term predictor, gamma escapes, rshift > 0 and the uint32 wrap in the Rice update
    (SURVEY.md section 8c fixture ii)."""
    rng = np.random.default_rng(seed)
    n = num_samples
    exc = np.zeros(n)
    exc[::147] = 1.0
    a = [1.0, -1.8 * np.cos(0.21), 0.81]
    body = lfilter([1.0], a, exc)
    body /= np.max(np.abs(body))
    pcm = np.empty((2, n), dtype=np.int64)
    for ch in range(2):
        x = 0.5 * body * (0.9 if ch else 1.0) + 0.002 * rng.standard_normal(n)
        pcm[ch] = np.rint(x * (2 ** 23 - 1))
    for pos in rng.integers(1000, n - 1000, 12):
        pcm[:, pos] = rng.choice([-1, 1]) * rng.integers(3_000_000, 8_000_000)
    pcm = np.clip(pcm, -(2 ** 23), 2 ** 23 - 1)
    return (pcm << 8).astype(np.int32)


def synth_long(num_channels: int, num_samples: int, bits: int, rate: int, file_index: int = 0,
               tile_seconds: int = 60, distinct_tiles: int = 6, out: np.ndarray | None = None) -> np.ndarray:
    """Hour-scale files: `distinct_tiles` independently synthesised tiles (each with the special
    passages), cycled with a per-repeat circular shift and channel polarity so that no two minutes
    are identical.  Deterministic in (file_index, shape).  `out` may be a preallocated (pinned) array."""
    tile_n = tile_seconds * rate
    if num_samples <= tile_n:
        pcm = synth_pcm(num_channels, num_samples, bits, rate, file_index)
        if out is None:
            return pcm
        out[:] = pcm
        return out
    tiles = [synth_pcm(num_channels, tile_n, bits, rate, file_index * 1000 + t) for t in range(distinct_tiles)]
    if out is None:
        out = np.empty((num_channels, num_samples), dtype=np.int32)
    rng = np.random.default_rng(SEED_BASE ^ (file_index + 77))
    pos, rep = 0, 0
    while pos < num_samples:
        tile = tiles[rep % distinct_tiles]
        if rep >= distinct_tiles:
            tile = np.roll(tile, int(rng.integers(1, tile_n - 1)), axis=1)
            if rng.integers(0, 2):
                tile = -np.maximum(tile, -(2 ** 31 - 2 ** (32 - bits)))      # keep -x representable
        take = min(tile_n, num_samples - pos)
        out[:, pos:pos + take] = tile[:, :take]
        pos += take
        rep += 1
    return out
