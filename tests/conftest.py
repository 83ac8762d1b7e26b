"""Shared fixtures.

CPU suite  (-m "not gpu"): oracle vs reference vs golden vectors, host logic, C-ABI symbol export,
                           kernel logic through the host simulator (tests/hostsim).
GPU suite  (-m gpu):       the parity tests proper - libsla_b200.so through its C ABI on a B200,
                           checked against the oracle / the prebuilt reference in oracle/_ref.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sla_b200 import capi, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
PRODUCT_SO = os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so")
HOSTSIM_SO = os.path.join(ROOT, "tests", "hostsim", "libsla_hostsim.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_stream():
    def load(name):
        with open(os.path.join(GOLDEN, name + ".sla"), "rb") as f:
            return f.read()
    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build("oracle")
    return binding.Oracle()


@pytest.fixture(scope="session")
def reflib():
    """The unmodified reference (prebuilt into oracle/_ref; rebuilt when /root/reference exists)."""
    from oracle import binding
    path = os.path.join(ROOT, "oracle", "_ref", "libsla_ref.so")
    if os.path.isdir("/root/reference"):
        binding.build("ref")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libsla_ref.so not built and /root/reference absent")
    return binding.reference_library()


@pytest.fixture(scope="session")
def refwb():
    from oracle import binding
    path = os.path.join(ROOT, "oracle", "_ref", "libsla_ref_wb.so")
    if not os.path.exists(path):
        if not os.path.isdir("/root/reference"):
            pytest.skip("white-box reference not built")
        binding.build("ref")
    return binding.RefWhitebox()


@pytest.fixture(scope="session")
def hostsim():
    subprocess.run(["make", "-s", "-C", ROOT, "hostsim"], check=True)
    return capi.SLALibrary(HOSTSIM_SO)


@pytest.fixture(scope="session")
def product():
    """libsla_b200.so through its C ABI; GPU tests fail loudly if it is missing."""
    if not _has_gpu():
        pytest.skip("no CUDA device")
    assert os.path.exists(PRODUCT_SO), "libsla_b200.so missing: run `make product`"
    return capi.SLALibrary(PRODUCT_SO)


def pcm_md5(pcm: np.ndarray) -> str:
    return hashlib.md5(np.ascontiguousarray(pcm).tobytes()).hexdigest()


def multi_silence(n=100000, seed=11):
    """zero runs that end off the 1024 grid, so that the leading-silence rule re-bases the segment grid
    more than once (SLAEncoder.c:393-408) and later silences start from unaligned segment starts"""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = (6000 * np.sin(2 * np.pi * 440 * t / 44100) + rng.normal(0, 300, n)).astype(np.int64)
    pcm = np.stack([x, (0.8 * x + rng.normal(0, 200, n)).astype(np.int64)])
    for a, b in ((5000, 9000), (30011, 47000), (60000, 75555), (90000, 92047)):
        pcm[:, a:b] = 0
    return (np.clip(pcm, -32768, 32767) << 16).astype(np.int32)


def signal_set():
    """(name, pcm, bits, rate) used by several suites; small enough for the oracle in seconds."""
    return [
        ("multi_silence", multi_silence(), 16, 44100),
        ("s16_special", synth.synth_pcm(2, 60000, 16, 44100, 0, clear_low_bits=4), 16, 44100),
        ("s24_impulsive", synth.impulsive_24bit(30000), 24, 96000),
        ("ch8_24bit", synth.synth_pcm(8, 20000, 24, 48000, 3, specials=False), 24, 48000),
        ("mono16_odd", synth.synth_pcm(1, 30001, 16, 44100, 5, specials=False), 16, 44100),
    ]
